"""`Decoder` with the constructor of the reference's Modules/istftnet.py:660-667."""
from __future__ import annotations

from .config import DecoderConfig
from .decoder import B200Decoder


class Decoder(B200Decoder):
    def __init__(self, dim_in=512, F0_channel=512, style_dim=64, dim_out=80,
                 resblock_kernel_sizes=(3, 7, 11), upsample_rates=(10, 6),
                 upsample_initial_channel=512, resblock_dilation_sizes=((1, 3, 5), (1, 3, 5), (1, 3, 5)),
                 upsample_kernel_sizes=(20, 12), gen_istft_n_fft=20, gen_istft_hop_size=5, precision="fp32", fp16_storage=True):
        cfg = DecoderConfig(type="istftnet", dim_in=dim_in, style_dim=style_dim,
                            resblock_kernel_sizes=list(resblock_kernel_sizes), upsample_rates=list(upsample_rates),
                            upsample_initial_channel=upsample_initial_channel,
                            resblock_dilation_sizes=[list(d) for d in resblock_dilation_sizes],
                            upsample_kernel_sizes=list(upsample_kernel_sizes),
                            gen_istft_n_fft=gen_istft_n_fft, gen_istft_hop_size=gen_istft_hop_size)
        super().__init__(cfg, precision, fp16_storage)
