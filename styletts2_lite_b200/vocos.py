"""`Decoder` with the constructor of the reference's Modules/vocos.py:364-368 (the third `decoder.type`, models.py:555-562)."""
from __future__ import annotations

from .config import DecoderConfig
from .decoder import B200Decoder


class Decoder(B200Decoder):
    def __init__(self, dim_in=512, style_dim=64, dim_out=80, intermediate_dim=1536, num_layers=8, gen_istft_n_fft=1024,
                 gen_istft_hop_size=256, precision="fp32", fp16_storage=True):
        cfg = DecoderConfig(type="vocos", dim_in=dim_in, style_dim=style_dim, upsample_rates=[], upsample_kernel_sizes=[],
                            resblock_kernel_sizes=[], resblock_dilation_sizes=[], intermediate_dim=intermediate_dim,
                            num_layers=num_layers, gen_istft_n_fft=gen_istft_n_fft, gen_istft_hop_size=gen_istft_hop_size)
        super().__init__(cfg, precision, fp16_storage)
