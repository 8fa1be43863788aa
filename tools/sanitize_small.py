"""Small forwards of both decoder variants in every precision (run under compute-sanitizer on the GPU box):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
Covers the TMA pipeline kernel (ragged tiles, residual/accumulate rings, ConvTranspose path), the register-staged fused
kernel (256-channel layers, widest ups, istft mirror), the plain tensor-core conv and the fp32 SIMT path."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from styletts2_lite_b200 import synth  # noqa: E402
from styletts2_lite_b200.config import DecoderConfig  # noqa: E402
from styletts2_lite_b200.decoder import B200Decoder  # noqa: E402

for variant in ("hifigan", "istftnet"):
    cfg = DecoderConfig.hifigan() if variant == "hifigan" else DecoderConfig.istftnet()
    m = B200Decoder(cfg, "fp32")
    m.load_state_dict(synth.make_state_dict(cfg, 0, True))
    m = m.cuda().eval()
    for B, T in ((2, 5), (1, 37)):
        inp = {k: v.cuda() for k, v in synth.make_inputs(B, T, 7, cfg, with_noise=False).items()}
        for prec in ("fp32", "bf16", "fp16"):
            with torch.no_grad():
                out = m(inp["asr"], inp["F0_curve"], inp["N"], inp["s"], seed=3, precision=prec)
            torch.cuda.synchronize()
            assert bool(torch.isfinite(out).all()), (variant, B, T, prec)
            print(variant, B, T, prec, tuple(out.shape), float(out.abs().max()))
print("sanitize_small: ok")
