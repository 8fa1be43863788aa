"""A few decoder forwards at the bench shape for ncu captures (run on the GPU box):
    python tools/one_forward.py [--forwards 4] [--batch 64] [--frames 200] [--precision bf16] [--variant hifigan]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from styletts2_lite_b200 import synth  # noqa: E402
from styletts2_lite_b200.config import DecoderConfig  # noqa: E402
from styletts2_lite_b200.decoder import B200Decoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--forwards", type=int, default=4)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--frames", type=int, default=200)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--variant", default="hifigan")
a = ap.parse_args()
cfg = DecoderConfig.hifigan() if a.variant == "hifigan" else DecoderConfig.istftnet()
m = B200Decoder(cfg, a.precision)
m.load_state_dict(synth.make_state_dict(cfg, 0, True))
m = m.cuda().eval()
inp = {k: v.cuda() for k, v in synth.make_inputs(a.batch, a.frames, 1002, cfg, with_noise=False).items()}
for i in range(a.forwards):
    out = m(inp["asr"], inp["F0_curve"], inp["N"], inp["s"], seed=1 + i)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.abs().max()), "launches", m.last_launch_count())
