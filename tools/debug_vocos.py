"""Tap-by-tap comparison of the vocos variant against the oracle at a given (B, T) (run on the GPU box)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from styletts2_lite_b200 import synth  # noqa: E402
from styletts2_lite_b200.config import DecoderConfig  # noqa: E402
from styletts2_lite_b200.decoder import B200Decoder  # noqa: E402
from oracle import decoder_np as O  # noqa: E402

B, T, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cfg = DecoderConfig.vocos()
sdt = synth.make_state_dict(cfg, 0, True)
m = B200Decoder(cfg, "fp32")
m.load_state_dict(sdt)
m = m.cuda().eval()
inp = synth.make_inputs(B, T, seed, cfg, with_noise=False)
names = {"encode": (T, 1024), "decode.0": (T, 1024), "decode.2": (T, 1024), "decode.3": (2 * T, 512), "generator.convnext.0.dwconv": (2 * T, 512),
         "generator.convnext.0": (2 * T, 512), "generator.convnext.7": (2 * T, 512), "generator.final_layer_norm": (2 * T, 512),
         "generator.stft.out": (2 * T, 1202)}
bufs = {n: m.set_tap(n, B, r, c) for n, (r, c) in names.items()}
with torch.no_grad():
    out = m(inp["asr"].cuda(), inp["F0_curve"].cuda(), inp["N"].cuda(), inp["s"].cuda()).cpu().numpy()
taps = {}
ref = O.decoder_forward({k: v.numpy() for k, v in sdt.items()}, cfg, inp["asr"].numpy(), inp["F0_curve"].numpy(), inp["N"].numpy(),
                        inp["s"].numpy(), None, taps=taps)
for n in names:
    a = bufs[n].cpu().numpy()
    r = taps[n]
    if r.shape != a.shape:
        r = r.transpose(0, 2, 1)
    print("%-32s max|ref| %.4f  max err %.3e" % (n, np.abs(r).max(), np.abs(a - r).max()))
print("out  max|ref| %.4f max err %.3e" % (np.abs(ref).max(), np.abs(out - ref).max()))
