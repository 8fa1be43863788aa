"""Per-launch event profile of one decoder forward (run on the GPU box):
    python tools/profile_layers.py [--batch 64] [--frames 200] [--precision bf16] > gpurun_out/layers.txt
Groups launches by (category, algorithmic flops, bytes) = layer shape and prints time and throughput."""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from styletts2_lite_b200 import synth  # noqa: E402
from styletts2_lite_b200.config import DecoderConfig  # noqa: E402
from styletts2_lite_b200.decoder import B200Decoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--frames", type=int, default=200)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--variant", default="hifigan")
a = ap.parse_args()
cfg = {"hifigan": DecoderConfig.hifigan, "istftnet": DecoderConfig.istftnet, "vocos": DecoderConfig.vocos}[a.variant]()
m = B200Decoder(cfg, a.precision)
m.load_state_dict(synth.make_state_dict(cfg, 0, True))
m = m.cuda().eval()
inp = {k: v.cuda() for k, v in synth.make_inputs(a.batch, a.frames, 1002, cfg, with_noise=False).items()}
ap_reps = 5
for _ in range(4):
    m(inp["asr"], inp["F0_curve"], inp["N"], inp["s"], seed=1)
m.set_profiling(True)
runs = []
for i in range(ap_reps):     # per-launch median over 5 forwards (clock / power state noise)
    m(inp["asr"], inp["F0_curve"], inp["N"], inp["s"], seed=2 + i)
    runs.append(m.get_profile_launches())
recs = []
for j in range(len(runs[0])):
    ms = sorted(r[j][1] for r in runs)[ap_reps // 2]
    recs.append((runs[0][j][0], ms, runs[0][j][2], runs[0][j][3]))
groups = collections.OrderedDict()
for cat, ms, fl, by in recs:
    g = groups.setdefault((cat, fl, by), [0, 0.0])
    g[0] += 1
    g[1] += ms
tot = sum(r[1] for r in recs)
print("total %.3f ms, %d launches" % (tot, len(recs)))
print("%-11s %5s %9s %8s %9s %9s %6s" % ("category", "n", "ms_total", "ms_each", "TFLOP/s", "GB/s", "share"))
for (cat, fl, by), (n, ms) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
    print("%-11s %5d %9.3f %8.4f %9.1f %9.1f %6.3f   GF=%.1f MB=%.1f" % (cat, n, ms, ms / n, fl * n / ms / 1e9 if ms else 0,
                                                                  by * n / ms / 1e6 if ms else 0, ms / tot, fl / 1e9, by / 1e6))
