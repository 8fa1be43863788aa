"""F0Ntrain on the GPU box: time per call, per-kernel-category profile and, beside it, the numpy/torch
CPU restatement on the host cores (bounded sample).  cfg-3 shape by default (32 utterances x 8 s).
    python tools/bench_predictor.py [--batch 32] [--frames 320] [--precision fp16] [--cpu]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from styletts2_lite_b200 import _lib, synth  # noqa: E402
from styletts2_lite_b200.predictor import B200F0NPredictor  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--frames", type=int, default=320)
ap.add_argument("--precision", default="fp16")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--cpu", action="store_true", help="also time the CPU restatement (torch CPU kernels) on a bounded sample")
a = ap.parse_args()

m = B200F0NPredictor(precision=a.precision)
m.load_state_dict(synth.make_predictor_state_dict(seed=0))
m = m.cuda().eval()
inp = {k: v.cuda() for k, v in synth.make_predictor_inputs(a.batch, a.frames, seed=2300).items()}
with torch.no_grad():
    for _ in range(3):
        m.F0Ntrain(inp["en"], inp["s"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        m.F0Ntrain(inp["en"], inp["s"])
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
lib = _lib.load()
_lib.check(lib.st2_decoder_set_profiling(m._handle, 1))
with torch.no_grad():
    m.F0Ntrain(inp["en"], inp["s"])
n = lib.st2_profile_num_categories()
pm, fl, by = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
ln = (C.c_int64 * n)()
_lib.check(lib.st2_decoder_get_profile(m._handle, pm, ln, fl, by))
_lib.check(lib.st2_decoder_set_profiling(m._handle, 0))
cats = {lib.st2_profile_category_name(i).decode(): {"ms": round(pm[i], 4), "launches": int(ln[i])} for i in range(n) if ln[i]}
secs = a.batch * a.frames / 40.0
out = {"path": "F0Ntrain", "batch": a.batch, "frames": a.frames, "audio_s": secs, "precision": a.precision, "ms": round(ms, 4),
       "audio_s_per_s": round(secs / ms * 1e3, 1), "launches": m.last_launch_count(), "categories": cats}
if a.cpu:
    # torch CPU kernels on the reference's own layer types (nn.LSTM / conv1d / instance_norm): what the reference executes
    from oracle import predictor_torch as PT
    sd = synth.make_predictor_state_dict(seed=0)
    bs = min(a.batch, 4)
    x, s = inp["en"][:bs].cpu(), inp["s"][:bs].cpu()
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        PT.f0n_train(sd, x, s)
        t0 = time.perf_counter()
        PT.f0n_train(sd, x, s)
        dt = time.perf_counter() - t0
    out["cpu"] = {"sample": "%d x %d frames" % (bs, a.frames), "ms": round(dt * 1e3, 2), "cores": os.cpu_count(),
                  "audio_s_per_s": round(bs * a.frames / 40.0 / dt, 1)}
print(json.dumps(out))
