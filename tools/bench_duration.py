"""Duration half of the predictor (inference.py:242-245) on the GPU box: time per call and per-category
profile at the cfg-3 shape (32 utterances x 64 tokens), with the torch-CPU restatement beside it.
    python tools/bench_duration.py [--batch 32] [--tokens 64] [--precision fp16] [--cpu]
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
ap.add_argument("--tokens", type=int, default=64)
ap.add_argument("--precision", default="fp16")
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--cpu", action="store_true")
a = ap.parse_args()

m = B200F0NPredictor(precision=a.precision, duration=True)
m.load_state_dict(synth.make_predictor_state_dict(seed=0, duration=True))
m = m.cuda().eval()
inp = {k: v.cuda() for k, v in synth.make_duration_inputs(a.batch, a.tokens, seed=4100).items()}
with torch.no_grad():
    for _ in range(3):
        m.predict_duration(inp["t_en"], inp["s"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        m.predict_duration(inp["t_en"], inp["s"])
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
lib = _lib.load()
_lib.check(lib.st2_decoder_set_profiling(m._handle, 1))
with torch.no_grad():
    m.predict_duration(inp["t_en"], inp["s"])
n = lib.st2_profile_num_categories()
pm, fl, by = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
ln = (C.c_int64 * n)()
_lib.check(lib.st2_decoder_get_profile(m._handle, pm, ln, fl, by))
_lib.check(lib.st2_decoder_set_profiling(m._handle, 0))
cats = {lib.st2_profile_category_name(i).decode(): {"ms": round(pm[i], 4), "launches": int(ln[i])} for i in range(n) if ln[i]}
out = {"path": "duration half (inference.py:242-245)", "batch": a.batch, "tokens": a.tokens, "precision": a.precision,
       "ms": round(ms, 4), "tokens_per_s": round(a.batch * a.tokens / ms * 1e3, 1), "launches": m.last_launch_count(), "categories": cats}
if a.cpu:
    from oracle import predictor_torch as PT
    sd = synth.make_predictor_state_dict(seed=0, duration=True)
    bs = min(a.batch, 8)
    x, s = inp["t_en"][:bs].cpu(), inp["s"][:bs].cpu()
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        PT.predict_duration(sd, x, s)
        t0 = time.perf_counter()
        PT.predict_duration(sd, x, s)
        dt = time.perf_counter() - t0
    out["cpu"] = {"sample": "%d x %d tokens" % (bs, a.tokens), "ms": round(dt * 1e3, 2), "cores": os.cpu_count(),
                  "tokens_per_s": round(bs * a.tokens / dt, 1)}
print(json.dumps(out))
