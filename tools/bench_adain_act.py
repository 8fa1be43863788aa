"""Unit bench of the standalone AdaIN + activation path (st2_adain_act: in_stats -> adain_coef -> affine_act) at the generator
shapes of BASELINE configs[1] (run on the GPU box):
    python tools/bench_adain_act.py
SURVEY.md section 7's minimum slice / VERDICT r1 weak item 7: the fused AdaIN kernels are HBM-bound and should sit at >= 0.70 of
the measured copy peak.  Algorithmic bytes: statistics pass reads x once (4 B / element), the affine pass reads x again and writes y
(4 + 2 B / element for a 16-bit output).  Every call is timed with CUDA events over 20 iterations after 3 warm-ups with a 256 MB
L2 flush between iterations outside the events; the three kernels are also timed one by one through the same C-ABI call with the
stages it does not need switched off (h = NULL skips the statistics pass)."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from styletts2_lite_b200 import _lib  # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
lib = _lib.load()
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(iters):
        flush.fill_(i & 0xFF)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


rows = []
for (B, T, Cc, act) in ((64, 120000, 32, "snake"), (64, 60000, 64, "snake"), (64, 20000, 128, "snake"), (64, 4000, 256, "snake"),
                        (64, 400, 1024, "lrelu")):
    x = torch.randn(B, T, Cc, device=dev)
    h = torch.randn(B, 2 * Cc, device=dev) * 0.1
    alpha = (0.6 + 0.8 * torch.rand(Cc, device=dev))
    y = torch.empty(B, T, Cc, device=dev, dtype=torch.bfloat16)
    scratch = torch.empty(_lib.check(lib.st2_adain_scratch_bytes(B, T, Cc)), dtype=torch.uint8, device=dev)

    def call(with_stats):
        _lib.check(lib.st2_adain_act(_lib.ptr(x), Cc, _lib.ptr(h) if with_stats else None, 2 * Cc if with_stats else 0, _lib.ptr(alpha),
                                     _lib.ACT[act], C.c_float(0.2), _lib.ptr(y), Cc, _lib.DTYPE["bf16"], B, T, Cc, _lib.ptr(scratch),
                                     stream), "adain_act")
    n = B * T * Cc
    ms_all = timed(lambda: call(True))
    ms_aff = timed(lambda: call(False))                 # coefficient kernel (identity) + affine pass only
    ms_stats = ms_all - ms_aff
    rows.append({"B": B, "T": T, "C": Cc, "act": act, "MB_in": round(n * 4 / 1e6, 1),
                 "all_ms": round(ms_all, 4), "all_gbs": round(n * 10 / ms_all / 1e6, 1), "all_frac": round(n * 10 / ms_all / 1e6 / PEAK, 3),
                 "affine_ms": round(ms_aff, 4), "affine_gbs": round(n * 6 / ms_aff / 1e6, 1), "affine_frac": round(n * 6 / ms_aff / 1e6 / PEAK, 3),
                 "stats_ms": round(ms_stats, 4), "stats_gbs": round(n * 4 / ms_stats / 1e6, 1), "stats_frac": round(n * 4 / ms_stats / 1e6 / PEAK, 3)})
    del x, y
    torch.cuda.empty_cache()
print(json.dumps({"peak_gbs": PEAK, "rows": rows}, indent=1))
