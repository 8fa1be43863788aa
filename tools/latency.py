"""Single-utterance latency of the drop-in decoder, eager launches vs CUDA-graph replay (run on the GPU box):
    python tools/latency.py [--frames 120] [--batch 1] [--precision bf16]
The reference synthesises one sentence at a time (inference.py:234-272); at that size a forward is ~270 short launches."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from styletts2_lite_b200 import synth  # noqa: E402
from styletts2_lite_b200.config import DecoderConfig  # noqa: E402
from styletts2_lite_b200.decoder import B200Decoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--frames", type=int, default=120)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--variant", default="hifigan")
ap.add_argument("--iters", type=int, default=50)
a = ap.parse_args()
cfg = DecoderConfig.hifigan() if a.variant == "hifigan" else DecoderConfig.istftnet()
m = B200Decoder(cfg, a.precision)
m.load_state_dict(synth.make_state_dict(cfg, 0, True))
m = m.cuda().eval()
inp = {k: v.cuda() for k, v in synth.make_inputs(a.batch, a.frames, 1002, cfg, with_noise=False).items()}
res = {}
for mode in (False, True):
    with torch.no_grad():
        for i in range(5):
            m(inp["asr"], inp["F0_curve"], inp["N"], inp["s"], seed=i, cuda_graph=mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.iters):
            m(inp["asr"], inp["F0_curve"], inp["N"], inp["s"], seed=100 + i, cuda_graph=mode)
        e1.record()
        torch.cuda.synchronize()
    res["graph" if mode else "eager"] = e0.elapsed_time(e1) / a.iters
secs = a.batch * a.frames * 600 / 24000
print(json.dumps({"variant": a.variant, "batch": a.batch, "frames": a.frames, "audio_s": secs, "precision": a.precision,
                  "eager_ms": round(res["eager"], 4), "graph_ms": round(res["graph"], 4),
                  "eager_rtf_inv": round(secs / res["eager"] * 1e3, 1), "graph_rtf_inv": round(secs / res["graph"] * 1e3, 1)}))
