"""BASELINE.json configs[2] on the GPU box: inference.py:239-270 as one pass
    TextEncoder -> duration half (DurationEncoder, lstm, duration_proj) -> durations -> length regulation of d and t_en
    -> F0Ntrain -> Decoder
for 32 utterances x 8 s (T=320, 64 tokens each, seeded integer durations as SURVEY.md 8(d) cfg 3 prescribes).
Times the pass with CUDA events (inputs resident), reports the three parts and checks the 16-bit result against the fp32 path.
    python tools/bench_chain.py [--batch 32] [--frames 320] [--tokens 64] [--iters 10]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from styletts2_lite_b200 import length_regulator as LR, synth  # noqa: E402
from styletts2_lite_b200.config import DecoderConfig  # noqa: E402
from styletts2_lite_b200.decoder import B200Decoder  # noqa: E402
from styletts2_lite_b200.predictor import B200F0NPredictor  # noqa: E402
from styletts2_lite_b200.text_encoder import B200TextEncoder  # noqa: E402


def run_chain(B=32, L=64, T=320, precision="bf16", iters=10, graph=False):
    """One pass = inference.py:239-270 on the GPU for B utterances of L tokens / T frames; returns the result dict."""
    cfg = DecoderConfig.hifigan()
    dec = B200Decoder(cfg, precision)
    dec.load_state_dict(synth.make_state_dict(cfg, 0, True))
    dec = dec.cuda().eval()
    pred = B200F0NPredictor(precision=precision, duration=True)
    pred.load_state_dict(synth.make_predictor_state_dict(seed=0, duration=True))
    pred = pred.cuda().eval()
    text = B200TextEncoder(precision=precision)
    text.load_state_dict(synth.make_text_state_dict(seed=0))
    text = text.cuda().eval()
    tokens = synth.make_tokens(B, L, seed=5300).cuda()
    ci = synth.make_chain_inputs(B, L, T, seed=3100)
    dur = ci["dur"].to(torch.int32).cuda()
    # random-init weights predict ~25 frames for every token, so the seeded integer durations stand in for round(duration)
    # (SURVEY.md 8(d) cfg 3); the duration half still runs and its `d` is what gets regulated
    s, noise = ci["s"].cuda(), ci["noise"].cuda()

    def chain(prec, seed=None, tape=None, ev=None):
        gr = graph and tape is None                                            # graph replay of every module (timed passes only)
        t_en = text(tokens, precision=prec, cuda_graph=gr)                     # inference.py:239
        if ev: ev[5].record()
        d, duration = pred.predict_duration(t_en, s, precision=prec, cuda_graph=gr)   # inference.py:242-245
        LR.round_durations(duration)                                           # inference.py:257 (result replaced by `dur`)
        if ev: ev[1].record()
        en = LR.length_regulate(d.transpose(1, 2).contiguous(), dur, T)        # inference.py:266
        asr = LR.length_regulate(t_en, dur, T)                                 # inference.py:269
        if ev: ev[2].record()
        f0, n = pred.F0Ntrain(en, s, precision=prec, cuda_graph=gr)            # inference.py:267
        if ev: ev[3].record()
        out = dec(asr, f0, n, s, noise=tape, seed=seed, precision=prec, cuda_graph=gr)   # inference.py:270
        if ev: ev[4].record()
        return out

    with torch.no_grad():
        ref = chain("fp32", tape=noise)
        got = chain(precision, tape=noise)
        e = (got - ref).double()
        snr = float(10 * torch.log10((ref.double() ** 2).sum() / (e ** 2).sum()))
        for i in range(3):
            chain(precision, seed=i)
        torch.cuda.synchronize()
        parts = np.zeros(5)
        tot = 0.0
        for i in range(iters):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            ev[0].record()
            chain(precision, seed=10 + i, ev=ev)
            torch.cuda.synchronize()
            parts += [ev[0].elapsed_time(ev[5]), ev[5].elapsed_time(ev[1])] + [ev[j].elapsed_time(ev[j + 1]) for j in range(1, 4)]
            tot += ev[0].elapsed_time(ev[4])
    secs = B * T / 40.0
    return {"path": "TextEncoder -> duration half -> length regulator -> F0Ntrain -> Decoder (inference.py:239-270)", "batch": B,
            "tokens": L, "frames": T, "audio_s": secs, "precision": precision, "cuda_graphs": bool(graph),
            "ms": round(tot / iters, 3),
            "audio_s_per_s": round(secs / (tot / iters) * 1e3, 1),
            "ms_parts": {"text_encoder": round(parts[0] / iters, 4), "duration_half": round(parts[1] / iters, 4),
                         "length_regulator": round(parts[2] / iters, 4), "f0n_predictor": round(parts[3] / iters, 4),
                         "decoder": round(parts[4] / iters, 4)},
            "snr_db_vs_fp32_path": round(snr, 2)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--frames", type=int, default=320)
    ap.add_argument("--tokens", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--graph", action="store_true", help="replay per-module CUDA graphs in the timed passes")
    a = ap.parse_args()
    print(json.dumps(run_chain(a.batch, a.tokens, a.frames, a.precision, a.iters, a.graph)))
