"""A paragraph -> PCM through pipeline.B200Synthesizer (run on the GPU box):
    python tools/bench_paragraph.py [--sentences 8] [--tokens 100] [--precision bf16]
The reference (StyleTTS2.generate, inference.py:303-319) walks the sentences one by one through every module; here the text
modules and the duration logic take the sentences as one padded batch, the regulator / F0Ntrain / Decoder run per sentence with graph
replay, and the trim / concatenate / normalise / PCM_16 step is one device pass.  Prints one JSON line: wall-clock ms per
paragraph (host clock around generate() + a synchronize, second and later calls: graphs captured), audio seconds, and the same
paragraph with the sentence loop the reference uses (B = 1 through every module, same library)."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from styletts2_lite_b200 import synth  # noqa: E402
from styletts2_lite_b200 import length_regulator as LR  # noqa: E402
from styletts2_lite_b200 import postprocess as PP  # noqa: E402
from styletts2_lite_b200.config import DecoderConfig  # noqa: E402
from styletts2_lite_b200.decoder import B200Decoder  # noqa: E402
from styletts2_lite_b200.pipeline import B200Synthesizer  # noqa: E402
from styletts2_lite_b200.predictor import B200F0NPredictor  # noqa: E402
from styletts2_lite_b200.text_encoder import B200TextEncoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sentences", type=int, default=8)
ap.add_argument("--tokens", type=int, default=100)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--graphs", action="store_true", help="graph replay for the per-sentence forwards (shapes repeat in this benchmark)")
a = ap.parse_args()
dev = torch.device("cuda")
te = B200TextEncoder(channels=512, kernel_size=5, depth=3, n_symbols=178)
te.load_state_dict(synth.make_text_state_dict(seed=0))
pr = B200F0NPredictor(style_dim=128, d_hid=512, nlayers=3, max_dur=50, dropout=0.2, duration=True)
pr.load_state_dict(synth.make_predictor_state_dict(seed=0, duration=True))
cfg = DecoderConfig.hifigan()
dec = B200Decoder(cfg, a.precision)
dec.load_state_dict(synth.make_state_dict(cfg, 0, True))
te, pr, dec = te.to(dev).eval(), pr.to(dev).eval(), dec.to(dev).eval()
g = torch.Generator().manual_seed(3)
lens = [int(x) for x in torch.randint(a.tokens // 2, a.tokens + 1, (a.sentences,), generator=g)]
lens[0] = a.tokens
sents = [synth.make_tokens(1, n, seed=6000 + i)[0] for i, n in enumerate(lens)]
s = synth.make_duration_inputs(1, 4, seed=4800)["s"].to(dev)
z = torch.randn(a.sentences, max(lens), generator=g).to(dev)
seeds = list(range(50, 50 + a.sentences))
syn = B200Synthesizer(te, pr, dec, precision=a.precision, cuda_graph=a.graphs)
if a.graphs:                                   # every sentence has its own frame count: room for all of them
    dec.max_graphs = pr._graphs.max_graphs = te._graphs.max_graphs = 4 * a.sentences


def batched():
    return syn.generate(sents, s, speed=1.0, stabilize=True, duration_noise=z, decoder_seeds=seeds)


def looped():
    """inference.py:303-319 with the same modules: one sentence at a time through everything, prev_d_mean handed on."""
    prev, waves = 0.0, []
    with torch.no_grad():
        for b, tok in enumerate(sents):
            n = lens[b]
            t_en = te(tok.unsqueeze(0).to(dev), precision=a.precision, cuda_graph=a.graphs)
            d, duration = pr.predict_duration(t_en, s, precision=a.precision, cuda_graph=a.graphs)
            duration, mean = LR.smooth_durations(duration, z[b:b + 1, :n].contiguous(), t=0.2, speed=1.0, prev_d_mean=prev)
            prev = float(mean[0])                                   # .item(): the reference returns duration.mean() to Python too
            pd, tot = LR.round_durations(duration)
            F = int(tot[0])
            asr = LR.length_regulate(t_en, pd, F)
            en = LR.length_regulate(d.transpose(1, 2).contiguous(), pd, F)
            f0, nn_ = pr.F0Ntrain(en, s, precision=a.precision, cuda_graph=a.graphs)
            waves.append(dec(asr, f0, nn_, s, seed=seeds[b], precision=a.precision, cuda_graph=a.graphs).reshape(-1))
        lengths = torch.tensor([w.numel() for w in waves], dtype=torch.int32, device=dev)
        batch = torch.zeros(len(waves), int(lengths.max()), device=dev)
        for b, w in enumerate(waves):
            batch[b, :w.numel()] = w
        return PP.assemble(batch, lengths)


def wall(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.iters):
        r, pcm = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / a.iters * 1e3, pcm


ms_b, pcm = wall(batched)
ms_l, pcm_l = wall(looped)
audio_s = pcm.numel() / 24000.0
print(json.dumps({"sentences": a.sentences, "tokens": lens, "precision": a.precision, "audio_s": round(audio_s, 2),
                  "graphs": bool(a.graphs), "batched_ms": round(ms_b, 2), "looped_ms": round(ms_l, 2), "audio_s_per_s_batched": round(audio_s / ms_b * 1e3, 1),
                  "pcm_samples": int(pcm.numel()), "same_length_as_loop": bool(pcm.numel() == pcm_l.numel())}))
