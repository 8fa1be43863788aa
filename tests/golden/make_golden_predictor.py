"""Golden fixtures for SURVEY.md §8(f) N1 (`ProsodyPredictor.F0Ntrain`) and the chained cfg-3 slice, from the
UNMODIFIED reference.  Authoring container only (needs /root/reference):

    python tests/golden/make_golden_predictor.py

models.py imports `munch`, which is not installed: a 4-line attribute-dict shim stands in for it.
The reference ProsodyPredictor is built as at models.py:565 / inference.py:120 with config_example.yaml values
(style_dim 128, hidden_dim 512, n_layer 3, max_dur 50, dropout 0.2); the F0Ntrain sub-modules take the synthetic
weights of styletts2_lite_b200.synth.make_predictor_state_dict through the reference's own load_state_dict.

Fixtures
  f0n_B2_T6_w0_i2001.npz      F0Ntrain(en, s) on a tiny case, with the LSTM output and per-block taps
  f0n_B1_T120_w0_i2002.npz    F0Ntrain at the 3 s size of cfg 1 (outputs only)
  chain_B2_L9_T16_w0.npz      inference.py:257-270 chained: integer durations -> alignment matrix -> en / asr by the
                              reference's matmul -> F0Ntrain -> hifigan Decoder (noise tape), i.e. the cfg-3 data flow
                              after the text modules, at a size that fits a fixture.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

_m = types.ModuleType("munch")
_m.Munch = type("Munch", (dict,), {"__getattr__": dict.get, "__setattr__": dict.__setitem__})
sys.modules.setdefault("munch", _m)

from styletts2_lite_b200.config import DecoderConfig, PredictorConfig, F0N_PREFIXES, DUR_PREFIXES  # noqa: E402
from styletts2_lite_b200 import synth  # noqa: E402


def build_predictor(sd):
    import warnings
    warnings.simplefilter("ignore")
    from models import ProsodyPredictor
    p = ProsodyPredictor(style_dim=128, d_hid=512, nlayers=3, max_dur=50, dropout=0.2)
    ref = {k: v for k, v in p.state_dict().items() if k.startswith(F0N_PREFIXES)}
    assert set(ref) == set(sd), set(ref) ^ set(sd)
    for k in ref:
        assert tuple(ref[k].shape) == tuple(sd[k].shape), (k, ref[k].shape, sd[k].shape)
    missing, unexpected = p.load_state_dict(sd, strict=False)
    assert not unexpected and all(not k.startswith(F0N_PREFIXES) for k in missing)
    return p.eval()


def build_full_predictor(sd):
    """ProsodyPredictor with both halves loaded (strict: every parameter of the reference module is in `sd`)."""
    import warnings
    warnings.simplefilter("ignore")
    from models import ProsodyPredictor
    p = ProsodyPredictor(style_dim=128, d_hid=512, nlayers=3, max_dur=50, dropout=0.2)
    ref = p.state_dict()
    assert set(ref) == set(sd), set(ref) ^ set(sd)
    for k in ref:
        assert tuple(ref[k].shape) == tuple(sd[k].shape), (k, ref[k].shape, sd[k].shape)
    p.load_state_dict(sd)
    return p.eval()


def run_duration(B, L, wseed, iseed, tap_names=()):
    """inference.py:236-245 after the TextEncoder: mask -> predictor.text_encoder -> predictor.lstm -> duration_proj ->
    sigmoid-sum, for an equal-length batch."""
    sd = synth.make_predictor_state_dict(PredictorConfig(), seed=wseed, perturb=True, duration=True)
    p = build_full_predictor(sd)
    inp = synth.make_duration_inputs(B, L, seed=iseed)
    lengths = torch.full((B,), L, dtype=torch.long)
    mask = p.length_to_mask(lengths)                                    # inference.py:237 (all False here)
    taps, hooks = {}, []
    mods = dict(p.named_modules())
    def grab(o):
        o = o[0] if isinstance(o, tuple) else o
        if isinstance(o, torch.nn.utils.rnn.PackedSequence):            # the inner LSTMs run on packed sequences (models.py:503)
            o = torch.nn.utils.rnn.pad_packed_sequence(o, batch_first=True)[0]
        return o.detach().clone()

    for n in tap_names:
        hooks.append(mods[n].register_forward_hook(lambda mod, i, o, n=n: taps.__setitem__(n, grab(o))))
    with torch.no_grad():
        d = p.text_encoder(inp["t_en"], inp["s"], lengths, mask)        # inference.py:242
        x, _ = p.lstm(d)                                                # inference.py:243
        duration = torch.sigmoid(p.duration_proj(x)).sum(axis=-1)       # inference.py:244-245
    for h in hooks:
        h.remove()
    return d, duration, taps, sd


def run_f0n(B, T, wseed, iseed, tap_names=()):
    sd = synth.make_predictor_state_dict(PredictorConfig(), seed=wseed, perturb=True)
    p = build_predictor(sd)
    inp = synth.make_predictor_inputs(B, T, seed=iseed)
    taps, hooks = {}, []
    mods = dict(p.named_modules())
    for n in tap_names:
        hooks.append(mods[n].register_forward_hook(
            lambda mod, i, o, n=n: taps.__setitem__(n, (o[0] if isinstance(o, tuple) else o).detach().clone())))
    with torch.no_grad():
        f0, n = p.F0Ntrain(inp["en"], inp["s"])
    for h in hooks:
        h.remove()
    return f0, n, taps, p


def main():
    torch.set_num_threads(8)
    f0, n, taps, _ = run_f0n(2, 6, 0, 2001, ["shared", "F0.0", "F0.1", "F0.2", "N.1"])
    d = {"F0": f0.numpy(), "N": n.numpy()}
    for k, v in taps.items():
        d["tap:" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "f0n_B2_T6_w0_i2001.npz"), **d)
    print("f0n small", f0.shape, float(f0.abs().max()), float(n.abs().max()))

    f0, n, _, _ = run_f0n(1, 120, 0, 2002)
    np.savez_compressed(os.path.join(HERE, "f0n_B1_T120_w0_i2002.npz"), F0=f0.numpy(), N=n.numpy())
    print("f0n 3 s", f0.shape, float(f0.abs().max()))

    # ---- duration half
    d, dur, taps, full_sd = run_duration(2, 7, 0, 4001, ["text_encoder.lstms.0", "text_encoder.lstms.1", "lstm"])
    dd = {"d": d.numpy(), "duration": dur.numpy()}
    for k, v in taps.items():
        dd["tap:" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "dur_B2_L7_w0_i4001.npz"), **dd)
    print("duration small", d.shape, dur.numpy().round(3))
    d, dur, _, _ = run_duration(1, 64, 0, 4002)
    np.savez_compressed(os.path.join(HERE, "dur_B1_L64_w0_i4002.npz"), d=d.numpy(), duration=dur.numpy())
    print("duration L=64", float(dur.mean()))

    # ---- TextEncoder
    from models import TextEncoder
    from styletts2_lite_b200.config import TextEncoderConfig
    tsd = synth.make_text_state_dict(TextEncoderConfig(), seed=0, perturb=True)
    te = TextEncoder(channels=512, kernel_size=5, depth=3, n_symbols=178)                 # models.py:563
    ref = te.state_dict()
    assert set(ref) == set(tsd), set(ref) ^ set(tsd)
    for k in ref:
        assert tuple(ref[k].shape) == tuple(tsd[k].shape), (k, ref[k].shape, tsd[k].shape)
    te.load_state_dict(tsd)
    te = te.eval()
    for (B, L, seed, name) in ((2, 9, 5001, "text_B2_L9_w0_i5001.npz"), (1, 64, 5002, "text_B1_L64_w0_i5002.npz")):
        tok = synth.make_tokens(B, L, seed=seed)
        lengths = torch.full((B,), L, dtype=torch.long)
        taps_t = {}
        hk = te.cnn[0].register_forward_hook(lambda mod, i, o: taps_t.__setitem__("cnn.0", o.detach().clone()))
        with torch.no_grad():
            out = te(tok, lengths, te.length_to_mask(lengths))                            # inference.py:239
        hk.remove()
        np.savez_compressed(os.path.join(HERE, name), out=out.numpy(), **{"tap:cnn.0": taps_t["cnn.0"].numpy()})
        print("text encoder", out.shape, float(out.abs().max()))

    # ---- chained slice of inference.py:257-270 (cfg 3 after the text modules)
    from make_golden import build_reference, NoiseTape
    B, L, T = 2, 9, 16
    ci = synth.make_chain_inputs(B, L, T, seed=3003)
    dur, d_tok, t_en, s, noise = ci["dur"], ci["d"], ci["t_en"], ci["s"], ci["noise"]   # dur stands in for pred_dur (inference.py:257)
    psd = synth.make_predictor_state_dict(PredictorConfig(), seed=0, perturb=True)
    pred = build_predictor(psd)
    cfg = DecoderConfig.hifigan()
    dec = build_reference(cfg, synth.make_state_dict(cfg, seed=0, perturb=True))
    en_l, asr_l, f0_l, n_l, out_l = [], [], [], [], []
    with torch.no_grad():
        for b in range(B):                                             # the reference handles one sentence at a time
            aln = torch.zeros(L, T)                                    # inference.py:259-263
            c = 0
            for i in range(L):
                aln[i, c:c + int(dur[b, i])] = 1
                c += int(dur[b, i])
            en = d_tok[b:b + 1].transpose(-1, -2) @ aln.unsqueeze(0)   # inference.py:266
            f0p, np_ = pred.F0Ntrain(en, s[b:b + 1])                   # inference.py:267
            asr = t_en[b:b + 1] @ aln.unsqueeze(0)                     # inference.py:269
            with NoiseTape(noise[b:b + 1]):
                out = dec(asr, f0p, np_, s[b:b + 1])                   # inference.py:270
            en_l.append(en); asr_l.append(asr); f0_l.append(f0p); n_l.append(np_); out_l.append(out)
    np.savez_compressed(os.path.join(HERE, "chain_B2_L9_T16_w0.npz"), dur=dur.numpy().astype(np.int32),
                        en=torch.cat(en_l).numpy(), asr=torch.cat(asr_l).numpy(), F0=torch.cat(f0_l).numpy(),
                        N=torch.cat(n_l).numpy(), out=torch.cat(out_l).numpy())
    print("chain", torch.cat(out_l).shape, float(torch.cat(out_l).abs().max()))

    # schema of the F0Ntrain subset (drop-in contract of the predictor module)
    import json
    path = os.path.join(HERE, "state_dict_schema.json")
    schema = json.load(open(path))
    schema["predictor_f0n"] = {"num_params": sum(v.numel() for v in psd.values()),
                               "state_dict": {k: list(v.shape) for k, v in psd.items()}}
    schema["text_encoder"] = {"num_params": sum(v.numel() for v in tsd.values()),
                              "state_dict": {k: list(v.shape) for k, v in tsd.items()}}
    schema["predictor"] = {"num_params": sum(v.numel() for v in full_sd.values()),
                           "state_dict": {k: list(v.shape) for k, v in full_sd.items()}}
    json.dump(schema, open(path, "w"), sort_keys=True)


if __name__ == "__main__":
    main()
