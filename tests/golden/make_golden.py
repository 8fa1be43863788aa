"""Generate the committed golden fixtures from the UNMODIFIED reference.

Runs only in the authoring container (needs /root/reference, which does not exist
on the GPU box):   python tests/golden/make_golden.py

For each case it builds the reference Decoder (Modules/hifigan.py:416 or
Modules/istftnet.py:660), loads the synthetic state_dict of
styletts2_lite_b200.synth through the reference's own load_state_dict, replays a
shared noise tape into the three random draws (torch.rand hifigan.py:126,
torch.randn_like :213 and :267) and stores inputs-by-seed + reference outputs.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from styletts2_lite_b200.config import DecoderConfig  # noqa: E402
from styletts2_lite_b200 import synth  # noqa: E402


def build_reference(cfg, sd):
    import warnings
    warnings.simplefilter("ignore")
    kw = dict(dim_in=cfg.dim_in, style_dim=cfg.style_dim, dim_out=80,
              resblock_kernel_sizes=cfg.resblock_kernel_sizes, upsample_rates=cfg.upsample_rates,
              upsample_initial_channel=cfg.upsample_initial_channel,
              resblock_dilation_sizes=cfg.resblock_dilation_sizes,
              upsample_kernel_sizes=cfg.upsample_kernel_sizes)
    if cfg.is_istft:
        from Modules.istftnet import Decoder
        kw.update(gen_istft_n_fft=cfg.gen_istft_n_fft, gen_istft_hop_size=cfg.gen_istft_hop_size)
    else:
        from Modules.hifigan import Decoder
    m = Decoder(**kw)
    ref_sd = m.state_dict()
    assert set(ref_sd.keys()) == set(sd.keys()), (set(ref_sd) ^ set(sd))
    for k in ref_sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), (k, ref_sd[k].shape, sd[k].shape)
    m.load_state_dict(sd)
    return m.eval()


class NoiseTape:
    """Replays `noise` into the first randn_like of each forward; the second draw
    (hifigan.py:267, discarded by Generator) gets zeros."""

    def __init__(self, noise):
        self.noise = noise
        self.captured = {}

    def __enter__(self):
        self._randn_like = torch.randn_like
        self._sin = torch.sin
        tape = self

        def randn_like(t, *a, **k):
            if t.dim() == 3 and t.shape[-1] == 9:
                assert tuple(t.shape) == tuple(tape.noise.shape)
                return tape.noise.clone()
            return torch.zeros_like(t)

        def sin(t, *a, **k):
            if t.dim() == 3 and t.shape[-1] == 9 and "phase" not in tape.captured:
                tape.captured["phase"] = t.detach().clone()
            return tape._sin(t, *a, **k)

        torch.randn_like = randn_like
        torch.sin = sin
        return self

    def __exit__(self, *a):
        torch.randn_like = self._randn_like
        torch.sin = self._sin


def run_case(cfg, B, T, wseed, iseed, perturb, tap_names=()):
    sd = synth.make_state_dict(cfg, seed=wseed, perturb=perturb)
    m = build_reference(cfg, sd)
    inp = synth.make_inputs(B, T, seed=iseed, cfg=cfg)
    taps = {}
    hooks = []
    mods = dict(m.named_modules())
    for n in tap_names:
        hooks.append(mods[n].register_forward_hook(
            lambda mod, i, o, n=n: taps.__setitem__(n, (o[0] if isinstance(o, tuple) else o).detach().clone())))
    with torch.no_grad(), NoiseTape(inp["noise"]) as tape:
        out = m(inp["asr"], inp["F0_curve"], inp["N"], inp["s"])
    for h in hooks:
        h.remove()
    return out, taps, tape.captured, inp


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def dump_schema(out_dir):
    """state_dict keys/shapes of the reference Decoders (the drop-in contract)."""
    import json
    schema = {}
    for cfg in (DecoderConfig.hifigan(), DecoderConfig.istftnet()):
        sd = synth.make_state_dict(cfg, seed=0, perturb=False)
        m = build_reference(cfg, sd)
        schema[cfg.type] = {"num_params": sum(p.numel() for p in m.parameters()),
                            "state_dict": {k: list(v.shape) for k, v in m.state_dict().items()}}
    with open(os.path.join(out_dir, "state_dict_schema.json"), "w") as f:
        json.dump(schema, f, sort_keys=True)


def main():
    torch.set_num_threads(8)
    out_dir = HERE
    dump_schema(out_dir)
    if "--schema-only" in sys.argv:
        return
    # ---- hifigan small, with intermediate taps (batch element 0 only for the big ones)
    cfg = DecoderConfig.hifigan()
    tap_names = ["encode", "decode.0", "decode.3", "generator.m_source", "generator.noise_res.0",
                 "generator.ups.0", "generator.resblocks.0", "generator.resblocks.5",
                 "generator.noise_res.3", "generator.resblocks.11"]
    out, taps, cap, inp = run_case(cfg, 2, 5, 0, 1001, True, tap_names)
    d = {"out": out.numpy(), "phase": cap["phase"].numpy()}
    for n, v in taps.items():
        d["tap:" + n] = v[:1].numpy()
    np.savez_compressed(os.path.join(out_dir, "hifigan_B2_T5_w0_i1001.npz"), **d)
    print("hifigan small", out.shape, float(out.abs().max()))

    # ---- hifigan, reference-like init (g=||v||, alpha=1)
    out, _, _, _ = run_case(cfg, 1, 4, 3, 1002, False)
    np.savez_compressed(os.path.join(out_dir, "hifigan_B1_T4_w3_i1002_plain.npz"), out=out.numpy())

    # ---- hifigan cfg 1 of BASELINE.json: B=1, T=120 (3 s)
    out, taps, cap, _ = run_case(cfg, 1, 120, 0, 1000 + 1, True, ["generator.m_source"])
    np.savez_compressed(os.path.join(out_dir, "hifigan_B1_T120_w0_i1001.npz"), out=out.numpy(),
                        phase_sha256=np.array(sha(cap["phase"].numpy())),
                        har_source=taps["generator.m_source"].numpy())
    print("hifigan cfg1", out.shape, float(out.abs().max()))

    # ---- SineGen phase at 10 s (T=400): checksum only (phase up to ~1e5 rad)
    inp = synth.make_inputs(2, 400, seed=1004, cfg=cfg, with_noise=False)
    from Modules.hifigan import SineGen
    sg = SineGen(24000, cfg.upsample_scale, harmonic_num=8, voiced_threshold=10)
    f0 = torch.repeat_interleave(inp["F0_curve"], cfg.upsample_scale, dim=1)[:, :, None]
    cap = {}
    _sin = torch.sin

    def sin(t, *a, **k):
        cap.setdefault("phase", t.detach().clone())
        return _sin(t, *a, **k)
    torch.sin = sin
    try:
        with torch.no_grad():
            sg(f0)
    finally:
        torch.sin = _sin
    ph = cap["phase"].numpy()
    np.savez_compressed(os.path.join(out_dir, "sinegen_phase_B2_T400_i1004.npz"),
                        phase_sha256=np.array(sha(ph)), phase_head=ph[:, :1200], phase_tail=ph[:, -1200:],
                        phase_absmax=np.array(np.abs(ph).max()))
    print("phase 10 s absmax", np.abs(ph).max())

    # ---- istftnet small
    cfg = DecoderConfig.istftnet()
    tap_names = ["decode.3", "generator.m_source", "generator.noise_res.0", "generator.noise_res.1",
                 "generator.resblocks.5", "generator.conv_post"]
    out, taps, cap, _ = run_case(cfg, 2, 5, 0, 1005, True, tap_names)
    d = {"out": out.numpy()}
    for n, v in taps.items():
        d["tap:" + n] = v[:1].numpy()
    np.savez_compressed(os.path.join(out_dir, "istftnet_B2_T5_w0_i1005.npz"), **d)
    print("istftnet small", out.shape, float(out.abs().max()))

    # ---- length regulator, restating inference.py:257-268 with torch exactly as written there
    g = torch.Generator().manual_seed(11)
    L, C = 37, 24
    duration = torch.rand(1, L, generator=g) * 6.0
    duration[0, 3] = 0.2          # clamps to 1
    duration[0, 5] = 2.5          # half-to-even -> 2
    duration[0, 6] = 3.5          # -> 4
    t_en = torch.randn(1, C, L, generator=g)
    pred_dur = torch.round(duration.squeeze()).clamp(min=1)
    pred_aln_trg = torch.zeros(L, int(pred_dur.sum().data))
    c_frame = 0
    for i in range(pred_aln_trg.size(0)):
        pred_aln_trg[i, c_frame:c_frame + int(pred_dur[i].data)] = 1
        c_frame += int(pred_dur[i].data)
    asr = t_en @ pred_aln_trg.unsqueeze(0)
    np.savez_compressed(os.path.join(out_dir, "length_regulator_L37.npz"), duration=duration.numpy(),
                        t_en=t_en.numpy(), pred_dur=pred_dur.numpy(), asr=asr.numpy())
    print("length regulator", asr.shape)


if __name__ == "__main__":
    main()
