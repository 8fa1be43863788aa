"""CPU tests: the numpy oracle against the golden fixtures produced by the unmodified
reference (tests/golden/make_golden.py).  Tolerances:
  * SineGen phase, length regulator: bit-exact / torch.equal-style equality;
  * waveforms and intermediate activations: max-abs <= 5e-5 (the reference's own
    fp32 self-consistency level, SURVEY.md §0: 2.2e-5 between its fp32 and fp64 runs)."""
import numpy as np
import pytest

from styletts2_lite_b200.config import DecoderConfig
from styletts2_lite_b200 import synth
from oracle import decoder_np as O
from helpers import golden, np_inputs, np_state_dict, sha, rel_l2

WAVE_TOL = 5e-5


def test_phase_bit_exact_small():
    cfg = DecoderConfig.hifigan()
    g = golden("hifigan_B2_T5_w0_i1001.npz")
    inp = np_inputs(2, 5, 1001, cfg)
    ph = O.sinegen_phase(inp["F0_curve"], cfg.upsample_scale)
    assert ph.dtype == np.float32
    assert np.array_equal(ph, g["phase"])


def test_phase_bit_exact_10s_checksum():
    cfg = DecoderConfig.hifigan()
    g = golden("sinegen_phase_B2_T400_i1004.npz")
    inp = np_inputs(2, 400, 1004, cfg, with_noise=False)
    ph = O.sinegen_phase(inp["F0_curve"], cfg.upsample_scale)
    assert np.array_equal(ph[:, :1200], g["phase_head"])
    assert np.array_equal(ph[:, -1200:], g["phase_tail"])
    assert sha(ph) == str(g["phase_sha256"])
    assert float(np.abs(ph).max()) > 5e4          # the regime where 1 ulp = 0.004..0.008 rad


def test_hifigan_small_waveform_and_taps():
    cfg = DecoderConfig.hifigan()
    g = golden("hifigan_B2_T5_w0_i1001.npz")
    sd = np_state_dict(cfg, 0, True)
    inp = np_inputs(2, 5, 1001, cfg)
    taps = {}
    out = O.decoder_forward(sd, cfg, inp["asr"], inp["F0_curve"], inp["N"], inp["s"], inp["noise"], taps=taps)
    assert out.shape == (2, 1, 3000)
    assert np.abs(out - g["out"]).max() <= WAVE_TOL
    pairs = {"encode": "encode", "decode.0": "decode.0", "decode.3": "decode.3",
             "generator.noise_res.0": "generator.noise_res.0.iter2",
             "generator.resblocks.0": "generator.resblocks.0.iter2",
             "generator.resblocks.5": "generator.resblocks.5.iter2",
             "generator.noise_res.3": "generator.noise_res.3.iter2",
             "generator.resblocks.11": "generator.resblocks.11.iter2"}
    for gname, oname in pairs.items():
        assert rel_l2(g["tap:" + gname], taps[oname][:1]) <= 2e-5, gname
    har_ref = g["tap:generator.m_source"].transpose(0, 2, 1)
    assert np.abs(taps["har_source"][:1] - har_ref).max() <= 1e-6


def test_hifigan_reference_init():
    cfg = DecoderConfig.hifigan()
    g = golden("hifigan_B1_T4_w3_i1002_plain.npz")
    sd = np_state_dict(cfg, 3, False)
    inp = np_inputs(1, 4, 1002, cfg)
    out = O.decoder_forward(sd, cfg, inp["asr"], inp["F0_curve"], inp["N"], inp["s"], inp["noise"])
    assert np.abs(out - g["out"]).max() <= WAVE_TOL


def test_hifigan_cfg1_3s():
    """BASELINE.json configs[0]: B=1, ~3 s (T=120)."""
    cfg = DecoderConfig.hifigan()
    g = golden("hifigan_B1_T120_w0_i1001.npz")
    sd = np_state_dict(cfg, 0, True)
    inp = np_inputs(1, 120, 1001, cfg)
    ph = O.sinegen_phase(inp["F0_curve"], cfg.upsample_scale)
    assert sha(ph) == str(g["phase_sha256"])
    taps = {}
    out = O.decoder_forward(sd, cfg, inp["asr"], inp["F0_curve"], inp["N"], inp["s"], inp["noise"], taps=taps)
    assert np.abs(taps["har_source"] - g["har_source"].transpose(0, 2, 1)).max() <= 1e-6
    assert np.abs(out - g["out"]).max() <= 1e-4


def test_istftnet_small():
    cfg = DecoderConfig.istftnet()
    g = golden("istftnet_B2_T5_w0_i1005.npz")
    sd = np_state_dict(cfg, 0, True)
    inp = np_inputs(2, 5, 1005, cfg)
    taps = {}
    out = O.decoder_forward(sd, cfg, inp["asr"], inp["F0_curve"], inp["N"], inp["s"], inp["noise"], taps=taps)
    assert out.shape == (2, 1, 3000)
    assert rel_l2(g["tap:generator.noise_res.0"], taps["generator.noise_res.0.iter2"][:1]) <= 2e-5
    assert rel_l2(g["tap:generator.noise_res.1"], taps["generator.noise_res.1.iter2"][:1]) <= 2e-5
    assert np.abs(out - g["out"]).max() <= WAVE_TOL


def test_length_regulator_golden():
    g = golden("length_regulator_L37.npz")
    pd = O.round_durations(g["duration"][0])
    assert np.array_equal(pd, g["pred_dur"].astype(np.int64))
    assert pd[3] == 1 and pd[5] == 2 and pd[6] == 4          # clamp, half-to-even
    out = O.length_regulate(g["t_en"][0], pd)
    assert np.array_equal(out, g["asr"][0])
    # gather restatement == matmul restatement
    idx = np.repeat(np.arange(len(pd)), pd)
    assert np.array_equal(out, g["t_en"][0][:, idx])


def test_length_regulator_batch_ragged():
    rng = np.random.default_rng(0)
    B, C, L = 3, 8, 11
    src = rng.standard_normal((B, C, L)).astype(np.float32)
    dur = rng.integers(1, 5, (B, L))
    dur[1, 7:] = 0                                           # ragged: padded tokens
    dur[2, :] = 0                                            # empty utterance
    out = O.length_regulate_batch(src, dur)
    assert out.shape == (B, C, int(dur.sum(1).max()))
    for b in range(B):
        idx = np.repeat(np.arange(L), dur[b])
        assert np.array_equal(out[b, :, :len(idx)], src[b][:, idx])
        assert not out[b, :, len(idx):].any()


def test_bf16_rounding_helper():
    x = np.array([1.0, 1.00390625, 1.005859375, -3.1415927, 0.0], np.float32)
    r = O.round_bf16(x)
    import torch
    assert np.array_equal(r, torch.from_numpy(x).bfloat16().float().numpy())


def test_torch_cpu_port_matches_golden():
    """oracle/decoder_torch.py (the CPU speed baseline of bench.py) is pinned to the same reference outputs."""
    import torch
    from oracle import decoder_torch as OT
    from styletts2_lite_b200 import synth
    for cfg, name, B, T, seed in ((DecoderConfig.hifigan(), "hifigan_B2_T5_w0_i1001.npz", 2, 5, 1001),
                                  (DecoderConfig.istftnet(), "istftnet_B2_T5_w0_i1005.npz", 2, 5, 1005)):
        g = golden(name)
        W = OT.TorchWeights(synth.make_state_dict(cfg, 0, True))
        inp = synth.make_inputs(B, T, seed, cfg)
        out = OT.decoder_forward(W, cfg, inp["asr"], inp["F0_curve"], inp["N"], inp["s"], inp["noise"]).numpy()
        assert np.abs(out - g["out"]).max() <= WAVE_TOL, name


# ---------------------------------------------------------------- §8(f) N1: F0Ntrain
def test_predictor_oracle_matches_reference_fixtures():
    from oracle import predictor_np as P
    sd = {k: v.numpy() for k, v in synth.make_predictor_state_dict(seed=0).items()}
    g = golden("f0n_B2_T6_w0_i2001.npz")
    inp = synth.make_predictor_inputs(2, 6, seed=2001)
    taps = {}
    f0, n = P.f0n_train(sd, inp["en"].numpy(), inp["s"].numpy(), taps=taps)
    assert np.abs(f0 - g["F0"]).max() <= 1e-5 and np.abs(n - g["N"]).max() <= 1e-5
    assert np.abs(taps["shared"] - g["tap:shared"]).max() <= 5e-6
    for k in ("F0.0", "F0.1", "F0.2", "N.1"):
        assert np.abs(taps[k] - g["tap:" + k]).max() <= 2e-5, k
    g = golden("f0n_B1_T120_w0_i2002.npz")
    inp = synth.make_predictor_inputs(1, 120, seed=2002)
    f0, n = P.f0n_train(sd, inp["en"].numpy(), inp["s"].numpy())
    assert np.abs(f0 - g["F0"]).max() <= 1e-5 and np.abs(n - g["N"]).max() <= 1e-5


def test_chain_fixture_is_consistent_with_oracle():
    """The chained cfg-3 fixture: oracle regulator reproduces en / asr bit-exactly, oracle F0Ntrain the F0 / N."""
    from oracle import predictor_np as P
    g = golden("chain_B2_L9_T16_w0.npz")
    ci = synth.make_chain_inputs(2, 9, 16, seed=3003)
    en = O.length_regulate_batch(ci["d"].numpy().transpose(0, 2, 1), g["dur"].astype(np.int64), 16)
    assert np.array_equal(en, g["en"])
    assert np.array_equal(O.length_regulate_batch(ci["t_en"].numpy(), g["dur"].astype(np.int64), 16), g["asr"])
    sd = {k: v.numpy() for k, v in synth.make_predictor_state_dict(seed=0).items()}
    f0, n = P.f0n_train(sd, en, ci["s"].numpy())
    assert np.abs(f0 - g["F0"]).max() <= 1e-5 and np.abs(n - g["N"]).max() <= 1e-5


def test_predictor_torch_port_matches_reference_fixture():
    import torch
    from oracle import predictor_torch as PT
    g = golden("f0n_B1_T120_w0_i2002.npz")
    inp = synth.make_predictor_inputs(1, 120, seed=2002)
    with torch.no_grad():
        f0, n = PT.f0n_train(synth.make_predictor_state_dict(seed=0), inp["en"], inp["s"])
    assert np.abs(f0.numpy() - g["F0"]).max() <= 1e-6 and np.abs(n.numpy() - g["N"]).max() <= 1e-6


def test_duration_oracle_matches_reference_fixtures():
    """§8(f) N2: numpy restatement and torch-CPU port of inference.py:242-245 against the reference fixtures."""
    import torch
    from oracle import predictor_np as P, predictor_torch as PT
    sdt = synth.make_predictor_state_dict(seed=0, duration=True)
    sd = {k: v.numpy() for k, v in sdt.items()}
    g = golden("dur_B2_L7_w0_i4001.npz")
    inp = synth.make_duration_inputs(2, 7, seed=4001)
    taps = {}
    d, dur = P.predict_duration(sd, inp["t_en"].numpy(), inp["s"].numpy(), taps=taps)
    assert np.abs(d - g["d"]).max() <= 2e-5 and np.abs(dur - g["duration"]).max() <= 1e-5
    assert np.abs(taps["text_encoder.lstms.0"] - g["tap:text_encoder.lstms.0"]).max() <= 5e-6
    assert np.abs(taps["lstm"] - g["tap:lstm"]).max() <= 1e-5
    with torch.no_grad():
        d2, dur2 = PT.predict_duration(sdt, inp["t_en"], inp["s"])
    assert np.abs(d2.numpy() - g["d"]).max() <= 2e-5 and np.abs(dur2.numpy() - g["duration"]).max() <= 1e-5
    g = golden("dur_B1_L64_w0_i4002.npz")
    inp = synth.make_duration_inputs(1, 64, seed=4002)
    d, dur = P.predict_duration(sd, inp["t_en"].numpy(), inp["s"].numpy())
    assert np.abs(d - g["d"]).max() <= 2e-5 and np.abs(dur - g["duration"]).max() <= 2e-5


def test_text_encoder_oracle_matches_reference_fixtures():
    """§8(f) N3: numpy restatement and torch-CPU port of TextEncoder.forward against the reference fixtures."""
    import torch
    from oracle import predictor_np as P, predictor_torch as PT
    sdt = synth.make_text_state_dict(seed=0)
    sd = {k: v.numpy() for k, v in sdt.items()}
    for B, L, seed, name in ((2, 9, 5001, "text_B2_L9_w0_i5001.npz"), (1, 64, 5002, "text_B1_L64_w0_i5002.npz")):
        g = golden(name)
        tok = synth.make_tokens(B, L, seed=seed)
        taps = {}
        out = P.text_encoder(sd, tok.numpy(), taps=taps)
        assert np.abs(out - g["out"]).max() <= 5e-6 and np.abs(taps["cnn.0"] - g["tap:cnn.0"]).max() <= 1e-5
        with torch.no_grad():
            assert np.abs(PT.text_encoder(sdt, tok).numpy() - g["out"]).max() <= 5e-6


def test_vocos_oracle_matches_reference_fixtures():
    """Vocos decoder variant (Modules/vocos.py:364-422; tests/golden/make_golden_vocos.py): ConvNeXt blocks, final LayerNorm,
    ISTFTHead with 'same' padding -- output and taps of the small case, output of the 3 s case."""
    cfg = DecoderConfig.vocos()
    sd = {k: v.numpy() for k, v in synth.make_state_dict(cfg, 0, True).items()}
    g = golden("vocos_B2_T6_w0_i1000.npz")
    inp = {k: v.numpy() for k, v in synth.make_inputs(2, 6, 1000, cfg, with_noise=False).items()}
    taps = {}
    out = O.decoder_forward(sd, cfg, inp["asr"], inp["F0_curve"], inp["N"], inp["s"], None, taps=taps)
    assert out.shape == (2, 1, 3600) and np.abs(out - g["out"]).max() <= 2e-6
    assert np.abs(taps["decode.3"] - g["tap:decode.3"]).max() <= 2e-5
    for i in (0, 7):
        assert np.abs(taps["generator.convnext.%d" % i] - g["tap:generator.convnext.%d" % i]).max() <= 2e-5
    assert np.abs(taps["generator.final_layer_norm"] - g["tap:generator.final_layer_norm"]).max() <= 2e-5
    assert np.abs(taps["generator.stft.out"].transpose(0, 2, 1) - g["tap:generator.stft.out"]).max() <= 2e-5
    g = golden("vocos_B1_T120_w0_i1001.npz")
    inp = {k: v.numpy() for k, v in synth.make_inputs(1, 120, 1001, cfg, with_noise=False).items()}
    out = O.decoder_forward(sd, cfg, inp["asr"], inp["F0_curve"], inp["N"], inp["s"], None)
    assert out.shape == (1, 1, 72000) and np.abs(out - g["out"]).max() <= 5e-6


def test_duration_smoothing_oracle_matches_reference_statements():
    """inference.py:248-257 lifted from the reference file and executed as is (tests/golden/make_golden_smooth.py): noise mix,
    z-score outlier replacement on duration[1:-2], / speed, rounding -- sentence lengths 2 .. 200 incl. the empty / single-element
    slices, a previous mean, t = 0."""
    g = golden("smooth_cases.npz")
    replaced = 0
    for k in range(int(g["n_cases"])):
        dur, z = g["duration_%d" % k][0], g["noise_%d" % k][0]
        out, mean = O.smooth_durations(dur, z, float(g["t_%d" % k]), float(g["speed_%d" % k]), float(g["prev_%d" % k]))
        ref = g["out_%d" % k][0]
        assert np.abs(out - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), k
        assert abs(mean - g["mean_%d" % k]) <= 1e-5
        safe = np.abs(ref - np.floor(ref) - 0.5) > 1e-4              # away from rounding ties
        assert np.array_equal(O.round_durations(out)[safe], g["pred_%d" % k].astype(np.int64)[safe]), k
        replaced += int((np.abs(ref - dur / g["speed_%d" % k]) > 5.0).sum())
    assert replaced >= 5                                             # the spikes of the fixture were really replaced


def test_ragged_text_modules_match_reference_padded_batches():
    """Padded batches (lengths 9, 6, 4) of TextEncoder.forward and of the duration half, as the reference's modules handle them
    (masked_fill_ + pack_padded_sequence, models.py:258-285, :485-520, :426-439): fixtures of tests/golden/make_golden_ragged.py,
    which also assert that the padded batch equals the one-sentence-at-a-time inference.py calls on the valid tokens."""
    from oracle import predictor_np as P
    g = golden("text_ragged_B3_L9_w0_i5101.npz")
    sd = {k: v.numpy() for k, v in synth.make_text_state_dict(seed=0).items()}
    lens = g["lengths"]
    taps = {}
    out = P.text_encoder(sd, g["tokens"], taps=taps, lengths=lens)
    assert np.abs(out - g["out"]).max() <= 5e-6 and np.abs(taps["cnn.0"] - g["tap:cnn.0"]).max() <= 1e-5
    for b, n in enumerate(lens):
        assert np.abs(out[b, :, :n] - g["single%d" % b][0]).max() <= 1e-5
        assert not out[b, :, n:].any()
        # a padded utterance equals the oracle on that utterance alone
        assert np.abs(P.text_encoder(sd, g["tokens"][b:b + 1, :n])[0] - out[b, :, :n]).max() <= 1e-6
    g = golden("dur_ragged_B3_L9_w0_i4101.npz")
    sd = {k: v.numpy() for k, v in synth.make_predictor_state_dict(seed=0, duration=True).items()}
    taps = {}
    d, dur = P.predict_duration(sd, g["t_en"], g["s"], taps=taps, lengths=g["lengths"])
    assert np.abs(d - g["d"]).max() <= 2e-5 and np.abs(dur - g["duration"]).max() <= 2e-5
    assert np.abs(taps["text_encoder.lstms.0"] - g["tap:text_encoder.lstms.0"]).max() <= 5e-6
    assert np.abs(taps["lstm"] - g["tap:lstm"]).max() <= 1e-5
    for b, n in enumerate(g["lengths"]):
        assert np.abs(dur[b, :n] - g["single_dur%d" % b][0]).max() <= 2e-5
        assert not d[b, n:].any()


# ---------------------------------------------------------------- round-2 fixtures (tests/golden/make_golden_r2.py)
def test_synth_plain_init_is_the_reference_init():
    """`init_weights` N(0, 0.01) (hifigan.py:37,47,318-319) never reaches the reference's forward: under the legacy weight_norm
    it rewrites the derived `.weight` only, which the pre-forward hook recomputes from weight_g / weight_v -- and those keep
    PyTorch's default U(+-1/sqrt(fan_in)) with g = ||v|| (reference_init_stats.json, measured on a freshly constructed
    reference Decoder).  synth.make_state_dict(perturb=False) draws exactly that distribution."""
    import json
    import math
    import os
    from helpers import GOLDEN
    st = json.load(open(os.path.join(GOLDEN, "reference_init_stats.json")))
    cfg = DecoderConfig.hifigan()
    sd = np_state_dict(cfg, 3, False)
    for k in ("generator.resblocks.0.convs1.0", "generator.resblocks.11.convs2.2", "generator.ups.0", "generator.conv_post",
              "generator.noise_res.0.convs1.0", "encode.conv1"):
        v, g, r = sd[k + ".weight_v"], sd[k + ".weight_g"], st[k + ".weight_v"]
        fan_in = v.shape[1] * v.shape[2]
        bound = 1.0 / math.sqrt(fan_in)
        # the reference's tensor is uniform on +-bound, not N(0, 0.01): abs-max just under the bound, std = bound / sqrt(3)
        assert r["absmax"] <= bound * (1 + 1e-6) and r["absmax"] >= 0.97 * bound, k
        assert abs(r["std"] - bound / math.sqrt(3)) <= (0.15 if v.size < 1000 else 0.02) * bound, k
        assert abs(v.std() - r["std"]) <= (0.15 if v.size < 1000 else 0.02) * bound and np.abs(v).max() <= bound * (1 + 1e-6), k
        # g = ||v|| over all dims but 0 in both
        assert np.abs(g.reshape(-1) - np.sqrt((v.reshape(v.shape[0], -1) ** 2).sum(1))).max() <= 1e-5, k
    eff = st["__effective__generator.resblocks.0.convs1.0.weight"]
    assert abs(eff["std"] - st["generator.resblocks.0.convs1.0.weight_v"]["std"]) <= 1e-6     # forward uses (g, v), not N(0, 0.01)


@pytest.mark.parametrize("name,variant,T,ws,iseed,perturb", [
    ("hifigan_B1_T120_w3_i1007_plain", "hifigan", 120, 3, 1007, False),
    ("hifigan_B1_T203_w0_i1009", "hifigan", 203, 0, 1009, True),
    ("istftnet_B1_T120_w0_i1006", "istftnet", 120, 0, 1006, True),
    ("istftnet_B1_T203_w0_i1008", "istftnet", 203, 0, 1008, True),
    ("hifigan_B1_T400_w0_i1003", "hifigan", 400, 0, 1003, True)])
def test_torch_cpu_port_matches_round2_goldens(name, variant, T, ws, iseed, perturb):
    """3 s / ragged / 10 s fixtures of both decoder variants, incl. the reference's true initialisation: the torch CPU port
    (what `bench.py --impl reference` times) against the unmodified reference; SineGen phase checksum of the numpy oracle."""
    from oracle import decoder_torch as OT
    from styletts2_lite_b200 import synth
    cfg = DecoderConfig.hifigan() if variant == "hifigan" else DecoderConfig.istftnet()
    g = golden(name + ".npz")
    W = OT.TorchWeights(synth.make_state_dict(cfg, ws, perturb))
    inp = synth.make_inputs(1, T, iseed, cfg)
    out = OT.decoder_forward(W, cfg, inp["asr"], inp["F0_curve"], inp["N"], inp["s"], inp["noise"]).numpy()
    assert out.shape == g["out"].shape
    assert np.abs(out - g["out"]).max() <= WAVE_TOL, name
    ph = O.sinegen_phase(inp["F0_curve"].numpy(), cfg.upsample_scale)
    assert sha(ph) == str(g["phase_sha256"]), name
