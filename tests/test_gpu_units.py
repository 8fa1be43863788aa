"""-m gpu: per-kernel parity of the CUDA path against the numpy oracle / golden fixtures,
all through the C ABI (include/st2_b200.h).

Tolerances: length regulator and SineGen phase bit-exact; fp32 kernels: max-abs error
relative to the tensor's max-abs <= 2e-5 (fp32 accumulation-order noise)."""
import numpy as np
import pytest
import torch

from styletts2_lite_b200.config import DecoderConfig
from oracle import decoder_np as O
from helpers import golden, np_inputs, np_state_dict, sha

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import gpu_util as G
    from styletts2_lite_b200 import length_regulator as LR


def _relmax(ref, x):
    return float(np.abs(x - ref).max() / max(np.abs(ref).max(), 1e-30))


# ---------------------------------------------------------------- length regulator
def test_length_regulator_golden_bit_exact():
    g = golden("length_regulator_L37.npz")
    dur, tot = LR.round_durations(G.to_dev(g["duration"]))
    assert np.array_equal(dur.cpu().numpy()[0], g["pred_dur"].astype(np.int32))
    assert int(tot[0]) == g["asr"].shape[2]
    out = LR.length_regulate(G.to_dev(g["t_en"]), dur, int(tot[0]))
    assert torch.equal(out.cpu(), torch.from_numpy(g["asr"]))
    out_cl = LR.length_regulate(G.to_dev(g["t_en"]), dur, int(tot[0]), channels_last=True)
    assert torch.equal(out_cl.cpu().transpose(1, 2), torch.from_numpy(g["asr"]))


def test_duration_smoothing_reference_statements_and_ragged_batch():
    """st2_smooth_durations against the reference statements of inference.py:248-257 (fixture smooth_cases.npz, one sentence per
    case), then all cases as ONE padded batch with per-utterance n_tokens, previous means and tapes against the per-sentence
    oracle; finally the rounded durations."""
    g = golden("smooth_cases.npz")
    K = int(g["n_cases"])
    for k in range(K):
        t, speed, prev = float(g["t_%d" % k]), float(g["speed_%d" % k]), float(g["prev_%d" % k])
        out, mean = LR.smooth_durations(G.to_dev(g["duration_%d" % k]), G.to_dev(g["noise_%d" % k]), t, speed, prev)
        ref = g["out_%d" % k]
        assert np.abs(out.cpu().numpy() - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), k
        assert abs(float(mean[0]) - float(g["mean_%d" % k])) <= 1e-5
        pred, _ = LR.round_durations(out)
        safe = np.abs(ref[0] - np.floor(ref[0]) - 0.5) > 1e-4
        assert np.array_equal(pred.cpu().numpy()[0][safe], g["pred_%d" % k][safe]), k
    # the cases that share (t, speed) as one padded batch
    ks = [k for k in range(K) if float(g["t_%d" % k]) == np.float32(0.1) and float(g["speed_%d" % k]) == 1.0]
    L = max(g["duration_%d" % k].shape[1] for k in ks)
    dur = np.full((len(ks), L), 99.0, np.float32)                    # junk behind every utterance must not matter
    z = np.zeros((len(ks), L), np.float32)
    n_tok = np.array([g["duration_%d" % k].shape[1] for k in ks], np.int32)
    prev = np.array([float(g["prev_%d" % k]) for k in ks], np.float32)
    for i, k in enumerate(ks):
        dur[i, :n_tok[i]] = g["duration_%d" % k][0]
        z[i, :n_tok[i]] = g["noise_%d" % k][0]
    out, mean = LR.smooth_durations(G.to_dev(dur), G.to_dev(z), 0.1, 1.0, torch.from_numpy(prev), torch.from_numpy(n_tok))
    out = out.cpu().numpy()
    for i, k in enumerate(ks):
        ref, rmean = O.smooth_durations(g["duration_%d" % k][0], g["noise_%d" % k][0], 0.1, 1.0, float(prev[i]))
        assert np.abs(out[i, :n_tok[i]] - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), k
        assert not out[i, n_tok[i]:].any() and abs(float(mean[i]) - rmean) <= 1e-5
    with pytest.raises(Exception):
        LR.smooth_durations(G.to_dev(dur), None, 1.5)                # t outside [0, 1]


def test_length_regulator_ragged_empty_and_large():
    rng = np.random.default_rng(3)
    B, Cc, L = 5, 640, 300
    src = rng.standard_normal((B, Cc, L)).astype(np.float32)
    duration = (rng.random((B, L)) * 7).astype(np.float32)
    duration[0, :10] = [0.5, 1.5, 2.5, 3.5, 0.49, -3.0, 0.0, 2.4999, 2.5001, 4.5]   # ties, clamp
    n_tok = np.array([300, 1, 123, 0, 299], np.int32)                               # ragged + an empty utterance
    dur, tot = LR.round_durations(G.to_dev(duration), torch.from_numpy(n_tok))
    ref_dur = O.round_durations(duration)
    for b in range(B):
        ref_dur[b, n_tok[b]:] = 0
    assert np.array_equal(dur.cpu().numpy(), ref_dur.astype(np.int32))
    assert np.array_equal(tot.cpu().numpy(), ref_dur.sum(1).astype(np.int32))
    F = int(tot.max())
    out = LR.length_regulate(G.to_dev(src), dur, F).cpu().numpy()
    ref = O.length_regulate_batch(src, ref_dur, F)
    assert np.array_equal(out, ref)
    assert not out[3].any()
    # size-independent property: every output column equals some input column, in order
    b = 0
    idx = np.repeat(np.arange(L), ref_dur[b])
    assert np.array_equal(out[b, :, :len(idx)], src[b][:, idx])


def test_length_regulate_full_size_roundtrip():
    """cfg-3 shape (B=32, 8 s -> F=320, C=512+640): dur=1 everywhere is the identity;
    dur=k repeats every column k times."""
    B, Cc, L = 32, 512, 320
    src = torch.randn(B, Cc, L, device="cuda")
    ones = torch.ones(B, L, dtype=torch.int32, device="cuda")
    assert torch.equal(LR.length_regulate(src, ones, L), src)
    out = LR.length_regulate(src, ones * 3, 3 * L)
    assert torch.equal(out, src.repeat_interleave(3, dim=2))


# ---------------------------------------------------------------- SineGen
def test_sinegen_phase_bit_exact_small_and_golden():
    cfg = DecoderConfig.hifigan()
    g = golden("hifigan_B2_T5_w0_i1001.npz")
    inp = np_inputs(2, 5, 1001, cfg)
    ph, _ = G.sinegen_phase(inp["F0_curve"], cfg.upsample_scale)
    bad = int((ph != g["phase"]).sum())
    G.log("sinegen_phase_small", mismatches=bad, total=int(ph.size))
    assert bad == 0


@pytest.mark.parametrize("T,seed", [(400, 1004)])
def test_sinegen_phase_10s_golden_checksum(T, seed):
    cfg = DecoderConfig.hifigan()
    g = golden("sinegen_phase_B2_T400_i1004.npz")
    inp = np_inputs(2, T, seed, cfg, with_noise=False)
    ph, _ = G.sinegen_phase(inp["F0_curve"], cfg.upsample_scale)
    ref = O.sinegen_phase(inp["F0_curve"], cfg.upsample_scale)
    bad = int((ph != ref).sum())
    G.log("sinegen_phase_10s", mismatches=bad, total=int(ph.size), absmax=float(np.abs(ph).max()))
    assert bad == 0
    assert sha(ph) == str(g["phase_sha256"])


def test_sinegen_phase_60s_vs_oracle():
    cfg = DecoderConfig.hifigan()
    inp = np_inputs(1, 2400, 1006, cfg, with_noise=False)
    ph, _ = G.sinegen_phase(inp["F0_curve"], cfg.upsample_scale)
    ref = O.sinegen_phase(inp["F0_curve"], cfg.upsample_scale)
    bad = int((ph != ref).sum())
    G.log("sinegen_phase_60s", mismatches=bad, total=int(ph.size), absmax=float(np.abs(ph).max()))
    assert bad == 0


def test_sinegen_negative_and_zero_f0():
    cfg = DecoderConfig.hifigan()
    f0 = np.array([[0.0, -50.0, 120.5, 0.0, 9.99, 10.0, 10.01, 300.0]], np.float32)
    ph, _ = G.sinegen_phase(f0, cfg.upsample_scale)
    assert np.array_equal(ph, O.sinegen_phase(f0, cfg.upsample_scale))


def test_har_source_vs_oracle():
    cfg = DecoderConfig.hifigan()
    sd = np_state_dict(cfg, 0, True)
    W = O.Weights(sd)
    inp = np_inputs(2, 20, 1001, cfg)
    ref = O.source_module(W, inp["F0_curve"], cfg.upsample_scale, inp["noise"])[:, 0, :]
    got = G.har_source(inp["F0_curve"], inp["noise"], sd["generator.m_source.l_linear.weight"],
                       sd["generator.m_source.l_linear.bias"], cfg.upsample_scale)
    err = float(np.abs(got - ref).max())
    G.log("har_source", maxabs=err)
    assert err <= 1e-6


def test_har_source_device_noise_statistics():
    """noise=NULL: Philox noise drawn on the device; deterministic per seed, ~N(0,1)-driven."""
    cfg = DecoderConfig.hifigan()
    sd = np_state_dict(cfg, 0, True)
    f0 = np.zeros((2, 40), np.float32)                   # unvoiced: output = tanh(w . (0.1/3 * noise) + b)
    w = np.zeros(9, np.float32); w[0] = 1.0
    b = np.zeros(1, np.float32)
    a = G.har_source(f0, None, w, b, cfg.upsample_scale, seed=5)
    a2 = G.har_source(f0, None, w, b, cfg.upsample_scale, seed=5)
    c = G.har_source(f0, None, w, b, cfg.upsample_scale, seed=6)
    assert np.array_equal(a, a2) and not np.array_equal(a, c)
    z = np.arctanh(a.astype(np.float64)) / (0.1 / 3)
    G.log("philox_noise", mean=float(z.mean()), std=float(z.std()))
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02


# ---------------------------------------------------------------- AdaIN + activation
@pytest.mark.parametrize("Cc,T,act", [(32, 3000, "snake"), (256, 100, "snake"), (64, 1500, "lrelu"),
                                      (1024, 7, "lrelu"), (128, 601, "snake")])
def test_adain_act_vs_oracle(Cc, T, act):
    rng = np.random.default_rng(Cc + T)
    B = 3
    x = (rng.standard_normal((B, Cc, T)) * 2.0 + rng.standard_normal((B, Cc, 1)) * 3.0).astype(np.float32)
    s = rng.standard_normal((B, 128)).astype(np.float32)
    fw = (rng.standard_normal((2 * Cc, 128)) * 0.05).astype(np.float32)
    fb = (rng.standard_normal(2 * Cc) * 0.1).astype(np.float32)
    alpha = (0.6 + 0.8 * rng.random((1, Cc, 1))).astype(np.float32)
    ref = O.adain(x, s, fw, fb)
    ref = O.snake(ref, alpha) if act == "snake" else O.leaky_relu(ref, 0.2)
    h = (s @ fw.T + fb).astype(np.float32)
    got = G.cf(G.adain_act(G.cl(x), h, alpha if act == "snake" else None, act, slope=0.2, ld_pad=4))
    err = _relmax(ref, got)
    G.log("adain_act", C=Cc, T=T, act=act, relmax=err)
    assert err <= 2e-5


def test_adain_long_row_statistics():
    """1.44 M-element rows (60 s at C=32): mean/var must stay fp32-exact."""
    rng = np.random.default_rng(9)
    B, Cc, T = 1, 32, 1_440_000
    x = (rng.standard_normal((B, Cc, T)) * 0.5 + 100.0).astype(np.float32)      # large mean/std ratio
    h = np.zeros((B, 2 * Cc), np.float32)
    got = G.cf(G.adain_act(G.cl(x), h, None, "none"))
    ref = O.instance_norm(x)
    err = float(np.abs(got - ref).max())
    G.log("adain_long_row", maxabs=err)
    assert err <= 5e-5         # |x-mean|/std up to ~5; fp32 cancellation at mean/std = 200


@pytest.mark.parametrize("dt,tol", [("bf16", 4e-3), ("fp16", 6e-4)])
def test_adain_act_16bit_outputs(dt, tol):
    rng = np.random.default_rng(5)
    B, Cc, T = 2, 64, 777
    x = rng.standard_normal((B, Cc, T)).astype(np.float32)
    h = (rng.standard_normal((B, 2 * Cc)) * 0.3).astype(np.float32)
    alpha = (0.6 + 0.8 * rng.random((1, Cc, 1))).astype(np.float32)
    C_ = Cc
    gamma, beta = h[:, :C_, None], h[:, C_:, None]
    ref = O.snake(((1 + gamma) * O.instance_norm(x) + beta).astype(np.float32), alpha)
    got = G.cf(G.adain_act(G.cl(x), h, alpha, "snake", out_dtype=dt))
    err = _relmax(ref, got)
    G.log("adain_act_16", dtype=dt, relmax=err)
    assert err <= tol


# ---------------------------------------------------------------- convolutions (fp32 SIMT)
CONV_CASES = [
    # Cin, Cout, k, stride, pad, dil, T
    (32, 32, 11, 1, 25, 5, 700), (64, 64, 7, 1, 9, 3, 333), (128, 128, 3, 1, 1, 1, 257),
    (256, 256, 7, 1, 3, 1, 140), (514, 1024, 3, 1, 1, 1, 9), (1090, 512, 3, 1, 1, 1, 12),
    (1090, 1024, 1, 1, 0, 1, 7), (512, 64, 1, 1, 0, 1, 11), (1, 256, 60, 30, 15, 1, 3000),
    (1, 128, 12, 6, 3, 1, 1200), (1, 32, 1, 1, 0, 1, 500), (22, 256, 12, 6, 3, 1, 601),
    (22, 128, 1, 1, 0, 1, 301), (128, 22, 7, 1, 3, 1, 301), (32, 1, 7, 1, 3, 1, 400),
]


@pytest.mark.parametrize("Cin,Cout,k,stride,pad,dil,T", CONV_CASES)
def test_conv1d_fp32_vs_oracle(Cin, Cout, k, stride, pad, dil, T):
    rng = np.random.default_rng(Cin * 7 + Cout + k)
    B = 2
    x = rng.standard_normal((B, Cin, T)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin, k)) / np.sqrt(Cin * k)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    ref = O.conv1d(x, w, b, stride=stride, padding=pad, dilation=dil)
    got = G.cf(G.conv1d(G.cl(x), w, b, stride=stride, padding=pad, dilation=dil))
    assert got.shape == ref.shape
    err = _relmax(ref, got)
    G.log("conv1d_fp32", Cin=Cin, Cout=Cout, k=k, stride=stride, dil=dil, relmax=err)
    assert err <= 2e-5


CONVT_CASES = [
    # Cin, Cout, k, stride, pad, out_pad, T
    (512, 256, 20, 10, 5, 0, 24), (256, 128, 10, 5, 3, 1, 60), (128, 64, 6, 3, 2, 1, 130),
    (64, 32, 4, 2, 1, 0, 257), (256, 128, 12, 6, 3, 0, 100),
]


@pytest.mark.parametrize("Cin,Cout,k,stride,pad,opad,T", CONVT_CASES)
def test_conv_transpose1d_fp32_vs_oracle(Cin, Cout, k, stride, pad, opad, T):
    rng = np.random.default_rng(Cin + Cout + k)
    B = 2
    x = rng.standard_normal((B, Cin, T)).astype(np.float32)
    w = (rng.standard_normal((Cin, Cout, k)) / np.sqrt(Cin * k / stride)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    ref = O.conv_transpose1d(x, w, b, stride=stride, padding=pad, output_padding=opad)
    got = G.cf(G.conv1d(G.cl(x), w, b, stride=stride, padding=pad, output_padding=opad, transposed=True))
    assert got.shape == ref.shape
    assert not np.isnan(got).any()
    err = _relmax(ref, got)
    G.log("convT_fp32", Cin=Cin, Cout=Cout, k=k, stride=stride, relmax=err)
    assert err <= 2e-5


# ---------------------------------------------------------------- convolutions (tcgen05 tensor-core path)
# Same rounded operands as the oracle's emulation (RN to bf16/fp16), fp32 accumulation in TMEM:
# only the summation order differs -> same 2e-5 bound as the fp32 kernel.
TC_CONV_CASES = [
    (32, 32, 11, 1, 25, 5, 700), (32, 32, 3, 1, 1, 1, 131), (64, 64, 7, 1, 9, 3, 333), (64, 64, 11, 1, 5, 1, 1000),
    (128, 128, 3, 1, 1, 1, 257), (128, 128, 11, 1, 15, 3, 640), (256, 256, 7, 1, 3, 1, 140),
    (256, 256, 3, 1, 5, 5, 129), (1088, 512, 3, 1, 1, 1, 12), (1024, 1024, 3, 1, 1, 1, 130),
    (1088, 1024, 1, 1, 0, 1, 7), (512, 64, 1, 1, 0, 1, 11), (128, 22, 7, 1, 3, 1, 301),
]


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("Cin,Cout,k,stride,pad,dil,T", TC_CONV_CASES)
def test_conv1d_tensor_core_vs_oracle(Cin, Cout, k, stride, pad, dil, T, prec):
    rng = np.random.default_rng(Cin * 7 + Cout + k)
    B = 2
    x = rng.standard_normal((B, Cin, T)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin, k)) / np.sqrt(Cin * k)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    ref = O.conv1d(x, w, b, stride=stride, padding=pad, dilation=dil, operand=prec)
    got = G.cf(G.conv1d(G.cl(x), w, b, stride=stride, padding=pad, dilation=dil, precision=prec))
    assert got.shape == ref.shape
    nan = int(np.isnan(got).sum())
    err = _relmax(ref, np.nan_to_num(got))
    G.log("conv1d_tc", prec=prec, Cin=Cin, Cout=Cout, k=k, dil=dil, T=T, relmax=err, nan=nan)
    assert nan == 0
    assert err <= 2e-5


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("Cin,Cout,k,stride,pad,opad,T", CONVT_CASES)
def test_conv_transpose1d_tensor_core_vs_oracle(Cin, Cout, k, stride, pad, opad, T, prec):
    rng = np.random.default_rng(Cin + Cout + k)
    B = 2
    x = rng.standard_normal((B, Cin, T)).astype(np.float32)
    w = (rng.standard_normal((Cin, Cout, k)) / np.sqrt(Cin * k / stride)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32)
    ref = O.conv_transpose1d(x, w, b, stride=stride, padding=pad, output_padding=opad, operand=prec)
    got = G.cf(G.conv1d(G.cl(x), w, b, stride=stride, padding=pad, output_padding=opad, transposed=True, precision=prec))
    assert got.shape == ref.shape
    nan = int(np.isnan(got).sum())
    err = _relmax(ref, np.nan_to_num(got))
    G.log("convT_tc", prec=prec, Cin=Cin, Cout=Cout, k=k, stride=stride, relmax=err, nan=nan)
    assert nan == 0
    assert err <= 2e-5
