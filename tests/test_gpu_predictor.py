"""-m gpu: SURVEY.md §8(f) N1 -- `ProsodyPredictor.F0Ntrain` through the drop-in module / C ABI, and the chained
cfg-3 slice inference.py:257-270 (length regulator -> F0Ntrain -> Decoder) against fixtures of the unmodified reference.

fp32 path: max-abs <= 1e-4 against the reference's F0 / N (same bar as the decoder's fp32 path); 16-bit path:
per-layer relative L2 <= 1e-2 and output SNR >= 40 dB."""
import numpy as np
import pytest
import torch

from styletts2_lite_b200.config import DecoderConfig
from styletts2_lite_b200 import synth
from oracle import predictor_np as P
from helpers import golden, rel_l2, snr_db

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import gpu_util as G
    from styletts2_lite_b200 import _lib, length_regulator as LR
    from styletts2_lite_b200.decoder import B200Decoder
    from styletts2_lite_b200.predictor import B200F0NPredictor

_CACHE = {}


def _predictor():
    if "p" not in _CACHE:
        m = B200F0NPredictor(style_dim=128, d_hid=512, nlayers=3, max_dur=50, dropout=0.2)
        m.load_state_dict(synth.make_predictor_state_dict(seed=0))
        _CACHE["p"] = m.to("cuda").eval()
    return _CACHE["p"]


def _np_sd():
    return {k: v.numpy() for k, v in synth.make_predictor_state_dict(seed=0).items()}


def _run(m, inp, precision="fp32"):
    with torch.no_grad():
        f0, n = m.F0Ntrain(inp["en"].cuda(), inp["s"].cuda(), precision=precision)
    torch.cuda.synchronize()
    return f0.cpu().numpy(), n.cpu().numpy()


def test_f0n_small_fp32_golden_and_taps():
    g = golden("f0n_B2_T6_w0_i2001.npz")
    m = _predictor()
    inp = synth.make_predictor_inputs(2, 6, seed=2001)
    shapes = {"shared": (6, 512), "F0.0": (6, 512), "F0.1": (12, 256), "F0.2": (12, 256), "N.1": (12, 256)}
    bufs = {k: m.set_tap(k, 2, r, c) for k, (r, c) in shapes.items()}
    f0, n = _run(m, inp)
    m.clear_taps()
    assert np.abs(f0 - g["F0"]).max() <= 1e-4 and np.abs(n - g["N"]).max() <= 1e-4
    ref_shared = g["tap:shared"]                                   # [B, T, 512] (batch_first LSTM output)
    assert np.abs(bufs["shared"].cpu().numpy() - ref_shared).max() <= 2e-5
    for k in ("F0.0", "F0.1", "F0.2", "N.1"):
        assert rel_l2(g["tap:" + k], G.cf(bufs[k].cpu().numpy())) <= 2e-5, k


def test_f0n_3s_fp32_golden():
    g = golden("f0n_B1_T120_w0_i2002.npz")
    f0, n = _run(_predictor(), synth.make_predictor_inputs(1, 120, seed=2002))
    assert np.abs(f0 - g["F0"]).max() <= 1e-4 and np.abs(n - g["N"]).max() <= 1e-4


@pytest.mark.parametrize("B,T", [(11, 37), (1, 1), (17, 2)])
def test_f0n_ragged_batches_vs_oracle(B, T):
    """Batch sizes that do not fill the 8-utterance clusters of the LSTM kernel, odd lengths, a single frame."""
    inp = synth.make_predictor_inputs(B, T, seed=2100 + B)
    f0, n = _run(_predictor(), inp)
    rf0, rn = P.f0n_train(_np_sd(), inp["en"].numpy(), inp["s"].numpy())
    # InstanceNorm over 2-4 time steps is ill-conditioned (two nearly equal samples -> rstd up to 1/sqrt(eps) = 316
    # amplifies rounding differences of the producing conv): looser bound for the degenerate lengths only
    tol = 1e-4 if T >= 8 else 1e-3
    assert np.abs(f0 - rf0).max() <= tol and np.abs(n - rn).max() <= tol


def test_f0n_batch_independence():
    inp = synth.make_predictor_inputs(9, 50, seed=2200)
    m = _predictor()
    f0, n = _run(m, inp)
    one = {"en": inp["en"][4:5], "s": inp["s"][4:5]}
    f1, n1 = _run(m, one)
    assert np.array_equal(f0[4:5], f1) and np.array_equal(n[4:5], n1)


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
def test_f0n_tensor_core_path(prec):
    """cfg-3 batch shape (32 x 8 s): 16-bit operands on the convolutions against this library's fp32 path (itself
    within 1e-4 of the reference) -- SNR >= 40 dB on both outputs, and per-layer taps within 1e-2."""
    B, T = 32, 320
    inp = synth.make_predictor_inputs(B, T, seed=2300)
    m = _predictor()
    shapes = {"shared": (T, 512), "F0.0": (T, 512), "F0.1": (2 * T, 256), "F0.2": (2 * T, 256), "N.2": (2 * T, 256)}
    bufs = {k: m.set_tap(k, B, r, c) for k, (r, c) in shapes.items()}
    f0, n = _run(m, inp)
    ref = {k: v.cpu().numpy().copy() for k, v in bufs.items()}
    f0h, nh = _run(m, inp, precision=prec)
    m.clear_taps()
    for k in shapes:
        assert rel_l2(ref[k], bufs[k].cpu().numpy()) <= 1e-2, k
    assert snr_db(f0, f0h) >= 40.0 and snr_db(n, nh) >= 40.0


def test_f0n_error_behaviour():
    m = _predictor()
    with pytest.raises(_lib.St2Error):
        m.F0Ntrain(torch.zeros(1, 640, 4), torch.zeros(1, 128))            # CPU tensors: no CPU path
    with pytest.raises(ValueError):
        m.F0Ntrain(torch.zeros(1, 512, 4).cuda(), torch.zeros(1, 128).cuda())
    lib = _lib.load()
    import ctypes as C
    h = C.c_void_p()
    assert lib.st2_f0n_create(384, 128, C.byref(h)) == -1                  # ST2_ERR_INVALID, message set
    assert b"d_hid" in lib.st2_last_error()
    # a predictor handle is rejected by the decoder entry points
    assert lib.st2_decoder_workspace_bytes(m._handle, 1, 4, 0) < 0


def test_chain_regulator_predictor_decoder_matches_reference():
    """inference.py:257-270 on the GPU: durations -> (bit-exact) length regulation of d and t_en -> F0Ntrain ->
    Decoder, against the same chain run through the unmodified reference (tests/golden/make_golden_predictor.py)."""
    g = golden("chain_B2_L9_T16_w0.npz")
    B, L, T = 2, 9, 16
    ci = synth.make_chain_inputs(B, L, T, seed=3003)
    dur = ci["dur"].to(torch.int32).cuda()
    assert np.array_equal(dur.cpu().numpy(), g["dur"])
    en = LR.length_regulate(ci["d"].transpose(1, 2).contiguous().cuda(), dur, T)       # d^T @ alignment (inference.py:266)
    asr = LR.length_regulate(ci["t_en"].cuda(), dur, T)                                # t_en @ alignment (inference.py:269)
    assert torch.equal(en.cpu(), torch.from_numpy(g["en"])) and torch.equal(asr.cpu(), torch.from_numpy(g["asr"]))
    s = ci["s"].cuda()
    with torch.no_grad():
        f0, n = _predictor().F0Ntrain(en, s)
        assert np.abs(f0.cpu().numpy() - g["F0"]).max() <= 1e-4 and np.abs(n.cpu().numpy() - g["N"]).max() <= 1e-4
        cfg = DecoderConfig.hifigan()
        dec = B200Decoder(cfg, "fp32")
        dec.load_state_dict(synth.make_state_dict(cfg, 0, True))
        dec = dec.to("cuda").eval()
        out = dec(asr, f0, n, s, noise=ci["noise"].cuda())
    assert np.abs(out.cpu().numpy() - g["out"]).max() <= 1e-4


# ---------------------------------------------------------------- §8(f) N2: duration half
def _dur_predictor():
    if "d" not in _CACHE:
        m = B200F0NPredictor(style_dim=128, d_hid=512, nlayers=3, max_dur=50, dropout=0.2, duration=True)
        m.load_state_dict(synth.make_predictor_state_dict(seed=0, duration=True))
        _CACHE["d"] = m.to("cuda").eval()
    return _CACHE["d"]


def test_duration_small_fp32_golden_and_taps():
    g = golden("dur_B2_L7_w0_i4001.npz")
    m = _dur_predictor()
    inp = synth.make_duration_inputs(2, 7, seed=4001)
    bufs = {k: m.set_tap(k, 2, 7, 512) for k in ("text_encoder.lstms.0", "text_encoder.lstms.1", "lstm")}
    with torch.no_grad():
        d, dur = m.predict_duration(inp["t_en"].cuda(), inp["s"].cuda())
    m.clear_taps()
    assert np.abs(bufs["text_encoder.lstms.0"].cpu().numpy() - g["tap:text_encoder.lstms.0"]).max() <= 2e-5
    assert np.abs(bufs["text_encoder.lstms.1"].cpu().numpy() - g["tap:text_encoder.lstms.1"]).max() <= 5e-5   # AdaLayerNorm, [B,L,C]
    assert np.abs(bufs["lstm"].cpu().numpy() - g["tap:lstm"]).max() <= 2e-5
    assert np.abs(d.cpu().numpy() - g["d"]).max() <= 1e-4
    assert np.abs(dur.cpu().numpy() - g["duration"]).max() <= 1e-4
    # the same handle still serves F0Ntrain
    gf = golden("f0n_B2_T6_w0_i2001.npz")
    f0, n = _run(m, synth.make_predictor_inputs(2, 6, seed=2001))
    assert np.abs(f0 - gf["F0"]).max() <= 1e-4 and np.abs(n - gf["N"]).max() <= 1e-4


def test_duration_64_tokens_fp32_golden_and_rounding():
    g = golden("dur_B1_L64_w0_i4002.npz")
    m = _dur_predictor()
    inp = synth.make_duration_inputs(1, 64, seed=4002)
    with torch.no_grad():
        d, dur = m.predict_duration(inp["t_en"].cuda(), inp["s"].cuda())
    assert np.abs(d.cpu().numpy() - g["d"]).max() <= 1e-4 and np.abs(dur.cpu().numpy() - g["duration"]).max() <= 1e-4
    # inference.py:257: the integer durations agree wherever the reference value is not within 1e-3 of a rounding tie
    pred, _ = LR.round_durations(dur)
    ref = np.maximum(np.round(g["duration"]), 1).astype(np.int32)
    safe = np.abs(g["duration"] - np.floor(g["duration"]) - 0.5) > 1e-3
    assert np.array_equal(pred.cpu().numpy()[safe], ref[safe])


@pytest.mark.parametrize("prec", ["fp16"])
def test_duration_tensor_core_path_and_batch(prec):
    """cfg-3 batch shape (32 utterances x 64 tokens): ragged cluster fill (B=32 -> 4 clusters per direction), batch
    independence, and the fp16-operand input projections against the fp32 path."""
    B, L = 32, 64
    inp = synth.make_duration_inputs(B, L, seed=4100)
    m = _dur_predictor()
    with torch.no_grad():
        d, dur = m.predict_duration(inp["t_en"].cuda(), inp["s"].cuda())
        d1, dur1 = m.predict_duration(inp["t_en"][5:6].cuda(), inp["s"][5:6].cuda())
        dh, durh = m.predict_duration(inp["t_en"].cuda(), inp["s"].cuda(), precision=prec)
    assert torch.equal(d[5:6], d1) and torch.equal(dur[5:6], dur1)
    from oracle import predictor_np as PN
    sd = {k: v.numpy() for k, v in synth.make_predictor_state_dict(seed=0, duration=True).items()}
    rd, rdur = PN.predict_duration(sd, inp["t_en"][:3].numpy(), inp["s"][:3].numpy())
    assert np.abs(d[:3].cpu().numpy() - rd).max() <= 1e-4 and np.abs(dur[:3].cpu().numpy() - rdur).max() <= 1e-4
    assert rel_l2(d.cpu().numpy(), dh.cpu().numpy()) <= 1e-2
    assert snr_db(dur.cpu().numpy(), durh.cpu().numpy()) >= 40.0


def test_duration_needs_its_weights():
    m = _predictor()                                   # built without the duration half
    with pytest.raises(RuntimeError):
        m.predict_duration(torch.zeros(1, 512, 4).cuda(), torch.zeros(1, 128).cuda())
    lib = _lib.load()
    assert lib.st2_dur_workspace_bytes(m._handle, 1, 4, 0) == -2          # ST2_ERR_STATE


# ---------------------------------------------------------------- §8(f) N3: TextEncoder
def _text_encoder():
    if "t" not in _CACHE:
        from styletts2_lite_b200.text_encoder import B200TextEncoder
        m = B200TextEncoder(channels=512, kernel_size=5, depth=3, n_symbols=178)
        m.load_state_dict(synth.make_text_state_dict(seed=0))
        _CACHE["t"] = m.to("cuda").eval()
    return _CACHE["t"]


@pytest.mark.parametrize("B,L,seed,name", [(2, 9, 5001, "text_B2_L9_w0_i5001.npz"), (1, 64, 5002, "text_B1_L64_w0_i5002.npz")])
def test_text_encoder_fp32_golden(B, L, seed, name):
    g = golden(name)
    m = _text_encoder()
    tok = synth.make_tokens(B, L, seed=seed).cuda()
    lengths = torch.full((B,), L, dtype=torch.long)
    buf = m.set_tap("cnn.0", B, L, 512)
    with torch.no_grad():
        out = m(tok, lengths, torch.zeros(B, L, dtype=torch.bool))
    m.clear_taps()
    assert np.abs(G.cf(buf.cpu().numpy()) - g["tap:cnn.0"]).max() <= 2e-5
    assert out.shape == (B, 512, L) and np.abs(out.cpu().numpy() - g["out"]).max() <= 1e-4


def test_text_encoder_batch_tensor_core_and_errors():
    from oracle import predictor_np as PN
    B, L = 32, 64
    m = _text_encoder()
    tok = synth.make_tokens(B, L, seed=5100).cuda()
    with torch.no_grad():
        out = m(tok)
        one = m(tok[7:8])
        o16 = m(tok, precision="fp16")
    assert torch.equal(out[7:8], one)
    sd = {k: v.numpy() for k, v in synth.make_text_state_dict(seed=0).items()}
    ref = PN.text_encoder(sd, tok[:2].cpu().numpy())
    assert np.abs(out[:2].cpu().numpy() - ref).max() <= 1e-4
    assert rel_l2(out.cpu().numpy(), o16.cpu().numpy()) <= 1e-2
    with pytest.raises(ValueError):
        m(tok, torch.tensor([L] * (B - 1) + [L + 3]))                      # longer than the batch
    with pytest.raises(ValueError):
        m(tok, torch.tensor([L] * (B - 1) + [L - 3]), torch.zeros(B, L, dtype=torch.bool))   # mask disagrees with the lengths
    with pytest.raises(IndexError):
        m(torch.full((1, 4), 178, dtype=torch.long).cuda())                # id out of range, like nn.Embedding
    with pytest.raises(_lib.St2Error):
        m(tok.cpu())


def test_tokens_to_waveform_chain_runs_on_the_gpu():
    """inference.py:239-270 end to end in this library at a small size: TextEncoder -> duration half -> durations ->
    length regulation -> F0Ntrain -> Decoder.  Each stage is pinned to the reference by its own fixtures; here the stages are
    chained through device tensors and checked against the chained oracles."""
    from oracle import predictor_np as PN, decoder_np as O
    B, L, T = 2, 9, 16
    tok = synth.make_tokens(B, L, seed=5200)
    ci = synth.make_chain_inputs(B, L, T, seed=3003)
    with torch.no_grad():
        t_en = _text_encoder()(tok.cuda())
        pred = _dur_predictor()
        d, duration = pred.predict_duration(t_en, ci["s"].cuda())
        dur = ci["dur"].to(torch.int32).cuda()                             # seeded durations summing to T (SURVEY 8(d) cfg 3)
        en = LR.length_regulate(d.transpose(1, 2).contiguous(), dur, T)
        asr = LR.length_regulate(t_en, dur, T)
        f0, n = pred.F0Ntrain(en, ci["s"].cuda())
        cfg = DecoderConfig.hifigan()
        dec = B200Decoder(cfg, "fp32")
        dec.load_state_dict(synth.make_state_dict(cfg, 0, True))
        out = dec.to("cuda").eval()(asr, f0, n, ci["s"].cuda(), noise=ci["noise"].cuda())
    s_np = ci["s"].numpy()
    r_t = PN.text_encoder({k: v.numpy() for k, v in synth.make_text_state_dict(seed=0).items()}, tok.numpy())
    psd = {k: v.numpy() for k, v in synth.make_predictor_state_dict(seed=0, duration=True).items()}
    r_d, r_dur = PN.predict_duration(psd, r_t, s_np)
    assert np.abs(t_en.cpu().numpy() - r_t).max() <= 1e-4 and np.abs(duration.cpu().numpy() - r_dur).max() <= 2e-4
    r_en = O.length_regulate_batch(r_d.transpose(0, 2, 1), ci["dur"].numpy(), T)
    r_asr = O.length_regulate_batch(r_t, ci["dur"].numpy(), T)
    r_f0, r_n = PN.f0n_train(psd, r_en, s_np)
    assert np.abs(f0.cpu().numpy() - r_f0).max() <= 2e-4 and np.abs(n.cpu().numpy() - r_n).max() <= 2e-4
    r_out = O.decoder_forward({k: v.numpy() for k, v in synth.make_state_dict(cfg, 0, True).items()}, cfg, r_asr, r_f0, r_n, s_np,
                              ci["noise"].numpy())
    assert np.abs(out.cpu().numpy() - r_out).max() <= 5e-4


# ---------------------------------------------------------------- padded (ragged) token batches
def test_ragged_text_encoder_and_duration_reference_fixtures():
    """The reference's modules on a padded batch (lengths 9, 6, 4; tests/golden/make_golden_ragged.py): masked_fill_ +
    pack_padded_sequence semantics of models.py:258-285, :485-520, :426-439, eager and as a replayed graph."""
    g = golden("text_ragged_B3_L9_w0_i5101.npz")
    m = _text_encoder()
    tok = torch.from_numpy(g["tokens"]).cuda()
    lengths = torch.from_numpy(g["lengths"]).long()
    mask = torch.arange(9).unsqueeze(0) >= lengths.unsqueeze(1)
    buf = m.set_tap("cnn.0", 3, 9, 512)
    with torch.no_grad():
        out = m(tok, lengths, mask)
    m.clear_taps()
    assert np.abs(G.cf(buf.cpu().numpy()) - g["tap:cnn.0"]).max() <= 2e-5
    assert np.abs(out.cpu().numpy() - g["out"]).max() <= 1e-4
    with torch.no_grad():
        assert torch.equal(m(tok, lengths, cuda_graph=True), out)
        for b, n in enumerate(g["lengths"]):
            n = int(n)
            assert not bool(out[b, :, n:].any())
            assert np.abs(out[b, :, :n].cpu().numpy() - g["single%d" % b][0]).max() <= 1e-4
            one = m(tok[b:b + 1, :n].contiguous())                  # the sentence alone, as inference.py runs it
            assert float((one[0] - out[b, :, :n]).abs().max()) <= 1e-5

    g = golden("dur_ragged_B3_L9_w0_i4101.npz")
    p = _dur_predictor()
    t_en, s = torch.from_numpy(g["t_en"]).cuda(), torch.from_numpy(g["s"]).cuda()
    bufs = {k: p.set_tap(k, 3, 9, 512) for k in ("text_encoder.lstms.0", "lstm")}
    with torch.no_grad():
        d, dur = p.predict_duration(t_en, s, input_lengths=lengths)
    p.clear_taps()
    assert np.abs(bufs["text_encoder.lstms.0"].cpu().numpy() - g["tap:text_encoder.lstms.0"]).max() <= 2e-5
    assert np.abs(bufs["lstm"].cpu().numpy() - g["tap:lstm"]).max() <= 2e-5
    assert np.abs(d.cpu().numpy() - g["d"]).max() <= 1e-4 and np.abs(dur.cpu().numpy() - g["duration"]).max() <= 1e-4
    with torch.no_grad():
        dg, durg = p.predict_duration(t_en, s, input_lengths=lengths, mask=mask, cuda_graph=True)
        assert torch.equal(dg, d) and torch.equal(durg, dur)
        for b, n in enumerate(g["lengths"]):
            n = int(n)
            assert not bool(d[b, n:].any())
            d1, dur1 = p.predict_duration(t_en[b:b + 1, :, :n].contiguous(), s[b:b + 1])
            assert float((d1[0] - d[b, :n]).abs().max()) <= 1e-5 and float((dur1[0] - dur[b, :n]).abs().max()) <= 1e-4


@pytest.mark.parametrize("B,L", [(5, 33), (37, 64)])
def test_ragged_batches_equal_one_sentence_at_a_time(B, L):
    """Seeded ragged lengths (1 .. L, one utterance at L, one at 1): every utterance of the padded batch equals its own
    unpadded call -- both LSTM cluster configurations (4 and 8 utterances per cluster, clusters whose utterances all end
    early), fp32 and fp16 operands -- and the oracle on two of them."""
    from oracle import predictor_np as PN
    rng = np.random.RandomState(B * 100 + L)
    lens = rng.randint(1, L + 1, size=B)
    lens[0], lens[-1] = L, 1
    if B > 8:
        lens[8:16] = rng.randint(1, L // 3, size=8)             # a whole cluster that stops early
    tok = synth.make_tokens(B, L, seed=5600 + B)
    s = synth.make_duration_inputs(B, L, seed=4600 + B)["s"]
    lengths = torch.from_numpy(lens).long()
    te, pr = _text_encoder(), _dur_predictor()
    with torch.no_grad():
        t_en = te(tok.cuda(), lengths)
        d, dur = pr.predict_duration(t_en, s.cuda(), input_lengths=lengths)
        t16 = te(tok.cuda(), lengths, precision="fp16")
        d16, dur16 = pr.predict_duration(t_en, s.cuda(), input_lengths=lengths, precision="fp16")
        for b in range(B):
            n = int(lens[b])
            assert not bool(t_en[b, :, n:].any()) and not bool(d[b, n:].any()) and not bool(t16[b, :, n:].any())
            one = te(tok[b:b + 1, :n].cuda().contiguous())
            assert float((one[0] - t_en[b, :, :n]).abs().max()) <= 1e-5, b
            d1, dur1 = pr.predict_duration(one, s[b:b + 1].cuda())
            assert float((d1[0] - d[b, :n]).abs().max()) <= 2e-5 and float((dur1[0] - dur[b, :n]).abs().max()) <= 2e-4, b
    assert rel_l2(t_en.cpu().numpy(), t16.cpu().numpy()) <= 1e-2 and rel_l2(d.cpu().numpy(), d16.cpu().numpy()) <= 1e-2
    pick = [0, B - 1, B // 2]
    r_t = PN.text_encoder({k: v.numpy() for k, v in synth.make_text_state_dict(seed=0).items()}, tok[pick].numpy(),
                          lengths=lens[pick])
    psd = {k: v.numpy() for k, v in synth.make_predictor_state_dict(seed=0, duration=True).items()}
    r_d, r_dur = PN.predict_duration(psd, r_t, s[pick].numpy(), lengths=lens[pick])
    assert np.abs(t_en[pick].cpu().numpy() - r_t).max() <= 1e-4
    assert np.abs(d[pick].cpu().numpy() - r_d).max() <= 2e-4 and np.abs(dur[pick].cpu().numpy() - r_dur).max() <= 2e-4


# ---------------------------------------------------------------- sizes at the edges of the new modules
@pytest.mark.parametrize("B,L", [(1, 1), (3, 400), (65, 5)])
def test_text_and_duration_sizes_vs_oracle(B, L):
    """One token, the longest sentence the reference splits to (~400 tokens), and a batch that needs two waves of LSTM
    clusters (65 utterances -> 9 groups of 8 per direction > 15 co-resident clusters)."""
    from oracle import predictor_np as PN
    tok = synth.make_tokens(B, max(L, 2), seed=5400 + B)[:, :L]
    s = synth.make_duration_inputs(B, L, seed=4400 + B)["s"]
    with torch.no_grad():
        t_en = _text_encoder()(tok.cuda())
        d, dur = _dur_predictor().predict_duration(t_en, s.cuda())
    nb = min(B, 2)
    r_t = PN.text_encoder({k: v.numpy() for k, v in synth.make_text_state_dict(seed=0).items()}, tok[-nb:].numpy())
    psd = {k: v.numpy() for k, v in synth.make_predictor_state_dict(seed=0, duration=True).items()}
    r_d, r_dur = PN.predict_duration(psd, r_t, s[-nb:].numpy())
    # a single token makes every LayerNorm / InstanceNorm-free path exact but the conv 'same' padding dominant: same bound
    assert np.abs(t_en[-nb:].cpu().numpy() - r_t).max() <= 1e-4
    assert np.abs(d[-nb:].cpu().numpy() - r_d).max() <= 2e-4 and np.abs(dur[-nb:].cpu().numpy() - r_dur).max() <= 2e-4


def test_f0n_10s_and_two_waves_vs_oracle():
    inp = synth.make_predictor_inputs(66, 400, seed=2500)
    f0, n = _run(_predictor(), inp)
    rf0, rn = P.f0n_train(_np_sd(), inp["en"][-1:].numpy(), inp["s"][-1:].numpy())
    assert np.abs(f0[-1:] - rf0).max() <= 1e-4 and np.abs(n[-1:] - rn).max() <= 1e-4
    one_f0, one_n = _run(_predictor(), {"en": inp["en"][:1], "s": inp["s"][:1]})
    assert np.array_equal(f0[:1], one_f0) and np.array_equal(n[:1], one_n)


def test_cuda_graph_replay_of_text_encoder_and_predictor_matches_eager():
    """cuda_graph=True captures each forward once per shape and replays it (graphs.py): bit-identical to the eager call, on
    new inputs of the same shape as well, and re-captured after the weights are re-packed."""
    tok = synth.make_tokens(1, 37, seed=5500).cuda()
    tok2 = synth.make_tokens(1, 37, seed=5501).cuda()
    te, pr = _text_encoder(), _dur_predictor()
    te._graphs.clear()                      # the modules are shared with the tests above
    pr._graphs.clear()
    s = synth.make_duration_inputs(1, 37, seed=4500)["s"].cuda()
    with torch.no_grad():
        for t in (tok, tok2, tok):
            e = te(t)
            assert torch.equal(te(t, cuda_graph=True), e)
            d_e, dur_e = pr.predict_duration(e, s)
            d_g, dur_g = pr.predict_duration(e, s, cuda_graph=True)
            assert torch.equal(d_g, d_e) and torch.equal(dur_g, dur_e)
            en = d_e.transpose(1, 2).contiguous()[:, :, :16].repeat(1, 1, 2).contiguous()      # any [1,640,32] tensor
            f_e, n_e = pr.F0Ntrain(en, s)
            f_g, n_g = pr.F0Ntrain(en, s, cuda_graph=True)
            assert torch.equal(f_g, f_e) and torch.equal(n_g, n_e)
        assert len(te._graphs) == 1 and len(pr._graphs) == 2
        te.load_state_dict(synth.make_text_state_dict(seed=1))         # new weights: old graphs must not be replayed
        e1 = te(tok)
        assert not torch.equal(e1, e) and torch.equal(te(tok, cuda_graph=True), e1)
        te.load_state_dict(synth.make_text_state_dict(seed=0))


# ---------------------------------------------------------------- sentences of one text: StyleTTS2.generate on the device
def test_chained_duration_smoothing_vs_sentence_loop_oracle():
    """st2_smooth_durations_chained: sentence b takes the mean duration of sentence b - 1 as its previous mean (the loop of
    StyleTTS2.generate, inference.py:312-313) -- against the per-sentence oracle chained on the host."""
    from oracle import decoder_np as O
    rng = np.random.RandomState(11)
    lens = np.array([17, 5, 33, 2, 24], np.int32)
    L = int(lens.max())
    dur = np.full((len(lens), L), 7.0, np.float32)
    z = rng.randn(len(lens), L).astype(np.float32)
    for b, n in enumerate(lens):
        dur[b, :n] = 1.0 + 6.0 * rng.rand(n)
    dur[2, 5] = 60.0                                                 # an outlier that gets replaced
    out, means = LR.smooth_durations(torch.from_numpy(dur).cuda(), torch.from_numpy(z).cuda(), t=0.2, speed=1.1, prev_d_mean=3.3,
                                     n_tokens=torch.from_numpy(lens), chained=True)
    out, means = out.cpu().numpy(), means.cpu().numpy()
    prev = 3.3
    for b, n in enumerate(lens):
        ref, prev = O.smooth_durations(dur[b, :n], z[b, :n], 0.2, 1.1, float(prev))
        assert np.abs(out[b, :n] - ref).max() <= 3e-6 * max(1.0, np.abs(ref).max()), b
        assert abs(means[b] - prev) <= 1e-5 and not out[b, n:].any()


def test_synthesizer_batches_what_the_reference_loops_over():
    """pipeline.B200Synthesizer: the sentences of one text as ONE padded batch through TextEncoder / duration half / chained smoothing /
    rounding, then per sentence through the length regulator, F0Ntrain and the Decoder -- against the reference's control flow
    (inference.py:224-272, :303-319: one sentence at a time, prev_d_mean handed on) written out with the same modules at B = 1,
    fp32, shared tapes.  Durations must agree as integers, waveforms to fp32 noise; generate() must be the device
    post-processing of those waveforms, bit for bit."""
    from styletts2_lite_b200.pipeline import B200Synthesizer
    from oracle import postprocess_np as PPN
    te, pr = _text_encoder(), _dur_predictor()
    cfg = DecoderConfig.hifigan()
    dec = B200Decoder(cfg, "fp32")
    dec.load_state_dict(synth.make_state_dict(cfg, 0, True))
    dec = dec.to("cuda").eval()
    lens = [23, 9, 15]
    sents = [synth.make_tokens(1, n, seed=5700 + n)[0] for n in lens]
    s = synth.make_duration_inputs(1, 4, seed=4700)["s"].cuda()
    L = max(lens)
    z = torch.randn(len(lens), L, generator=torch.Generator().manual_seed(9)).cuda()
    seeds = [41, 42, 43]
    syn = B200Synthesizer(te, pr, dec, precision="fp32", cuda_graph=False)
    waves, pred_dur, means = syn.infer_sentences(sents, s, speed=0.9, t=0.2, duration_noise=z, decoder_seeds=seeds)
    prev = 0.0
    with torch.no_grad():
        for b, tok in enumerate(sents):                                # the reference loop
            n = lens[b]
            t_en = te(tok.unsqueeze(0).cuda())
            d, duration = pr.predict_duration(t_en, s)
            duration, mean = LR.smooth_durations(duration, z[b:b + 1, :n].contiguous(), t=0.2, speed=0.9, prev_d_mean=prev)
            prev = float(mean[0])
            pd, tot = LR.round_durations(duration)
            dsm = duration[0].cpu().numpy()
            safe = np.abs(dsm - np.floor(dsm) - 0.5) > 1e-3            # away from rounding ties the integer durations agree
            assert np.array_equal(pd[0].cpu().numpy()[safe], pred_dur[b, :n].cpu().numpy()[safe]), b
            if not np.array_equal(pd[0].cpu().numpy(), pred_dur[b, :n].cpu().numpy()):
                continue                                               # a tie rounded the other way: frame counts differ
            F = int(tot[0])
            asr = LR.length_regulate(t_en, pd, F)
            en = LR.length_regulate(d.transpose(1, 2).contiguous(), pd, F)
            f0, nn_ = pr.F0Ntrain(en, s)
            w = dec(asr, f0, nn_, s, seed=seeds[b]).reshape(-1)
            assert w.shape == waves[b].shape and float((w - waves[b]).abs().max()) <= 2e-4, b
            assert abs(prev - float(means[b])) <= 1e-4
    r, pcm = syn.generate(sents, s, speed=0.9, stabilize=True, duration_noise=z, decoder_seeds=seeds)
    ref_r, ref_pcm = PPN.postprocess([w.cpu().numpy() for w in waves])
    assert np.array_equal(pcm.cpu().numpy(), ref_pcm) and np.array_equal(r.cpu().numpy(), ref_r)
