"""waveform post-processing + wire format (inference.py:314-319, Demo/infer.py:51-54).
CPU: the numpy restatement against the fixture produced by executing the reference's statements (make_golden_post.py).
GPU: st2_postprocess through the C ABI, bit-exact against the restatement (equal lengths, ragged, short, empty, silent)."""
import hashlib
import importlib.util
import io
import os
import wave

import numpy as np
import pytest
import torch

from oracle import postprocess_np as P
from helpers import GOLDEN, golden

_spec = importlib.util.spec_from_file_location("make_golden_post", os.path.join(GOLDEN, "make_golden_post.py"))
MGP = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(MGP)


def test_oracle_matches_reference_statement_fixture():
    g = golden("post_3sent.npz")
    wavs = MGP.sentences()
    r, pcm = P.postprocess(wavs)
    assert r.dtype == np.float64 and pcm.dtype == np.int16
    assert hashlib.sha256(np.ascontiguousarray(r).tobytes()).hexdigest() == str(g["r_sha256"])
    assert np.array_equal(pcm, g["pcm"])
    assert np.array_equal(r[3990:4200], g["r_head"]) and np.array_equal(r[-4200:-3990], g["r_tail"])
    assert np.abs(r).max() == 1.0 and np.all(r[:4000] == 0) and np.all(r[-4000:] == 0)
    assert abs(int(pcm[4000 + 16000 + 1000])) == 32767        # the kept peak of sentence 1 (sample 5000 of it) maps to full scale


def test_wav_container_is_readable_pcm16_24k():
    _, pcm = P.postprocess(MGP.sentences())
    with wave.open(io.BytesIO(P.wav_bytes(pcm)), "rb") as f:
        assert (f.getnchannels(), f.getsampwidth(), f.getframerate(), f.getnframes()) == (1, 2, 24000, len(pcm))
        assert np.array_equal(np.frombuffer(f.readframes(len(pcm)), "<i2"), pcm)


def test_half_even_rounding_of_the_pcm_conversion():
    # x * 32767 exactly halfway: lrint rounds to even (libsndfile uses lrint, not floor(x + 0.5))
    x = np.array([0.5 / 32767.0, 1.5 / 32767.0, 2.5 / 32767.0, -0.5 / 32767.0, -1.5 / 32767.0, 1.0, -1.0])
    assert P.pcm16(x).tolist() == np.rint(x * 32767.0).astype(np.int16).tolist()
    assert P.pcm16(np.array([1.0, -1.0])).tolist() == [32767, -32767]


# ---------------------------------------------------------------- GPU
def _gpu_post(wavs, S=None, lengths=True):
    from styletts2_lite_b200 import postprocess as PP
    S = S or max(len(w) for w in wavs)
    t = torch.zeros(len(wavs), 1, S)
    for i, w in enumerate(wavs):
        t[i, 0, :len(w)] = torch.from_numpy(w)
        t[i, 0, len(w):] = 7.0                       # garbage past the sentence must never be read
    ln = torch.tensor([len(w) for w in wavs], dtype=torch.int32) if lengths else None
    r, pcm = PP.assemble(t.cuda(), None if ln is None else ln.cuda())
    torch.cuda.synchronize()
    return r.cpu().numpy(), pcm.cpu().numpy()


@pytest.mark.gpu
def test_gpu_postprocess_bit_exact_fixture_and_ragged():
    g = golden("post_3sent.npz")
    wavs = MGP.sentences()
    r, pcm = _gpu_post(wavs)
    assert r.dtype == np.float64 and pcm.dtype == np.int16
    assert hashlib.sha256(np.ascontiguousarray(r).tobytes()).hexdigest() == str(g["r_sha256"])
    assert np.array_equal(pcm, g["pcm"])


@pytest.mark.gpu
@pytest.mark.parametrize("lens", [(24000,) * 4, (48000, 8001, 8000, 7999, 300, 24600), (9000,), (144000, 60000)])
def test_gpu_postprocess_bit_exact_vs_oracle(lens):
    wavs = MGP.sentences(seed=5 + len(lens), lens=lens) if len(lens) >= 3 and min(lens) > 5000 else None
    if wavs is None:
        rng = np.random.default_rng(len(lens))
        wavs = [np.tanh(rng.standard_normal(n) * 0.5).astype(np.float32) for n in lens]
    r_ref, pcm_ref = P.postprocess(wavs)
    equal = len(set(lens)) == 1
    r, pcm = _gpu_post(wavs, lengths=not equal)
    assert r.shape == r_ref.shape
    assert np.array_equal(r, r_ref) and np.array_equal(pcm, pcm_ref)


@pytest.mark.gpu
def test_gpu_postprocess_empty_and_silent_batches():
    from styletts2_lite_b200 import postprocess as PP
    r, pcm = PP.assemble(torch.zeros(0, 1, 24000, device="cuda"))
    assert r.numel() == 8000 and pcm.numel() == 8000 and float(r.abs().max()) == 0.0 and int(pcm.abs().max()) == 0
    r, pcm = PP.assemble(torch.zeros(2, 1, 12000, device="cuda"))
    assert r.numel() == 16000 and float(r.abs().max()) == 0.0 and int(pcm.abs().max()) == 0     # silent input: zeros, not NaN
    with pytest.raises(Exception):
        PP.assemble(torch.zeros(1, 1, 24000))                                                     # CPU tensor: no fallback
    # the container helper
    wavs = MGP.sentences()
    _, pcm = _gpu_post(wavs)
    assert PP.wav_bytes(torch.from_numpy(pcm)) == P.wav_bytes(pcm)
