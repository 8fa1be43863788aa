"""Fixture for SURVEY.md 8(f) N4 (waveform post-processing + wire format):  python tests/golden/make_golden_post.py

inference.py cannot be imported here (librosa / nltk / noisereduce are missing and it downloads at import time) and `soundfile`
is not installed, so -- like the length-regulator fixture of make_golden.py -- this script executes the reference's statements
exactly as written (inference.py:314-319, Demo/infer.py:51) on seeded sentence waveforms, and restates the one call it cannot
make, soundfile.write (Demo/infer.py:54), from libsndfile's published conversion (pcm.c d2s_array: lrint(x * 0x7FFF)).
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_statements(list_of_sentence_wavs):
    list_wav = []
    for wav in list_of_sentence_wavs:
        wav = wav[4000:-4000]  # Remove weird pulse and silent tokens          (inference.py:315)
        list_wav.append(wav)
    final_wav = np.concatenate(list_wav)                                       # (inference.py:318)
    final_wav = np.concatenate([np.zeros([4000]), final_wav, np.zeros([4000])], axis=0)  # add padding (inference.py:319)
    r = final_wav
    r = r / np.max(np.abs(r))  # Normalize                                      (Demo/infer.py:51)
    return r


def sentences(seed=77, lens=(600 * 40, 600 * 23, 600 * 61)):
    """Seeded sentence waveforms (multiples of 600 samples, as the decoder produces; decoder-like amplitudes).  The tests
    regenerate them with this function instead of storing them."""
    rng = np.random.default_rng(seed)
    wavs = [np.tanh(rng.standard_normal(n) * 0.4).astype(np.float32) for n in lens]
    wavs[1][5000] = np.float32(0.98765)           # the peak sits inside the kept part of sentence 1
    wavs[2][100] = np.float32(-0.999)             # ... and a larger one inside a trimmed part must not count
    return wavs


def main():
    import hashlib
    wavs = sentences()
    r = reference_statements(wavs)
    pcm = np.rint(r * 32767.0).astype(np.int16)   # libsndfile d2s_array, normalisation on (soundfile.write default for float data)
    np.savez_compressed(os.path.join(HERE, "post_3sent.npz"), pcm=pcm, r_head=r[3990:4200], r_tail=r[-4200:-3990],
                        r_sha256=np.array(hashlib.sha256(np.ascontiguousarray(r).tobytes()).hexdigest()))
    print("post_3sent", r.shape, r.dtype, int(np.abs(pcm).max()), float(np.abs(r).max()))


if __name__ == "__main__":
    main()
