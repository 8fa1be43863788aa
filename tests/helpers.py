"""Shared helpers for the test-suite (CPU + GPU)."""
import hashlib
import os

import numpy as np

from styletts2_lite_b200.config import DecoderConfig
from styletts2_lite_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def np_state_dict(cfg, seed=0, perturb=True):
    return {k: v.numpy() for k, v in synth.make_state_dict(cfg, seed, perturb).items()}


def np_inputs(B, T, seed, cfg=None, with_noise=True):
    return {k: v.numpy() for k, v in synth.make_inputs(B, T, seed, cfg, with_noise).items()}


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def snr_db(ref, x):
    ref = ref.astype(np.float64)
    err = x.astype(np.float64) - ref
    return 10.0 * np.log10((ref ** 2).sum() / max((err ** 2).sum(), 1e-300))


def rel_l2(ref, x):
    ref = ref.astype(np.float64)
    return float(np.linalg.norm(x.astype(np.float64) - ref) / max(np.linalg.norm(ref), 1e-300))
