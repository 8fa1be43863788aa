"""Deterministic synthetic weights and inputs (no checkpoint, no dataset).

SURVEY.md §8(d) "Synthetic inputs": random-init weights of the reference
architecture and seeded asr / F0 / N / style / noise tensors.  Weights are
drawn by this module (torch CPU generator, reproducible across machines with
the same torch build) in the reference's own state_dict schema, so the very
same dict can be `load_state_dict`-ed into the reference Decoder (that is how
tests/golden/make_golden.py produced the committed fixtures) and into the
B200 drop-in.

`perturb=True` moves weight_g away from ||v||, alpha away from 1 and scales
biases, so that the weight-norm fold, the Snake alphas and every bias path
are actually exercised by parity tests (with the reference's init,
g == ||v|| and alpha == 1, those bugs would be invisible).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

from .config import DecoderConfig, PredictorConfig, param_specs, predictor_param_specs, duration_param_specs, TextEncoderConfig, text_encoder_param_specs, buffer_specs, HARMONICS, HIDDEN_DIM, STYLE_DIM


def _fan_in(shape):
    f = 1
    for d in shape[1:]:
        f *= d
    return max(f, 1)


def make_state_dict(cfg: DecoderConfig, seed: int = 0, perturb: bool = True) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = _draw(param_specs(cfg), g, perturb)
    if cfg.is_vocos:
        sd["generator.stft.istft.window"] = torch.hann_window(cfg.gen_istft_n_fft)       # vocos.py:192 (periodic)
    else:
        for name, shape in buffer_specs(cfg):
            sd[name] = stft_buffers(cfg)[name]
    return {k: v.float().contiguous() for k, v in sd.items()}


def make_predictor_state_dict(cfg: PredictorConfig | None = None, seed: int = 0, perturb: bool = True,
                              duration: bool = False) -> Dict[str, torch.Tensor]:
    """The F0Ntrain subset of a ProsodyPredictor state_dict (models.py:407-419), drawn like make_state_dict;
    `duration=True` adds the duration half (text_encoder / lstm / duration_proj, models.py:399-405), drawn after it from
    the same generator so the F0Ntrain tensors do not depend on the flag."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    cfg = cfg or PredictorConfig()
    sd = _draw(predictor_param_specs(cfg), g, perturb)
    if duration:
        sd.update(_draw(duration_param_specs(cfg), g, perturb))
    return {k: v.float().contiguous() for k, v in sd.items()}


def make_duration_inputs(B: int, L: int, seed: int = 4000, cfg: PredictorConfig | None = None) -> Dict[str, torch.Tensor]:
    """t_en [B, d_hid, L] (TextEncoder output, inference.py:239) and the style s [B, style_dim]."""
    cfg = cfg or PredictorConfig()
    g = torch.Generator(device="cpu").manual_seed(seed)
    return {"t_en": torch.randn(B, cfg.d_hid, L, generator=g), "s": torch.randn(B, cfg.style_dim, generator=g)}


def make_predictor_inputs(B: int, T: int, seed: int = 2000, cfg: PredictorConfig | None = None) -> Dict[str, torch.Tensor]:
    """en [B, d_hid+style_dim, T] (the length-regulated DurationEncoder output, inference.py:267) and s [B, style_dim]."""
    cfg = cfg or PredictorConfig()
    g = torch.Generator(device="cpu").manual_seed(seed)
    return {"en": torch.randn(B, cfg.d_hid + cfg.style_dim, T, generator=g), "s": torch.randn(B, cfg.style_dim, generator=g)}


def _draw(specs, g, perturb: bool) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    # pass 1: everything that does not depend on another tensor
    for name, shape, kind in specs:
        if kind == "conv":
            bound = 1.0 / math.sqrt(_fan_in(shape))
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "linear":
            bound = 1.0 / math.sqrt(shape[1])
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind.startswith("linear_bias"):
            bound = 1.0 / math.sqrt(int(kind.split(":")[1]))
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "alpha":
            if perturb:
                sd[name] = 0.6 + 0.8 * torch.rand(shape, generator=g)
            else:
                sd[name] = torch.ones(shape)
        elif kind == "trunc02":                             # Generator._init_weights: trunc_normal_(std=0.02) (vocos.py:154-157)
            w = torch.randn(shape, generator=g).clamp_(-2, 2) * 0.02
            sd[name] = w * 3.0 if perturb else w                # the reference init is too small to exercise the blocks
        elif kind == "small_bias":                          # constant_(bias, 0) in the reference init
            sd[name] = 0.05 * torch.randn(shape, generator=g) if perturb else torch.zeros(shape)
        elif kind.startswith("layer_scale:"):
            v = float(kind.split(":")[1])
            sd[name] = v * (0.5 + torch.rand(shape, generator=g)) if perturb else torch.full(shape, v)
        elif kind == "normal":                              # nn.Embedding init
            sd[name] = torch.randn(shape, generator=g)
        elif kind == "gamma":                               # LayerNorm scale: 1 in the reference init
            sd[name] = 0.8 + 0.4 * torch.rand(shape, generator=g) if perturb else torch.ones(shape)
        elif kind == "beta":
            sd[name] = 0.2 * (torch.rand(shape, generator=g) - 0.5) if perturb else torch.zeros(shape)
        elif kind.startswith("lstm:"):                      # nn.LSTM init: U(-1/sqrt(hidden), 1/sqrt(hidden))
            bound = 1.0 / math.sqrt(int(kind.split(":")[1]))
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    # pass 2: g and conv biases
    for name, shape, kind in specs:
        if kind.startswith("g:"):
            v = sd[kind[2:]]
            nrm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(shape)
            if perturb:
                nrm = nrm * (0.85 + 0.3 * torch.rand(shape, generator=g))
            sd[name] = nrm
        elif kind.startswith("bias:"):
            w = sd[kind[5:]]
            bound = 1.0 / math.sqrt(_fan_in(w.shape))
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


def stft_buffers(cfg: DecoderConfig) -> Dict[str, torch.Tensor]:
    """The five CustomSTFT buffers, built the way istftnet.py:141-203 builds them
    (periodic Hann, float64 DFT basis rounded to fp32, inverse scaled by 1/n_fft)."""
    n = cfg.gen_istft_n_fft
    bins = n // 2 + 1
    window = torch.hann_window(n, periodic=True, dtype=torch.float32)
    wn = window.numpy()
    kk = np.arange(bins)
    nn_ = np.arange(n)
    ang = 2 * np.pi * np.outer(kk, nn_) / n
    fr = torch.from_numpy(np.cos(ang) * wn).float().unsqueeze(1)
    fi = torch.from_numpy(-np.sin(ang) * wn).float().unsqueeze(1)
    ang_t = 2 * np.pi * np.outer(nn_, kk) / n
    inv_w = wn * (1.0 / n)
    br = torch.from_numpy(np.cos(ang_t).T * inv_w).float().unsqueeze(1)
    bi = torch.from_numpy(np.sin(ang_t).T * inv_w).float().unsqueeze(1)
    p = "generator.stft."
    return {p + "window": window, p + "weight_forward_real": fr, p + "weight_forward_imag": fi,
            p + "weight_backward_real": br, p + "weight_backward_imag": bi}


def make_inputs(B: int, T: int, seed: int = 1000, cfg: DecoderConfig | None = None,
                with_noise: bool = True) -> Dict[str, torch.Tensor]:
    """asr [B,512,T], F0_curve [B,2T] (80..280 Hz with ~20 % unvoiced runs set to 0),
    N [B,2T], s [B,128], noise [B,S,9] (the SineGen `randn_like` draw, hifigan.py:213)."""
    cfg = cfg or DecoderConfig()
    g = torch.Generator(device="cpu").manual_seed(seed)
    asr = torch.randn(B, cfg.dim_in, T, generator=g)
    f0 = 80.0 + 200.0 * torch.rand(B, 2 * T, generator=g)
    # unvoiced runs: contiguous stretches of 3..12 frames until ~20 % is covered
    L = 2 * T
    for b in range(B):
        target = int(0.2 * L)
        covered = 0
        guard = 0
        while covered < target and guard < 1000:
            guard += 1
            ln = int(torch.randint(3, 13, (1,), generator=g))
            st = int(torch.randint(0, max(L - ln, 1), (1,), generator=g))
            f0[b, st:st + ln] = 0.0
            covered += ln
    n = torch.randn(B, 2 * T, generator=g)
    s = torch.randn(B, cfg.style_dim, generator=g)
    out = {"asr": asr, "F0_curve": f0, "N": n, "s": s}
    if with_noise:
        S = cfg.samples_per_frame * T
        out["noise"] = torch.randn(B, S, HARMONICS, generator=g)
    return out


def make_durations(B: int, L: int, F: int, seed: int = 7) -> torch.Tensor:
    """Seeded integer durations [B,L], each >= 1, each row summing to F
    (SURVEY.md §8(d) cfg 3: synthetic durations replace pred_dur after inference.py:257)."""
    assert F >= L
    g = torch.Generator(device="cpu").manual_seed(seed)
    dur = torch.ones(B, L, dtype=torch.int64)
    for b in range(B):
        extra = torch.randint(0, L, (F - L,), generator=g)
        dur[b] += torch.bincount(extra, minlength=L)
    return dur


def make_chain_inputs(B: int, L: int, T: int, seed: int = 3003) -> Dict[str, torch.Tensor]:
    """Inputs of the chained slice inference.py:257-270 (cfg 3 after the text modules): integer durations [B,L]
    summing to T, DurationEncoder output d [B,L,640], TextEncoder output t_en [B,512,L], style s, SineGen noise."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    return {"dur": make_durations(B, L, T, seed=seed + 8), "d": torch.randn(B, L, HIDDEN_DIM + STYLE_DIM, generator=g),
            "t_en": torch.randn(B, HIDDEN_DIM, L, generator=g), "s": torch.randn(B, STYLE_DIM, generator=g),
            "noise": torch.randn(B, 600 * T, HARMONICS, generator=g)}


def make_text_state_dict(cfg: TextEncoderConfig | None = None, seed: int = 0, perturb: bool = True) -> Dict[str, torch.Tensor]:
    """A TextEncoder state_dict (models.py:238-256), drawn like make_state_dict."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = _draw(text_encoder_param_specs(cfg or TextEncoderConfig()), g, perturb)
    return {k: v.float().contiguous() for k, v in sd.items()}


def make_tokens(B: int, L: int, seed: int = 5000, n_symbols: int = 178) -> torch.Tensor:
    """Token ids [B, L] int64 with the reference's leading / trailing pad token 0 (inference.py:231-232)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    t = torch.randint(1, n_symbols, (B, L), generator=g)
    t[:, 0] = 0
    t[:, -1] = 0
    return t
