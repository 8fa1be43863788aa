"""Duration -> frame length regulation (inference.py:257-268) on the GPU.

    duration = smooth(duration, t, prev_d_mean) / speed   # st2_smooth_durations (inference.py:248-255)
    pred_dur = round(duration).clamp(min=1)          # st2_round_durations
    asr = t_en @ alignment ; en = d^T @ alignment    # st2_length_regulate (bit-exact gather)

Batched and ragged: utterance b uses its first n_tokens[b] tokens and produces
sum(pred_dur[b]) frames; outputs are zero beyond that.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def round_durations(duration: torch.Tensor, n_tokens: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """duration [B,L] fp32 (CUDA) -> (pred_dur [B,L] int32, total_frames [B] int32)."""
    if not duration.is_cuda:
        raise _lib.St2Error("length regulator has no CPU path: inputs must be CUDA tensors")
    lib = _lib.load()
    duration = duration.detach().float().contiguous()
    B, L = duration.shape
    dur = torch.empty(B, L, dtype=torch.int32, device=duration.device)
    tot = torch.empty(B, dtype=torch.int32, device=duration.device)
    nt = None if n_tokens is None else n_tokens.to(device=duration.device, dtype=torch.int32).contiguous()
    with torch.cuda.device(duration.device):
        _lib.check(lib.st2_round_durations(_lib.ptr(duration), _lib.ptr(nt), _lib.ptr(dur), _lib.ptr(tot), B, L,
                                           _stream(duration.device)), "st2_round_durations")
    return dur, tot


def smooth_durations(duration: torch.Tensor, noise: Optional[torch.Tensor] = None, t: float = 0.1, speed: float = 1.0,
                     prev_d_mean=0.0, n_tokens: Optional[torch.Tensor] = None, chained: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """inference.py:248-255 on the device, per utterance: mix with N(mean or prev_d_mean, std) draws (weight t), replace the
    z-score outliers of duration[1:-2], divide by speed.  duration [B,L] fp32 (CUDA); noise [B,L] the N(0,1) tape that stands in
    for `normal_` (drawn with torch.randn on the device when omitted and t > 0); prev_d_mean a float or a [B] tensor (0 = none).
    Returns (smoothed duration [B,L], mean duration [B] -- inference.py:272, the next split's prev_d_mean).
    chained=True: the B rows are the sentences of one text in order and sentence b takes the mean of sentence b - 1 as its previous
    mean (the loop of StyleTTS2.generate, inference.py:312-313); prev_d_mean then seeds sentence 0 only."""
    if not duration.is_cuda:
        raise _lib.St2Error("length regulator has no CPU path: inputs must be CUDA tensors")
    speed = min(max(float(speed), 0.0001), 2.0)                   # inference.py:226
    if not 0.0 <= float(t) <= 1.0:
        raise ValueError("t must lie in [0, 1]")
    lib = _lib.load()
    dev = duration.device
    duration = duration.detach().float().contiguous()
    B, L = duration.shape
    if noise is None and t > 0:
        noise = torch.randn(B, L, device=dev, dtype=torch.float32)
    if noise is not None:
        noise = noise.to(device=dev, dtype=torch.float32).contiguous()
        if noise.shape != (B, L):
            raise ValueError("noise must be [B,L]")
    if isinstance(prev_d_mean, torch.Tensor):
        prev = prev_d_mean.to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        if prev.numel() != B:
            raise ValueError("prev_d_mean must have one entry per utterance")
    else:
        prev = None if float(prev_d_mean) == 0.0 else torch.full((B,), float(prev_d_mean), device=dev, dtype=torch.float32)
    fn = lib.st2_smooth_durations_chained if chained else lib.st2_smooth_durations
    nt = None if n_tokens is None else n_tokens.to(device=dev, dtype=torch.int32).contiguous()
    out = torch.empty_like(duration)
    mean = torch.empty(B, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(fn(_lib.ptr(duration), _lib.ptr(nt), _lib.ptr(noise), _lib.ptr(prev), float(t), speed,
                      _lib.ptr(out), _lib.ptr(mean), B, L, _stream(dev)), "st2_smooth_durations")
    return out, mean


def length_regulate(src: torch.Tensor, pred_dur: torch.Tensor, n_frames: int, channels_last: bool = False) -> torch.Tensor:
    """src [B,C,L] fp32, pred_dur [B,L] int32 -> [B,C,n_frames] (or [B,n_frames,C])."""
    if not src.is_cuda:
        raise _lib.St2Error("length regulator has no CPU path: inputs must be CUDA tensors")
    lib = _lib.load()
    src = src.detach().float().contiguous()
    pred_dur = pred_dur.to(device=src.device, dtype=torch.int32).contiguous()
    B, Cc, L = src.shape
    if pred_dur.shape != (B, L):
        raise ValueError("pred_dur must be [B,L]")
    shape = (B, n_frames, Cc) if channels_last else (B, Cc, n_frames)
    out = torch.empty(shape, dtype=torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        _lib.check(lib.st2_length_regulate(_lib.ptr(src), _lib.ptr(pred_dur), _lib.ptr(out), B, Cc, L, n_frames,
                                           1 if channels_last else 0, _stream(src.device)), "st2_length_regulate")
    return out


def regulate(duration: torch.Tensor, t_en: torch.Tensor, d: torch.Tensor, n_tokens: Optional[torch.Tensor] = None):
    """The four lines inference.py:257-268 as one call: duration [B,L], t_en [B,512,L],
    d [B,L,640] -> (asr [B,512,F], en [B,640,F], pred_dur, total_frames), F = max_b frames."""
    dur, tot = round_durations(duration, n_tokens)
    F = int(tot.max().item())
    asr = length_regulate(t_en, dur, F)
    en = length_regulate(d.transpose(-1, -2), dur, F)
    return asr, en, dur, tot
