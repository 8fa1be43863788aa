"""Drop-in for the F0 / energy half of the reference `ProsodyPredictor`.

Host-side mirror of `ProsodyPredictor.F0Ntrain(x, s)` (models.py:448-461), the call right before the Decoder
(inference.py:267): same sub-module names (`shared`, `F0`, `N`, `F0_proj`, `N_proj`), so the matching entries of a
reference `predictor` state_dict load verbatim (`load_state_dict(..., strict=False)` ignores the duration half --
`text_encoder`, `lstm`, `duration_proj` -- which is row N2 and stays in PyTorch).  The arithmetic runs in the sm_100a
kernels behind include/st2_b200.h (`st2_f0n_*`); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib
from .config import DUR_PREFIXES, F0N_PREFIXES, PredictorConfig, duration_param_specs, predictor_param_specs
from .decoder import _register
from .graphs import GraphReplay


class B200F0NPredictor(nn.Module):
    """`F0Ntrain(en [B, d_hid+style_dim, T], s [B, style_dim]) -> (F0 [B, 2T], N [B, 2T])`.
    precision 'fp32': SIMT everywhere; 'bf16' / 'fp16': convolutions and the LSTM input projection on tcgen05 with
    fp16 operands, recurrence in fp32."""

    def __init__(self, style_dim: int = 128, d_hid: int = 512, nlayers: int = 3, max_dur: int = 50, dropout: float = 0.1,
                 precision: str = "fp32", duration: bool = False):
        super().__init__()
        if precision not in _lib.PREC:
            raise ValueError("precision must be one of %s" % list(_lib.PREC))
        self.cfg = PredictorConfig(d_hid=d_hid, style_dim=style_dim)
        self.precision = precision
        for name, shape, _ in predictor_param_specs(self.cfg):
            _register(self, name, torch.zeros(shape))
        # duration=True adds the duration half (row N2): `text_encoder` (DurationEncoder), `lstm`, `duration_proj`
        self.has_duration = duration
        self._prefixes = F0N_PREFIXES + (DUR_PREFIXES if duration else ())
        if duration:
            for name, shape, _ in duration_param_specs(self.cfg, nlayers, max_dur):
                _register(self, name, torch.zeros(shape))
        self._handle: Optional[C.c_void_p] = None
        self._dirty = True
        self._workspace: Optional[torch.Tensor] = None
        self._taps: Dict[str, torch.Tensor] = {}
        self._graphs = GraphReplay()
        self.train(False)

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._dirty = True
        return r

    def load_state_dict(self, state_dict, strict: bool = True, *a, **k):
        """Accepts a full reference `predictor` state_dict: entries outside F0Ntrain are ignored."""
        sub = {key: v for key, v in state_dict.items() if key.startswith(self._prefixes)}
        r = super().load_state_dict(sub, strict, *a, **k)
        self._dirty = True
        return r

    def _sync(self, device: torch.device) -> None:
        lib = _lib.load()
        self._graphs.clear()                      # finalize re-allocates the packed weights captured graphs point to
        if self._handle is None:
            h = C.c_void_p()
            _lib.check(lib.st2_f0n_create(self.cfg.d_hid, self.cfg.style_dim, C.byref(h)), "st2_f0n_create")
            self._handle = h
        keep = []
        for name, t in self.state_dict().items():
            if t.device != device:
                raise _lib.St2Error("parameter %s is on %s but inputs are on %s; call .to(device)" % (name, t.device, device))
            t = t.detach().float().contiguous()
            keep.append(t)
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
            _lib.check(lib.st2_decoder_set_weight(self._handle, name.encode(), _lib.ptr(t), shape, t.dim()),
                       "st2_decoder_set_weight(%s)" % name)
        stream = torch.cuda.current_stream(device).cuda_stream
        _lib.check(lib.st2_decoder_finalize(self._handle, C.c_void_p(stream)), "st2_decoder_finalize")
        del keep
        self._dirty = False

    def set_tap(self, name: str, B: int, rows: int, C_: int) -> torch.Tensor:
        """Debug tap ('shared', 'F0.1', ...): returns the channels-last [B, rows, C] buffer the next call fills."""
        dev = next(self.parameters()).device
        buf = torch.zeros(B, rows, C_, device=dev, dtype=torch.float32)
        if self._handle is None or self._dirty:
            self._sync(dev)
        _lib.check(_lib.load().st2_decoder_set_tap(self._handle, name.encode(), _lib.ptr(buf), buf.numel()), "set_tap")
        self._taps[name] = buf
        return buf

    def clear_taps(self) -> None:
        for name in list(self._taps):
            _lib.load().st2_decoder_set_tap(self._handle, name.encode(), None, 0)
        self._taps.clear()

    def _launch_f0n(self, x_, s_, ws, B, T, prec):
        lib = _lib.load()
        dev = x_.device
        f0 = torch.empty(B, 2 * T, dtype=torch.float32, device=dev)
        n = torch.empty(B, 2 * T, dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.st2_f0n_forward(self._handle, _lib.ptr(x_), _lib.ptr(s_), _lib.ptr(f0), _lib.ptr(n), B, T, prec,
                                       _lib.ptr(ws), ws.numel(), C.c_void_p(stream)), "st2_f0n_forward")
        return f0, n

    def _launch_dur(self, x_, s_, lens, ws, B, L, prec):
        lib = _lib.load()
        dev = x_.device
        d = torch.empty(B, L, self.cfg.d_hid + self.cfg.style_dim, dtype=torch.float32, device=dev)
        duration = torch.empty(B, L, dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.st2_dur_forward_ragged(self._handle, _lib.ptr(x_), _lib.ptr(s_), _lib.ptr(lens) if lens is not None else None,
                                              _lib.ptr(d), _lib.ptr(duration), B, L, prec, _lib.ptr(ws), ws.numel(),
                                              C.c_void_p(stream)), "st2_dur_forward")
        return d, duration

    def F0Ntrain(self, x: torch.Tensor, s: torch.Tensor, precision: Optional[str] = None, cuda_graph: bool = False):
        """`cuda_graph=True` replays a graph captured once per (B, T, precision) -- one-sentence latency (graphs.py)."""
        if self.training:
            raise RuntimeError("B200F0NPredictor is inference-only (dropout p=0.2 of models.py:409-416 is not implemented); call .eval()")
        if not x.is_cuda:
            raise _lib.St2Error("B200F0NPredictor has no CPU path: inputs must be CUDA tensors")
        lib = _lib.load()
        dev = x.device
        B, Cin, T = x.shape
        if Cin != self.cfg.d_hid + self.cfg.style_dim or tuple(s.shape) != (B, self.cfg.style_dim):
            raise ValueError("expected x [B,%d,T] and s [B,%d]; got %s %s" %
                             (self.cfg.d_hid + self.cfg.style_dim, self.cfg.style_dim, tuple(x.shape), tuple(s.shape)))
        prec = _lib.PREC[precision or self.precision]
        with torch.cuda.device(dev):
            if self._dirty or self._handle is None:
                self._sync(dev)
            x_, s_ = x.detach().float().contiguous(), s.detach().float().contiguous()
            need = _lib.check(lib.st2_f0n_workspace_bytes(self._handle, B, T, prec), "st2_f0n_workspace_bytes")
            if cuda_graph and not self._taps:
                return self._graphs.run(("f0n", B, T, prec, dev.index), (x_, s_),
                                        lambda: torch.empty(need, dtype=torch.uint8, device=dev),
                                        lambda ins, ws: self._launch_f0n(ins[0], ins[1], ws, B, T, prec))
            if self._workspace is None or self._workspace.numel() < need or self._workspace.device != dev:
                self._workspace = None
                self._workspace = torch.empty(need, dtype=torch.uint8, device=dev)
            return self._launch_f0n(x_, s_, self._workspace, B, T, prec)

    def predict_duration(self, t_en: torch.Tensor, s: torch.Tensor, precision: Optional[str] = None, cuda_graph: bool = False,
                         input_lengths: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None):
        """inference.py:242-245:
        `d = predictor.text_encoder(t_en, s, lengths, mask); x, _ = predictor.lstm(d);
        duration = sigmoid(predictor.duration_proj(x)).sum(-1)`.
        t_en [B, d_hid, L], s [B, style_dim] -> (d [B, L, d_hid+style_dim], duration [B, L]).
        With input_lengths (and / or its mask) the batch is padded and handled as ProsodyPredictor.forward does
        (models.py:422-442): masked rows of d are zero, every LSTM is packed, utterance b equals its own B = 1 call."""
        if not self.has_duration:
            raise RuntimeError("construct B200F0NPredictor(duration=True) to get the duration half")
        if self.training:
            raise RuntimeError("B200F0NPredictor is inference-only; call .eval()")
        if not t_en.is_cuda:
            raise _lib.St2Error("B200F0NPredictor has no CPU path: inputs must be CUDA tensors")
        lib = _lib.load()
        dev = t_en.device
        B, Cin, L = t_en.shape
        if Cin != self.cfg.d_hid or tuple(s.shape) != (B, self.cfg.style_dim):
            raise ValueError("expected t_en [B,%d,L] and s [B,%d]; got %s %s" % (self.cfg.d_hid, self.cfg.style_dim,
                                                                              tuple(t_en.shape), tuple(s.shape)))
        from .text_encoder import ragged_lengths
        lens = ragged_lengths(input_lengths, mask, B, L, dev)
        prec = _lib.PREC[precision or self.precision]
        with torch.cuda.device(dev):
            if self._dirty or self._handle is None:
                self._sync(dev)
            x_, s_ = t_en.detach().float().contiguous(), s.detach().float().contiguous()
            need = _lib.check(lib.st2_dur_workspace_bytes(self._handle, B, L, prec), "st2_dur_workspace_bytes")
            if cuda_graph and not self._taps:
                ins = (x_, s_) if lens is None else (x_, s_, lens)
                return self._graphs.run(("dur", B, L, prec, dev.index, lens is not None), ins,
                                        lambda: torch.empty(need, dtype=torch.uint8, device=dev),
                                        lambda i, ws: self._launch_dur(i[0], i[1], i[2] if len(i) > 2 else None, ws, B, L, prec))
            if self._workspace is None or self._workspace.numel() < need or self._workspace.device != dev:
                self._workspace = None
                self._workspace = torch.empty(need, dtype=torch.uint8, device=dev)
            return self._launch_dur(x_, s_, lens, self._workspace, B, L, prec)

    def last_launch_count(self) -> int:
        return int(_lib.load().st2_decoder_last_launch_count(self._handle)) if self._handle else 0

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().st2_decoder_destroy(self._handle)
                self._handle = None
        except Exception:
            pass
