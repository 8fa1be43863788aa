"""Drop-in for the reference `TextEncoder` (models.py:238-285, called at inference.py:239).

Same constructor (`channels, kernel_size, depth, n_symbols`), same `state_dict` keys (`embedding.weight`,
`cnn.{i}.0.weight_g/_v/bias`, `cnn.{i}.1.gamma/beta`, `lstm.*`), same `forward(x, input_lengths, m) -> [B, channels, L]`,
padded batches included: tokens behind `input_lengths[b]` are masked where the reference masks (models.py:262, :266, :283) and
the LSTM runs over each utterance's own length (pack_padded_sequence, models.py:270-277).  The arithmetic runs in the sm_100a
kernels behind include/st2_b200.h (`st2_text_*`); there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib
from .config import TextEncoderConfig, text_encoder_param_specs
from .decoder import _register
from .graphs import GraphReplay


def ragged_lengths(input_lengths, m, B: int, L: int, device) -> Optional[torch.Tensor]:
    """The int32 device vector the `_ragged` entry points take, or None when the batch has no padding.  input_lengths [B] as the
    reference passes it; m [B, L'] the reference's length_to_mask(input_lengths) (models.py:296-299), checked against it."""
    if input_lengths is None and m is None:
        return None
    if input_lengths is None:
        mm = m.to("cpu").bool()
        lens = (~mm).sum(dim=1)
    else:
        lens = input_lengths.detach().to("cpu").to(torch.int64).reshape(-1)
    if lens.numel() != B:
        raise ValueError("input_lengths must have one entry per utterance (%d), got %d" % (B, lens.numel()))
    if int(lens.min()) < 1 or int(lens.max()) > L:
        raise ValueError("input_lengths must lie in [1, %d]" % L)        # pack_padded_sequence rejects length 0 as well
    if m is not None:
        mm = m.to("cpu").bool()
        want = torch.arange(mm.shape[-1]).unsqueeze(0) >= lens.unsqueeze(1)
        if mm.shape[0] != B or mm.shape[-1] > L or not bool((mm == want).all()):
            raise ValueError("the mask is not length_to_mask(input_lengths)")
    if bool((lens == L).all()):
        return None
    return lens.to(torch.int32).to(device)


class B200TextEncoder(nn.Module):
    def __init__(self, channels: int = 512, kernel_size: int = 5, depth: int = 3, n_symbols: int = 178, actv=None,
                 precision: str = "fp32"):
        super().__init__()
        if precision not in _lib.PREC:
            raise ValueError("precision must be one of %s" % list(_lib.PREC))
        self.cfg = TextEncoderConfig(channels=channels, kernel_size=kernel_size, depth=depth, n_symbols=n_symbols)
        self.precision = precision
        for name, shape, kind in text_encoder_param_specs(self.cfg):
            _register(self, name, torch.ones(shape) if kind == "gamma" else torch.zeros(shape))
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
        r = super().load_state_dict(state_dict, strict, *a, **k)
        self._dirty = True
        return r

    def _sync(self, device: torch.device) -> None:
        lib = _lib.load()
        self._graphs.clear()                      # finalize re-allocates the packed weights captured graphs point to
        if self._handle is None:
            h = C.c_void_p()
            _lib.check(lib.st2_text_create(self.cfg.channels, self.cfg.kernel_size, self.cfg.depth, self.cfg.n_symbols,
                                           C.byref(h)), "st2_text_create")
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

    def _launch(self, tok, lens, ws, B, L, prec):
        lib = _lib.load()
        dev = tok.device
        out = torch.empty(B, self.cfg.channels, L, dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.st2_text_forward_ragged(self._handle, _lib.ptr(tok), _lib.ptr(lens) if lens is not None else None,
                                               _lib.ptr(out), B, L, prec, _lib.ptr(ws), ws.numel(), C.c_void_p(stream)),
                   "st2_text_forward")
        return out

    def forward(self, x: torch.Tensor, input_lengths: Optional[torch.Tensor] = None, m: Optional[torch.Tensor] = None,
                precision: Optional[str] = None, cuda_graph: bool = False) -> torch.Tensor:
        """x: token ids [B, L] (int64); input_lengths [B] and the padding mask m [B, L] (True behind an utterance) as in the
        reference call (inference.py:236-239).  m, when given, must be the mask of input_lengths (length_to_mask)."""
        if self.training:
            raise RuntimeError("B200TextEncoder is inference-only (Dropout(0.2) of models.py:250 is not implemented); call .eval()")
        if not x.is_cuda:
            raise _lib.St2Error("B200TextEncoder has no CPU path: inputs must be CUDA tensors")
        B, L = x.shape
        lens = ragged_lengths(input_lengths, m, B, L, x.device)
        if int(x.min()) < 0 or int(x.max()) >= self.cfg.n_symbols:
            raise IndexError("token id out of range [0, %d)" % self.cfg.n_symbols)          # nn.Embedding raises as well
        lib = _lib.load()
        dev = x.device
        prec = _lib.PREC[precision or self.precision]
        with torch.cuda.device(dev):
            if self._dirty or self._handle is None:
                self._sync(dev)
            tok = x.detach().to(torch.int64).contiguous()
            need = _lib.check(lib.st2_text_workspace_bytes(self._handle, B, L, prec), "st2_text_workspace_bytes")
            if cuda_graph and not self._taps:           # graph captured once per (B, L, precision): graphs.py
                ins = (tok,) if lens is None else (tok, lens)
                return self._graphs.run(("text", B, L, prec, dev.index, lens is not None), ins,
                                        lambda: torch.empty(need, dtype=torch.uint8, device=dev),
                                        lambda i, ws: (self._launch(i[0], i[1] if len(i) > 1 else None, ws, B, L, prec),))[0]
            if self._workspace is None or self._workspace.numel() < need or self._workspace.device != dev:
                self._workspace = None
                self._workspace = torch.empty(need, dtype=torch.uint8, device=dev)
            return self._launch(tok, lens, self._workspace, B, L, prec)

    def last_launch_count(self) -> int:
        return int(_lib.load().st2_decoder_last_launch_count(self._handle)) if self._handle else 0

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().st2_decoder_destroy(self._handle)
                self._handle = None
        except Exception:
            pass
