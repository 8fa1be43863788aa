"""Host-fed serving loop for the drop-in decoder: batches come from pinned host memory and the waveforms go back to
pinned host memory, with the copies of neighbouring batches overlapped with the decoder on separate CUDA streams.

The reference moves one sentence at a time (`.to(device)` ... `.cpu().numpy()`, inference.py:235-272); a server decoding
a queue of batches pays the host<->device copies on the critical path unless they are pipelined:

    copy-in stream   H2D(i+1)            H2D(i+2)
    compute stream            forward(i)           forward(i+1)
    copy-out stream                      D2H(i-1)            D2H(i)

Two device input sets and two pinned output buffers are enough.  Nothing here touches the arithmetic: `decoder` is a
`B200Decoder` and every forward is the same C-ABI call the module makes.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Iterator, Optional

import torch

_KEYS = ("asr", "F0_curve", "N", "s")


class PipelinedDecoder:
    """decode(batches) yields one pinned-host waveform tensor [B,1,600*T] per host batch, in order.

    `batches`: iterable of dicts with pinned (or pageable) CPU tensors asr [B,512,T], F0_curve [B,2T], N [B,2T], s [B,128];
    shapes may change between batches.  A yielded tensor is a reused pinned buffer: it stays valid until the generator is
    advanced again (copy it if it must outlive that)."""

    def __init__(self, decoder, device: Optional[torch.device] = None, precision: Optional[str] = None,
                 after_forward: Optional[Callable[[torch.Tensor], None]] = None):
        self.decoder = decoder
        self.device = device or next(decoder.parameters()).device
        self.precision = precision
        self.after_forward = after_forward          # e.g. the NCCL gather of a multi-GPU job, on the compute stream
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._dev_in = [None, None]
        self._host_out = [None, None]
        self._in_free = [None, None]                # event: the forward that read input set k has finished
        self._out_done = [None, None]               # event: the D2H into host buffer k has finished

    def _stage_in(self, k: int, batch: Dict[str, torch.Tensor]) -> torch.cuda.Event:
        with torch.cuda.stream(self.s_in):
            if self._in_free[k] is not None:
                self.s_in.wait_event(self._in_free[k])
            cur = self._dev_in[k]
            if cur is None or any(cur[n].shape != batch[n].shape for n in _KEYS):
                cur = {n: torch.empty(batch[n].shape, dtype=torch.float32, device=self.device) for n in _KEYS}
                self._dev_in[k] = cur
            for n in _KEYS:
                cur[n].copy_(batch[n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.s_in)
        return ev

    def decode(self, batches: Iterable[Dict[str, torch.Tensor]], seeds: Optional[Iterator[int]] = None):
        compute = torch.cuda.current_stream(self.device)
        it = iter(batches)
        nxt = next(it, None)
        if nxt is None:
            return
        ready = self._stage_in(0, nxt)
        i = 0
        pending = []                                # (host buffer index, event) of D2H copies not yet yielded
        while nxt is not None:
            k = i & 1
            following = next(it, None)
            ready_next = self._stage_in(k ^ 1, following) if following is not None else None   # H2D(i+1) overlaps forward(i)
            compute.wait_event(ready)
            d = self._dev_in[k]
            with torch.no_grad():
                out = self.decoder(d["asr"], d["F0_curve"], d["N"], d["s"], precision=self.precision,
                                   seed=(next(seeds) if seeds is not None else None))
            if self.after_forward is not None:
                self.after_forward(out)
            done = torch.cuda.Event()
            done.record(compute)
            self._in_free[k] = done
            with torch.cuda.stream(self.s_out):     # D2H(i) overlaps forward(i+1)
                self.s_out.wait_event(done)
                if self._out_done[k] is not None:
                    self._out_done[k].synchronize() # the consumer has had two batches to read host buffer k
                hb = self._host_out[k]
                if hb is None or hb.shape != out.shape:
                    hb = torch.empty(out.shape, dtype=torch.float32).pin_memory()
                    self._host_out[k] = hb
                hb.copy_(out, non_blocking=True)
                out.record_stream(self.s_out)
                ev = torch.cuda.Event()
                ev.record(self.s_out)
                self._out_done[k] = ev
            pending.append((k, ev))
            if len(pending) == 2:                   # hand out batch i-1 while batch i is still computing
                kk, e = pending.pop(0)
                e.synchronize()
                yield self._host_out[kk]
            nxt, ready, i = following, ready_next, i + 1
        for kk, e in pending:
            e.synchronize()
            yield self._host_out[kk]
