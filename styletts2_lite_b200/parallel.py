"""Utterance sharding across the GPUs of one box.

Every op of the decoder is per utterance, so ranks never exchange activations: rank r of W decodes utterances
[r*B/W, (r+1)*B/W) with replicated weights, and the only collective is the gather of the waveforms to rank 0 (NCCL over
NVLink on the GPU box, gloo in the CPU tests).

`ShardedGather` is the production path: rank 0 owns ONE preallocated [n_total, 1, S] buffer, every micro-batch a rank decodes is
sent straight into its slice of that buffer (point-to-point, no staging copies, no concatenation, no per-step allocation), and
the transfers are issued on a side stream behind an event of the compute stream, so the gather of micro-batch k overlaps the
forward of micro-batch k + 1.  `gather_waveforms` is the one-shot convenience form on top of the same buffer logic.

`direct=True` (CUDA, one box): the buffer is symmetric memory (torch.distributed._symmetric_memory: every rank maps rank 0's
allocation over NVLink / NVSwitch) and `target(j)` is the slice of it micro-batch j belongs to -- pass it as
`decoder(..., out=g.target(j))` and the decoder's last kernel stores the waveform straight into rank 0's HBM: the gather is
fused into the compute, there is no send / receive, no side stream and nothing to overlap.  A micro-batch that was not written
in place is copied there device-to-device by `submit`.  Falls back to the point-to-point path when symmetric memory cannot be
set up (CPU tensors, gloo, no peer access).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_utterances: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) slice of the utterance list for `rank` (first ranks get the remainder)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world %d/%d" % (rank, world))
    base, rem = divmod(n_utterances, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _dist_on() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


class ShardedGather:
    """Gather of a job of `n_total` utterances, decoded in micro-batches of at most `micro_batch` per rank, to rank `dst`.

        g = ShardedGather(n_total, S, micro_batch, device)
        for j, (lo, hi) in enumerate(g.my_micro_batches()):      # utterance indices of this rank, global numbering
            wave = decoder(...inputs of utterances lo..hi...)    # [hi - lo, 1, S]
            g.submit(j, wave)                                    # asynchronous: returns at once
        full = g.finish()                                        # [n_total, 1, S] on dst (None elsewhere), all transfers done

    The buffer on `dst` is allocated once and reused by every job of the same shape (`reset()` starts the next one)."""

    def __init__(self, n_total: int, samples: int, micro_batch: int, device: torch.device, dst: int = 0,
                 dtype: torch.dtype = torch.float32, direct: bool = False):
        self.n_total, self.S, self.mb, self.dst = int(n_total), int(samples), max(1, int(micro_batch)), dst
        self.device = torch.device(device)
        self.on = _dist_on()
        self.world = dist.get_world_size() if self.on else 1
        self.rank = dist.get_rank() if self.on else 0
        self.spans = [shard_range(self.n_total, r, self.world) for r in range(self.world)]
        self.direct = False
        self._hdl = None
        self._peer = None                      # dst's buffer as this rank sees it (direct mode)
        if direct and self.on and self.device.type == "cuda":
            self._setup_direct(dtype)
        if self.direct:
            self.full = self._peer if self.rank == dst else None
        else:
            self.full = (torch.empty((self.n_total, 1, self.S), dtype=dtype, device=self.device) if self.rank == dst else None)
        self.side = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self._works: List = []
        self._keep: List[torch.Tensor] = []

    def _setup_direct(self, dtype: torch.dtype) -> None:
        """Symmetric allocation of the gathered buffer (every rank allocates, only dst's copy is used) and the rendezvous that
        maps dst's copy into this process.  All ranks must agree, so the outcome is reduced over the group."""
        ok = 1
        try:
            import torch.distributed._symmetric_memory as symm
            with torch.cuda.device(self.device):
                buf = symm.empty((self.n_total, 1, self.S), dtype=dtype, device=self.device)
                hdl = symm.rendezvous(buf, dist.group.WORLD)
                peer = buf if self.rank == self.dst else hdl.get_buffer(self.dst, (self.n_total, 1, self.S), dtype)
        except Exception:  # noqa: BLE001  (no symmetric memory in this build / no peer access: use send / recv)
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            self._sym, self._hdl, self._peer, self.direct = buf, hdl, peer, True

    def target(self, j: int) -> Optional[torch.Tensor]:
        """Direct mode: the [hi - lo, 1, S] slice of dst's buffer that micro-batch j of THIS rank belongs to (peer memory on
        every rank but dst) -- the `out=` of the decoder call.  None otherwise."""
        if not self.direct:
            return None
        lo, hi = self.my_micro_batches()[j]
        return self._peer[lo:hi]

    def micro_batches_of(self, rank: int) -> List[Tuple[int, int]]:
        a, b = self.spans[rank]
        return [(lo, min(lo + self.mb, b)) for lo in range(a, b, self.mb)]

    def my_micro_batches(self) -> List[Tuple[int, int]]:
        return self.micro_batches_of(self.rank)

    def reset(self) -> None:
        self.finish()

    def submit(self, j: int, wave: torch.Tensor) -> None:
        """Micro-batch j of this rank is ready on the current stream: move it (and, on dst, everybody's j-th micro-batch)."""
        mine = self.my_micro_batches()
        lo, hi = mine[j]
        if tuple(wave.shape) != (hi - lo, 1, self.S):
            raise ValueError("micro-batch %d: expected %s, got %s" % (j, (hi - lo, 1, self.S), tuple(wave.shape)))
        wave = wave.contiguous()
        if self.direct:
            dst_view = self._peer[lo:hi]
            if wave.data_ptr() != dst_view.data_ptr():          # not decoded in place: one device-to-device copy over NVLink
                dst_view.copy_(wave, non_blocking=True)
            return

        def issue():
            if self.rank == self.dst:
                self.full[lo:hi].copy_(wave, non_blocking=True)
                if not self.on:
                    return
                ops = []
                for r in range(self.world):
                    if r == self.dst:
                        continue
                    theirs = self.micro_batches_of(r)
                    if j < len(theirs):
                        ops.append(dist.P2POp(dist.irecv, self.full[theirs[j][0]:theirs[j][1]], r))
                if ops:
                    self._works.extend(dist.batch_isend_irecv(ops))
            elif self.on:
                self._works.extend(dist.batch_isend_irecv([dist.P2POp(dist.isend, wave, self.dst)]))
                self._keep.append(wave)                     # alive until finish(): the send reads it asynchronously

        # ranks with fewer micro-batches than dst's j simply have nothing to send; dst posts only the receives that exist
        if self.side is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.side):
                self.side.wait_event(ev)
                issue()
                wave.record_stream(self.side)
        else:
            issue()

    def drain_remote(self) -> None:
        """dst only: post the receives of micro-batches that other ranks have beyond dst's own count (ragged shards)."""
        if self.direct or not (self.on and self.rank == self.dst):
            return
        mine = len(self.my_micro_batches())
        ops = []
        for r in range(self.world):
            if r == self.dst:
                continue
            for lo, hi in self.micro_batches_of(r)[mine:]:
                ops.append(dist.P2POp(dist.irecv, self.full[lo:hi], r))
        if ops:
            ctx = torch.cuda.stream(self.side) if self.side is not None else _Null()
            with ctx:
                self._works.extend(dist.batch_isend_irecv(ops))

    def finish(self) -> Optional[torch.Tensor]:
        """Wait for every transfer issued so far (the current stream waits on the side stream); returns the buffer on dst."""
        if self.direct:
            # every rank's stores so far are ordered before the barrier on its stream; after it dst may read the buffer
            self._hdl.barrier()
            return self.full
        self.drain_remote()
        for w in self._works:
            w.wait()
        self._works.clear()
        if self.side is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.side)
        self._keep.clear()
        return self.full


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_ONE_SHOT: Dict[tuple, ShardedGather] = {}


def gather_waveforms(wave: torch.Tensor, counts: Optional[List[int]] = None, dst: int = 0) -> Optional[torch.Tensor]:
    """Gather per-rank waveforms [b_r,1,S] to `dst`; returns the [sum b_r,1,S] buffer there (a cached buffer that the next
    call with the same shapes overwrites), None elsewhere.  `counts` = utterances per rank (equal shards by default; with
    ragged shards they must follow shard_range)."""
    if not _dist_on():
        return wave
    world = dist.get_world_size()
    counts = counts or [wave.shape[0]] * world
    n_total, S = int(sum(counts)), int(wave.shape[-1])
    key = (n_total, S, tuple(counts), wave.device, wave.dtype, dst)
    g = _ONE_SHOT.get(key)
    if g is None:
        g = ShardedGather(n_total, S, max(counts), wave.device, dst=dst, dtype=wave.dtype)
        if [b - a for a, b in g.spans] != list(counts):
            raise ValueError("counts %s do not follow shard_range (%s)" % (counts, [b - a for a, b in g.spans]))
        _ONE_SHOT[key] = g
    if wave.shape[0] > 0:
        g.submit(0, wave)
    return g.finish()
