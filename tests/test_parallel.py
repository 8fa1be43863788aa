"""CPU tests of the N>1 host logic (gloo, world_size 2): sharding, the one-shot gather and the micro-batched ShardedGather
(every micro-batch lands in its slice of ONE preallocated buffer on rank 0)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from styletts2_lite_b200.parallel import ShardedGather, gather_waveforms, shard_range


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 64, 1024, 1025):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def test_sharded_gather_single_process_is_a_copy():
    g = ShardedGather(5, 12, 2, torch.device("cpu"))
    full = torch.arange(5 * 12, dtype=torch.float32).reshape(5, 1, 12)
    assert g.my_micro_batches() == [(0, 2), (2, 4), (4, 5)]
    for j, (lo, hi) in enumerate(g.my_micro_batches()):
        g.submit(j, full[lo:hi].clone())
    assert torch.equal(g.finish(), full)
    with pytest.raises(ValueError):
        g.submit(0, full[0:1])


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        S = 48
        if mode in ("equal", "ragged"):
            n = 5 if mode == "ragged" else 6
            a, b = shard_range(n, rank, world)
            counts = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            full = torch.arange(n * S, dtype=torch.float32).reshape(n, 1, S)
            out = gather_waveforms(full[a:b].clone(), counts)
            again = gather_waveforms(full[a:b].clone() + 1.0, counts)        # the cached buffer is reused, not re-allocated
            if rank == 0:
                q.put(bool(torch.equal(again, full + 1.0)) and out.data_ptr() == again.data_ptr())
            else:
                assert out is None and again is None
        else:
            # micro-batched job: 11 utterances, micro-batches of 2 -> rank 0 has 3 micro-batches (6 utterances), rank 1 has 3 (5)
            n, mb = (11, 2) if mode == "micro" else (9, 4)                   # "micro_uneven": 5 + 4 utterances, 2 vs 1 micro-batches
            full = torch.arange(n * S, dtype=torch.float32).reshape(n, 1, S)
            g = ShardedGather(n, S, mb, torch.device("cpu"))
            for rep in range(2):                                             # the same object serves consecutive jobs
                for j, (lo, hi) in enumerate(g.my_micro_batches()):
                    g.submit(j, full[lo:hi].clone() + rep)
                out = g.finish()
                if rank == 0:
                    ok = bool(torch.equal(out, full + rep))
                    if rep == 1:
                        q.put(ok)
                    else:
                        assert ok
                else:
                    assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["equal", "ragged", "micro", "micro_uneven"])
def test_gather_to_rank0_gloo(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29711 + ["equal", "ragged", "micro", "micro_uneven"].index(mode)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
