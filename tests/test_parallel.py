"""CPU tests of the N>1 host logic (gloo, world_size 2): sharding and the gather to rank 0."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from styletts2_lite_b200.parallel import gather_waveforms, shard_range


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


def _worker(rank, world, port, ragged, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, S = (5 if ragged else 6), 48
        a, b = shard_range(n, rank, world)
        counts = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
        full = torch.arange(n * S, dtype=torch.float32).reshape(n, 1, S)
        out = gather_waveforms(full[a:b].clone(), counts)
        if rank == 0:
            q.put(bool(torch.equal(out, full)))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ragged", [False, True])
def test_gather_to_rank0_gloo(ragged):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29711 + (1 if ragged else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ragged, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
