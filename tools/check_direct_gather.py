"""torchrun check of parallel.ShardedGather(direct=True) on the GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_direct_gather.py

Every rank decodes its shard of a small job three ways -- into rank 0's buffer directly (`out=g.target(j)`), into a local tensor that
`submit` then copies to the peer buffer, and through the NCCL send / recv path -- and rank 0 checks that the three gathered
buffers are bit-identical and finite."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from styletts2_lite_b200 import synth  # noqa: E402
from styletts2_lite_b200.config import DecoderConfig  # noqa: E402
from styletts2_lite_b200.decoder import B200Decoder  # noqa: E402
from styletts2_lite_b200.parallel import ShardedGather  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = DecoderConfig.hifigan()
    m = B200Decoder(cfg, "bf16")
    m.load_state_dict(synth.make_state_dict(cfg, 0, True))
    m = m.to(dev).eval()
    N_UTT, T, MB = 5 * world + 1, 20, 2                      # ragged shards, uneven micro-batch counts
    S = 600 * T
    inp = {k: v.to(dev) for k, v in synth.make_inputs(N_UTT, T, seed=7, cfg=cfg, with_noise=False).items()}
    results = {}
    for mode in ("direct_out", "direct_copy", "p2p"):
        g = ShardedGather(N_UTT, S, MB, dev, direct=(mode != "p2p"))
        if mode != "p2p" and not g.direct:
            if rank == 0:
                print("symmetric memory not available: direct mode fell back to p2p")
        for rep in range(2):                                 # second pass reuses the buffer
            for j, (lo, hi) in enumerate(g.my_micro_batches()):
                with torch.no_grad():
                    out = m(inp["asr"][lo:hi], inp["F0_curve"][lo:hi], inp["N"][lo:hi], inp["s"][lo:hi], seed=100 + lo,
                            out=(g.target(j) if mode == "direct_out" else None))
                g.submit(j, out)
            full = g.finish()
        torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            results[mode] = full.clone()
        dist.barrier()
    if rank == 0:
        a, b, c = results["direct_out"], results["direct_copy"], results["p2p"]
        ok = bool(torch.isfinite(a).all()) and torch.equal(a, b) and torch.equal(a, c) and float(a.abs().max()) > 0
        print("direct gather check:", "OK" if ok else "MISMATCH", tuple(a.shape), float(a.abs().max()))
        if not ok:
            sys.exit(1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
