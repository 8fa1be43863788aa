#!/bin/bash
# 8-GPU bench (weak scaling line + cfg4 auxiliary), launched the way the driver launches it
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r2u_n$N.json 2> gpurun_out/bench_r2u_n$N.err; echo "bench N=$N rc=$?"
tail -c 2500 gpurun_out/bench_r2u_n$N.json; tail -3 gpurun_out/bench_r2u_n$N.err
