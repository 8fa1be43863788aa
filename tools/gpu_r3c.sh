#!/bin/bash
# direct (symmetric-memory) gather on 2 GPUs: correctness, then the bench both ways
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_direct_gather.py 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*" | tail -8
for mode in direct p2p; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --gather $mode > gpurun_out/bench_r3c_n2_$mode.json 2> gpurun_out/bench_r3c_n2_$mode.err; echo "bench $mode rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r3c_n2_$mode.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], d['config'].get('gather_used'))
c=d.get('cfg4_sharded_1024x10s'); print(c if c is None or 'error' in c else {k:c[k] for k in ('ms','audio_s_per_s','exposed_gather_tail_ms','gather','finite')})
PY
  tail -3 gpurun_out/bench_r3c_n2_$mode.err | cut -c1-300
done
