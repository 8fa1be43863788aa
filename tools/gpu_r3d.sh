#!/bin/bash
# A/B of the gather modes at 2 GPUs, order swapped, headline only
mkdir -p gpurun_out
for mode in p2p direct p2p direct; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 --no-aux --gather $mode 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), [round(x,1) for x in d['per_step_ms']])"
done
