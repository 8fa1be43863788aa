#!/bin/bash
# what do the AdaIN coefficient launches really cost inside an unprofiled forward?  (skip them: wrong results, valid timing)
mkdir -p gpurun_out
for cfg in "ST2_X=0" "ST2_SKIP_COEF=1" "ST2_X=0" "ST2_SKIP_COEF=1"; do
  echo "== $cfg"; env $cfg timeout 300 python bench.py --no-aux --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['gpu_launches']//10, [round(x,2) for x in d['per_step_ms']])"
done
