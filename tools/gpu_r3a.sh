#!/bin/bash
# conv_row computes its AdaIN coefficients itself: GPU suite + A/B of the bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r3a_suite.log 2>&1; echo "gpu suite rc=$?"; tail -4 gpurun_out/r3a_suite.log
for cfg in "ST2_NO_ROW_INLINE_COEF=1" "ST2_X=0" "ST2_NO_ROW_INLINE_COEF=1" "ST2_X=0"; do
  echo "== $cfg"; env $cfg timeout 300 python bench.py --no-aux --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['gpu_launches']//10, d['roofline']['frac'], [round(x,2) for x in d['per_step_ms']])"
done
