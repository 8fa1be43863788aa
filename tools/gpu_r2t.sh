#!/bin/bash
# fp16 noise source (bias-free) / last-stage output: decoder tests, then per-layer A/B and bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_decoder.py -q > gpurun_out/r2t_dec.log 2>&1; echo "decoder tests rc=$?"; tail -8 gpurun_out/r2t_dec.log
for cfg in "ST2_NO_SRC16=1 ST2_NO_OUT16=1" "ST2_X=0" "ST2_NO_SRC16=1 ST2_NO_OUT16=1" "ST2_X=0"; do
  echo "== $cfg"; env $cfg timeout 300 python tools/profile_layers.py 2>&1 | grep -E "^total" | head -3
done
timeout 600 python bench.py --no-aux --no-cpu-baseline > gpurun_out/bench_r2t.json 2> gpurun_out/bench_r2t.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2t.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'])
PY
