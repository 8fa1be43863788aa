#!/bin/bash
# round-2 GPU pass A: unit parity of conv_row, full GPU suite, per-layer A/B against conv_pipe, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
echo "== row unit tests"
timeout 600 python -m pytest tests/test_gpu_fused.py -x -q -k "row" > gpurun_out/r2a_row_tests.log 2>&1; rc=$?
tail -15 gpurun_out/r2a_row_tests.log
echo "row tests rc=$rc"
if [ $rc -ne 0 ]; then
  echo "== row tests failed: run them all (no -x) for the failure pattern"
  timeout 900 python -m pytest tests/test_gpu_fused.py -q -k "row" > gpurun_out/r2a_row_tests_all.log 2>&1
  tail -40 gpurun_out/r2a_row_tests_all.log
fi
echo "== full gpu suite (ST2_NO_ROW=${ST2_NO_ROW:-unset})"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gpu_tests.log 2>&1; echo "suite rc=$?"
tail -8 gpurun_out/r2a_gpu_tests.log
echo "== per-layer profiles"
timeout 300 python tools/profile_layers.py > gpurun_out/layers_r2a_row.txt 2> gpurun_out/layers_r2a_row.err; head -3 gpurun_out/layers_r2a_row.txt
ST2_NO_ROW=1 timeout 300 python tools/profile_layers.py > gpurun_out/layers_r2a_pipe.txt 2> gpurun_out/layers_r2a_pipe.err; head -3 gpurun_out/layers_r2a_pipe.txt
echo "== bench"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 1500 gpurun_out/bench_r2a.json
