#!/bin/bash
# conv_row: bias through the tensor core (C = 32, scale 1): parity, then A/B per layer and whole step
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fused.py tests/test_gpu_decoder.py -q -x > gpurun_out/r3f_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r3f_tests.log
for cfg in "ST2_NO_ROW_BIAS_MMA=1" "ST2_X=0" "ST2_NO_ROW_BIAS_MMA=1" "ST2_X=0"; do
  echo "== $cfg"; env $cfg timeout 300 python tools/profile_layers.py 2>&1 | grep -E "^total|^conv_row" | head -12
done
