#!/bin/bash
# conv_tc ring depth / N tile experiments (vocos + hifigan front half)
mkdir -p gpurun_out
for cfg in "ST2_X=0" "ST2_TC_SMEM_KB=192" "ST2_TC_SMEM_KB=144" "ST2_TC_BN=128" "ST2_TC_BN=128 ST2_TC_SMEM_KB=192" "ST2_X=0"; do
  echo "== $cfg"
  env $cfg timeout 300 python tools/profile_layers.py --variant vocos 2>&1 | grep -E "^total|^conv_tc" | head -5
  env $cfg timeout 300 python tools/profile_layers.py 2>&1 | grep -E "^total|^conv_tc" | head -4
done
