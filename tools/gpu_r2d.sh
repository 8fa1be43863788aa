#!/bin/bash
mkdir -p gpurun_out
i=0
for cfg in "ST2_X=0" "ST2_ROW_SUB=2 ST2_ROW_NA=4" "ST2_ROW_SUB=2 ST2_ROW_NA=3" "ST2_ROW_SUB=1 ST2_ROW_NA=4" "ST2_X=0"; do
  echo "== cfg $i: $cfg"
  env $cfg timeout 300 python tools/profile_layers.py > gpurun_out/sweep_r2d_$i.txt 2>&1 || { echo FAILED; tail -3 gpurun_out/sweep_r2d_$i.txt; }
  head -1 gpurun_out/sweep_r2d_$i.txt
  i=$((i+1))
done
