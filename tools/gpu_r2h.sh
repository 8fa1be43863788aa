#!/bin/bash
# A/B of conv_row build variants on one box: libst2_s{sections}p0.so
mkdir -p gpurun_out
for v in s0p0 s1p0 s0p0 s1p0; do
  ST2_B200_LIB=$PWD/styletts2_lite_b200/lib/libst2_$v.so timeout 300 python tools/profile_layers.py > gpurun_out/sweep_r2h_$v.txt 2>&1 || { echo "$v FAILED"; tail -3 gpurun_out/sweep_r2h_$v.txt; }
  echo "$v: $(head -1 gpurun_out/sweep_r2h_$v.txt)"
done
