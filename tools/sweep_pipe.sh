#!/bin/bash
# usage: tools/sweep_pipe.sh tag "ENV1=.. ENV2=.." "ENV.." ...   -> gpurun_out/sweep_<tag>_<i>.txt (per-layer profile per setting)
tag=$1; shift
i=0
for cfg in "$@"; do
  echo "== cfg $i: $cfg"
  env $cfg python tools/profile_layers.py > gpurun_out/sweep_${tag}_$i.txt 2>&1 || { echo "FAILED"; tail -3 gpurun_out/sweep_${tag}_$i.txt; }
  head -1 gpurun_out/sweep_${tag}_$i.txt
  i=$((i+1))
done
