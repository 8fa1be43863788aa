#!/bin/bash
mkdir -p gpurun_out
python tools/one_forward.py > gpurun_out/r2g_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2g_plain.log; exit 1; }
for spec in "c64k3:140" "c64k7:146" "c32k3:163" "c32k7:169"; do
  name=${spec%%:*}; skip=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:conv_row_kernel -s $skip -c 2 -f -o gpurun_out/prof_r2g_$name \
      python tools/one_forward.py > gpurun_out/r2g_ncu_$name.log 2>&1
  echo "$name rc=$?"
done
