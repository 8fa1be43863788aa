#!/bin/bash
# ncu --set full captures of representative conv_row launches (4th forward of tools/one_forward.py; 45 conv_row launches per forward)
mkdir -p gpurun_out
python tools/one_forward.py > gpurun_out/r2b_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2b_plain.log; exit 1; }
cat gpurun_out/r2b_plain.log | tail -1
for spec in "c64k3:140" "c64k11:152" "c32k3:163" "c32k11:175"; do
  name=${spec%%:*}; skip=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:conv_row_kernel -s $skip -c 2 -f -o gpurun_out/prof_r2b_$name \
      python tools/one_forward.py > gpurun_out/r2b_ncu_$name.log 2>&1
  echo "$name rc=$?"; tail -2 gpurun_out/r2b_ncu_$name.log
done
ls -la gpurun_out/prof_r2b_*
