#!/bin/bash
# postprocess tests + ncu captures of conv_row after the self-fed ring change
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_postprocess.py -q > gpurun_out/r2c_post.log 2>&1; echo "post rc=$?"; tail -3 gpurun_out/r2c_post.log
python tools/one_forward.py > gpurun_out/r2c_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2c_plain.log; exit 1; }
tail -1 gpurun_out/r2c_plain.log
for spec in "c64k3:140" "c64k11:152" "c32k3:163" "c32k11:175"; do
  name=${spec%%:*}; skip=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:conv_row_kernel -s $skip -c 2 -f -o gpurun_out/prof_r2c_$name \
      python tools/one_forward.py > gpurun_out/r2c_ncu_$name.log 2>&1
  echo "$name rc=$?"
done
ls -la gpurun_out/prof_r2c_*
