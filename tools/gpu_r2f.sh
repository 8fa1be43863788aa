#!/bin/bash
# GPU tests of the last changes + ncu capture of 128-channel conv_pipe launches (stage 1 of the 4th forward)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r2f_tests.log
python tools/one_forward.py > gpurun_out/r2f_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2f_plain.log; exit 1; }
# conv_pipe launches per forward: 43; stage 0: noise_res (N-pass / fused), ups...  capture a window in stage 1
ncu --set full --clock-control none --import-source on -k regex:conv_pipe_kernel -s 141 -c 12 -f -o gpurun_out/prof_r2f_pipe128 \
    python tools/one_forward.py > gpurun_out/r2f_ncu.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out/prof_r2f_*
