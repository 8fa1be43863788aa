#!/bin/bash
# pointwise convs over the flattened batch: full GPU suite + vocos profile + chain bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2v_suite.log 2>&1; echo "gpu suite rc=$?"; tail -4 gpurun_out/r2v_suite.log
timeout 300 python tools/profile_layers.py --variant vocos > gpurun_out/layers_r2v_vocos.txt 2>&1; head -8 gpurun_out/layers_r2v_vocos.txt
timeout 300 python tools/profile_layers.py 2>&1 | grep -E "^total|conv_tc" | head -14
