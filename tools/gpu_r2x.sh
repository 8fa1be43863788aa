#!/bin/bash
# vocos: GELU fused into the first pointwise GEMM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py -q -x -k "vocos" > gpurun_out/r2x_vocos.log 2>&1; echo "vocos tests rc=$?"; tail -4 gpurun_out/r2x_vocos.log
timeout 300 python tools/profile_layers.py --variant vocos > gpurun_out/layers_r2x_vocos.txt 2>&1; head -12 gpurun_out/layers_r2x_vocos.txt
