#!/bin/bash
# noise_conv tile size: decoder tests + source rows of the per-layer profile (twice: box noise)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_units.py -q -x > gpurun_out/r2w_dec.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2w_dec.log
for i in 1 2; do timeout 300 python tools/profile_layers.py 2>&1 | grep -E "^total|^source|^norm_coef" | head -8; done
