#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_predictor.py tests/test_gpu_decoder.py -q -x -k "synthesizer or chained or other_head or inline" > gpurun_out/r3e.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r3e.log
