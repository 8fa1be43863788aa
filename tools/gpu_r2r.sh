#!/bin/bash
# Vocos decoder variant: GPU parity tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decoder.py -q -x -k "vocos" > gpurun_out/r2r_vocos.log 2>&1; echo "vocos tests rc=$?"; tail -30 gpurun_out/r2r_vocos.log
