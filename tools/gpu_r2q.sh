#!/bin/bash
# ragged token batches: predictor / text-encoder GPU tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_predictor.py -q > gpurun_out/r2q_pred.log 2>&1; echo "predictor tests rc=$?"; tail -5 gpurun_out/r2q_pred.log
