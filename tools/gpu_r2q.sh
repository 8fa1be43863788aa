#!/bin/bash
# full GPU suite (ragged token batches, duration smoothing included)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2q_suite.log 2>&1; echo "gpu suite rc=$?"; tail -6 gpurun_out/r2q_suite.log
