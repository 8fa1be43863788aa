#!/bin/bash
# Vocos per-layer profile + the default bench line (with the vocos auxiliary)
mkdir -p gpurun_out
timeout 300 python tools/profile_layers.py --variant vocos > gpurun_out/layers_r2s_vocos.txt 2>&1; echo "profile rc=$?"; head -24 gpurun_out/layers_r2s_vocos.txt
timeout 900 python bench.py > gpurun_out/bench_r2s.json 2> gpurun_out/bench_r2s.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_r2s.json
