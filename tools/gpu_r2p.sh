#!/bin/bash
# profiles of the round-2 build: ncu launch list of the bench command + full captures of conv_row launches
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux"
$CMD > gpurun_out/r2p_bench_plain.json 2> gpurun_out/r2p_bench_plain.err || { echo "plain bench failed"; tail -5 gpurun_out/r2p_bench_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1400 -c 600 --csv \
    --log-file gpurun_out/r2p_launches.csv $CMD > gpurun_out/r2p_ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/one_forward.py > gpurun_out/r2p_plain.log 2>&1 || { echo "plain forward failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv_row_kernel -s 140 -c 12 -f -o gpurun_out/prof_r2p_row_c64 \
    python tools/one_forward.py > gpurun_out/r2p_ncu_c64.log 2>&1; echo "c64 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:conv_row_kernel -s 163 -c 12 -f -o gpurun_out/prof_r2p_row_c32 \
    python tools/one_forward.py > gpurun_out/r2p_ncu_c32.log 2>&1; echo "c32 rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2p.json 2> gpurun_out/bench_r2p.err; echo "bench rc=$?"
ls -la gpurun_out/prof_r2p_* gpurun_out/r2p_launches.csv
