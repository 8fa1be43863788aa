#!/bin/bash
# final profiles of the round-2 build: GPU suite, ncu launch list of the bench command, one full capture of conv_row launches, bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3g_suite.log 2>&1; echo "gpu suite rc=$?"; tail -3 gpurun_out/r3g_suite.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux"
$CMD > gpurun_out/r3g_bench_plain.json 2> gpurun_out/r3g_bench_plain.err || { echo "plain bench failed"; tail -5 gpurun_out/r3g_bench_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1200 -c 520 --csv \
    --log-file gpurun_out/r3g_launches.csv $CMD > gpurun_out/r3g_ncu_bench.log 2>&1
echo "launch list rc=$?"
python tools/one_forward.py > gpurun_out/r3g_plain.log 2>&1 || { echo "plain forward failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:conv_row_kernel -s 140 -c 12 -f -o gpurun_out/prof_r3g_row_c32 \
    python tools/one_forward.py > gpurun_out/r3g_ncu_c32.log 2>&1; echo "c32 rc=$?"
python bench.py > gpurun_out/bench_r3g.json 2> gpurun_out/bench_r3g.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r3g_ref.json 2> gpurun_out/bench_r3g_ref.err; echo "reference arm rc=$?"
tail -c 600 gpurun_out/bench_r3g_ref.json
ls -la gpurun_out/prof_r3g_* gpurun_out/r3g_launches.csv
