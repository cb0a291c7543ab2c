#!/bin/bash
# ncu evidence for the final tree: launch list of the default bench command, one full capture of the pair kernel
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_r4n.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_r4n.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_i8_kernel -s 1 -c 1 \
    -o gpurun_out/prof_pair_r4n -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_r4n.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out/prof_pair_r4n.ncu-rep gpurun_out/launches_r4n.csv
