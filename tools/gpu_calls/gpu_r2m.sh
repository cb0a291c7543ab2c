#!/bin/bash
mkdir -p gpurun_out
for v in default o_s7 o_u2 o_ic o_all; do
  if [ $v == default ]; then unset B200_RMSD_LIB; else export B200_RMSD_LIB=$PWD/variants/$v.so; fi
  echo "== $v"; timeout 300 python tools/onevn_check.py 2>&1 | tail -2
done
unset B200_RMSD_LIB
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_i8_kernel -s 1 -c 1 -o gpurun_out/prof_pair_r2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_pair_r2.log 2>&1
tail -2 gpurun_out/ncu_pair_r2.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onevn_stream2_kernel -s 2 -c 1 -o gpurun_out/prof_onevn2_r2 -f python tools/onevn_check.py > gpurun_out/ncu_onevn2_r2.log 2>&1
tail -2 gpurun_out/ncu_onevn2_r2.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_r2.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r2.csv
