#!/bin/bash
# round-2 final evidence on one GPU: tests, default bench, launch list, ncu captures of the kernels added / changed
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu_r2y.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2y.log
tail -4 gpurun_out/pytest_gpu_r2y.log
(time python bench.py) > gpurun_out/bench_r2y_n1.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_r2y_n1.log
tail -c 600 gpurun_out/bench_r2y_n1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_r2b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_onevn_r2b.csv python tools/onevn_check.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:onevn_stream2_kernel -s 2 -c 1 -o gpurun_out/prof_onevn2_r2b -f python tools/onevn_check.py > gpurun_out/ncu_onevn2_r2b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hieragglo_kernel -s 1 -c 1 -o gpurun_out/prof_hieragglo_r2b -f python tools/hieragglo_time.py 10000 > gpurun_out/ncu_hieragglo_r2b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:avgcorr_kernel -s 1 -c 1 -o gpurun_out/prof_avgcorr_r2b -f python tools/rmsavgcorr_time.py 3000 1000 > gpurun_out/ncu_avgcorr_r2b.log 2>&1
ls -la gpurun_out/*r2b*
