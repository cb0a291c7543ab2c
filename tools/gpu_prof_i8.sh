#!/bin/bash
# bench (N=1) + ncu launch list + full capture of the tcgen05 pair kernel.  usage: tools/gpu_prof_i8.sh TAG [--tests]
TAG=${1:-i8b}
mkdir -p gpurun_out
if [ "$2" == "--tests" ]; then
  timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
  tail -4 gpurun_out/pytest_gpu_$TAG.log
fi
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_$TAG.log
tail -2 gpurun_out/bench_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_i8_kernel -s 1 -c 1 \
    -o gpurun_out/prof_pair_i8_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
