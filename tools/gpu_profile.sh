#!/bin/bash
# ncu launch list + full capture of the pair kernel (1 GPU). Outputs under gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 1 -c 2 \
    -o gpurun_out/prof_pair_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
