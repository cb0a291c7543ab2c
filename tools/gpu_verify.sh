#!/bin/bash
# full GPU test-suite + N=1 bench.  usage: tools/gpu_verify.sh TAG [bench args]
TAG=${1:-v}; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -6 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 "$@" > gpurun_out/bench_$TAG.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_$TAG.log
tail -3 gpurun_out/bench_$TAG.log
