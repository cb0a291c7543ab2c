#!/bin/bash
# i8 engine bring-up: targeted tests with short timeouts first (a hung kernel must not eat the box)
mkdir -p gpurun_out
TAG=${1:-i8a}
timeout 300 python -m pytest tests/test_gpu_i8.py -x -q --timeout 120 -k "packed_image or integer_cov" > gpurun_out/i8_stage1_$TAG.log 2>&1
echo "stage1 exit $?" >> gpurun_out/i8_stage1_$TAG.log
tail -30 gpurun_out/i8_stage1_$TAG.log
timeout 600 python -m pytest tests/test_gpu_i8.py -q --timeout 200 > gpurun_out/i8_stage2_$TAG.log 2>&1
echo "stage2 exit $?" >> gpurun_out/i8_stage2_$TAG.log
tail -30 gpurun_out/i8_stage2_$TAG.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q --timeout 300 > gpurun_out/parity_$TAG.log 2>&1
echo "parity exit $?" >> gpurun_out/parity_$TAG.log
tail -5 gpurun_out/parity_$TAG.log
B200_PAIR_ENGINE=2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_i8_$TAG.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_i8_$TAG.log
tail -3 gpurun_out/bench_i8_$TAG.log
