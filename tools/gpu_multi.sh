#!/bin/bash
# usage: gpu_multi.sh N   -- GPU tests (1 GPU) then the N-rank bench, as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_n$N.log
tail -3 gpurun_out/bench_n$N.log
