#!/bin/bash
# N-GPU bench (torchrun) + the single-process multi-device test.  usage: tools/gpu_r2k.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_r2.py -q -m gpu --timeout 600 -k "multi_device" > gpurun_out/pytest_multidev_n$N.log 2>&1
tail -3 gpurun_out/pytest_multidev_n$N.log
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 ) > gpurun_out/bench_r2z_n$N.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_r2z_n$N.log
tail -c 2500 gpurun_out/bench_r2z_n$N.log
