#!/bin/bash
# N-rank bench exactly as the driver launches it.  usage: tools/gpu_bench_n.sh N TAG [bench args]
N=${1:-2}; TAG=${2:-n$N}; shift; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/smi_$TAG.txt 2>&1
free -g >> gpurun_out/smi_$TAG.txt; nproc >> gpurun_out/smi_$TAG.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 3 --warmup 3 "$@" > gpurun_out/bench_$TAG.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_$TAG.log
grep -v "^\s" gpurun_out/bench_$TAG.log | tail -8
