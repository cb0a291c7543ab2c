#!/bin/bash
# ncu full capture of the tcgen05 pair kernel on the modes micro-benchmark (1 GPU).  usage: tools/gpu_ncu_i8.sh TAG [launch-skip]
TAG=${1:-x}; SKIP=${2:-4}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_i8_kernel -s $SKIP -c 1 \
    -o gpurun_out/prof_pair_i8_$TAG -f python tools/i8_modes.py > gpurun_out/ncu_i8_$TAG.log 2>&1
tail -3 gpurun_out/ncu_i8_$TAG.log
