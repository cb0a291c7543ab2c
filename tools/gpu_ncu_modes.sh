#!/bin/bash
# ncu full captures of the tcgen05 pair kernel for several (cta group, debug mode) combinations.
# usage: tools/gpu_ncu_modes.sh TAG "cg:mode cg:mode ..."
TAG=${1:-x}; COMBOS=${2:-"2:0 2:2 1:2"}
mkdir -p gpurun_out
for cm in $COMBOS; do
  cg=${cm%%:*}; m=${cm##*:}
  B200_I8_CTA_GROUP=$cg B200_I8_DEBUG_MODE=$m timeout 600 ncu --set full --clock-control none --import-source on \
      -k regex:pair_i8_kernel -s 4 -c 1 -o gpurun_out/prof_${TAG}_cg${cg}_m${m} -f python tools/i8_modes.py > gpurun_out/ncu_${TAG}_cg${cg}_m${m}.log 2>&1
  tail -2 gpurun_out/ncu_${TAG}_cg${cg}_m${m}.log
done
ls -la gpurun_out | tail
