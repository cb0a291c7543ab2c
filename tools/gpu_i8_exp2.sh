#!/bin/bash
# timing experiments with per-CTA clocks.  usage: tools/gpu_i8_exp2.sh TAG "cg:-:mode ..."
TAG=${1:-exp}; COMBOS=${2:-"2:1:2"}
mkdir -p gpurun_out
for crm in $COMBOS; do
  IFS=: read cg res m <<< "$crm"
  echo "=== cg $cg mode $m" >> gpurun_out/exp_$TAG.log
  B200_I8_CTA_GROUP=$cg B200_I8_DEBUG_MODE=$m B200_I8_CLOCKS=1 timeout 200 python tools/i8_modes.py >> gpurun_out/exp_$TAG.log 2>&1
done
cat gpurun_out/exp_$TAG.log
