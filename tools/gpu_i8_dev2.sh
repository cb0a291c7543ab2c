#!/bin/bash
# tcgen05 pair kernel dev loop: exactness + parity under short timeouts, then per-mode timings.
# usage: tools/gpu_i8_dev2.sh TAG ["cg:- cg:- ..."] ["dbg modes"]
TAG=${1:-dev}; COMBOS=${2:-"2:0 1:0"}; MODES=${3:-"0"}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_i8.py -x -q --timeout 120 -k "integer_cov or packed" > gpurun_out/dev_stage1_$TAG.log 2>&1
rc=$?; echo "stage1 exit $rc" >> gpurun_out/dev_stage1_$TAG.log
tail -15 gpurun_out/dev_stage1_$TAG.log
if [ $rc -ne 0 ]; then echo "stage1 failed: stopping"; exit 0; fi
timeout 900 python -m pytest tests/test_gpu_i8.py -q --timeout 300 > gpurun_out/dev_stage2_$TAG.log 2>&1
echo "stage2 exit $?" >> gpurun_out/dev_stage2_$TAG.log
tail -12 gpurun_out/dev_stage2_$TAG.log
for cr in $COMBOS; do
  cg=${cr%%:*}; res=${cr##*:}
  for m in $MODES; do
    echo "=== cg $cg mode $m" >> gpurun_out/modes_$TAG.log
    B200_I8_CTA_GROUP=$cg B200_I8_DEBUG_MODE=$m timeout 200 python tools/i8_modes.py >> gpurun_out/modes_$TAG.log 2>&1
  done
done
cat gpurun_out/modes_$TAG.log
