#!/bin/bash
# round-2 late: what a tensor SM would still cost if the per-pair solve ran elsewhere (debug mode 7) next to modes 0 and 2
mkdir -p gpurun_out; L=gpurun_out/i8_mode7_r4a.log; : > $L
for rep in 1 2; do
for m in 0 7 2; do
  B200_I8_DEBUG_MODE=$m B200_I8_CLOCKS=1 timeout 120 python tools/i8_modes.py >> $L 2>&1
done
done
cat $L
