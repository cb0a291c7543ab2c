#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/e2e_pipe_ab_r4d.log; : > $L
for rep in 1 2 3; do
  timeout 120 python tools/e2e_pipe_ab.py >> $L 2>&1
  B200_PIPE_DOWNSTREAM=1 timeout 120 python tools/e2e_pipe_ab.py >> $L 2>&1
done
cat $L
