#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/e2e_pipe_ab_r4c.log; : > $L
for rep in 1 2; do
  B200_UPLOAD_2D=1 timeout 120 python tools/e2e_pipe_ab.py >> $L 2>&1
  timeout 120 python tools/e2e_pipe_ab.py >> $L 2>&1
  B200_PIPE_UPSTREAMS=2 timeout 120 python tools/e2e_pipe_ab.py >> $L 2>&1
  B200_PIPE_ROWS=504 timeout 120 python tools/e2e_pipe_ab.py >> $L 2>&1
  B200_PIPE_ROWS=504 B200_PIPE_UPSTREAMS=2 timeout 120 python tools/e2e_pipe_ab.py >> $L 2>&1
done
B200_PIPE_TRACE=1 timeout 120 python tools/e2e_pipe_ab.py 10000 1000 1 2> gpurun_out/pipe_trace_r4c.txt >> $L
B200_PIPE_TRACE=1 B200_PIPE_UPSTREAMS=2 timeout 120 python tools/e2e_pipe_ab.py 10000 1000 1 2> gpurun_out/pipe_trace_up2_r4c.txt >> $L
cat $L
