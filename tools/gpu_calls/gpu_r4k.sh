#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/init_time_r4k.log; : > $L
nvidia-smi --query-gpu=persistence_mode --format=csv >> $L
for i in 1 2 3 4; do python tools/init_time.py >> $L 2>&1; done
sleep 5
for i in 1 2; do python tools/init_time.py >> $L 2>&1; done
cat $L
