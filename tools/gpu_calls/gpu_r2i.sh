#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_r2i.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_r2i.log
tail -c 1200 gpurun_out/bench_r2i.log
