#!/bin/bash
mkdir -p gpurun_out
(time python bench.py --config cfg3 --steps 3 --warmup 3) > gpurun_out/bench_r2u_cfg3.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_r2u_cfg3.log
tail -c 3000 gpurun_out/bench_r2u_cfg3.log
