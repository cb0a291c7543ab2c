#!/bin/bash
mkdir -p gpurun_out
(time B200_BENCH_CPPTRAJ_LOG=$PWD/gpurun_out/cpptraj_leg_r2v.log python bench.py) > gpurun_out/bench_r2v.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_r2v.log
tail -c 1500 gpurun_out/bench_r2v.log
