#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/cpptraj_leg.log
B200_BENCH_CPPTRAJ_LOG=$PWD/gpurun_out/cpptraj_leg.log python -c "
import json, bench
print(json.dumps(bench.cpptraj_leg(bench.CONFIGS['cfg2'], 3.0)))" > gpurun_out/cpptraj_leg.json 2>&1
cat gpurun_out/cpptraj_leg.json | cut -c1-900
grep -i "B200\|TIME\|took" gpurun_out/cpptraj_leg.log | head -40
