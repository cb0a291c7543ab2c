#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu_r2e.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2e.log
tail -12 gpurun_out/pytest_gpu_r2e.log
timeout 200 python tools/variant_check.py 10000 1000 2>&1 | grep "parity\|BEST"
timeout 600 python tools/e2e_host_modes.py > gpurun_out/e2e_host_modes_r2e.txt 2>&1
cat gpurun_out/e2e_host_modes_r2e.txt
