#!/bin/bash
mkdir -p gpurun_out
uname -r
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu_r2f.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2f.log
tail -5 gpurun_out/pytest_gpu_r2f.log
timeout 200 python tools/variant_check.py 10000 1000 2>&1 | grep "parity\|BEST"
timeout 600 python tools/e2e_host_modes.py > gpurun_out/e2e_host_modes_r2f.txt 2>&1
cat gpurun_out/e2e_host_modes_r2f.txt
B200_NO_POPULATE=1 timeout 600 python tools/e2e_host_modes.py 10000 1000 5 2>&1 | grep "fresh (cpptraj)"
