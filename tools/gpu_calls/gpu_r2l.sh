#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do B200_1VN_V2=$v timeout 300 python tools/onevn_check.py 2>&1 | tail -3; done
for v in 0 1; do B200_1VN_V2=$v timeout 300 python tools/onevn_check.py 20000 1000 2>&1 | tail -2; done
timeout 900 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest_gpu_r2l.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2l.log
tail -15 gpurun_out/pytest_gpu_r2l.log | cut -c1-300
