#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu_r2h.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2h.log
tail -60 gpurun_out/pytest_gpu_r2h.log | cut -c1-300
