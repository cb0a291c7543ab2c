#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/pipe_trace.py > gpurun_out/pipe_trace.log 2>&1
tail -38 gpurun_out/pipe_trace.log
timeout 600 python tools/e2e_host_modes.py 2>&1 | head -4
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_i8.py tests/test_gpu_r2.py -q -m gpu --timeout 600 -x 2>&1 | tail -3
