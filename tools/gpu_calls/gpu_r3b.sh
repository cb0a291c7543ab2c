#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu_r3b.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r3b.log
tail -3 gpurun_out/pytest_gpu_r3b.log
