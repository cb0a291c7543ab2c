#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hieragglo.py -q -m gpu --timeout 600 -x > gpurun_out/pytest_hier.log 2>&1; echo "exit $?" >> gpurun_out/pytest_hier.log
tail -15 gpurun_out/pytest_hier.log
timeout 900 python -m pytest tests/test_gpu_cpptraj_e2e.py -q -m gpu --timeout 600 -k "hier or cluster" > gpurun_out/pytest_hier_e2e.log 2>&1; echo "exit $?" >> gpurun_out/pytest_hier_e2e.log
tail -15 gpurun_out/pytest_hier_e2e.log
timeout 900 python tools/hieragglo_time.py 2000 10000 30000 > gpurun_out/hieragglo_time.log 2>&1
cat gpurun_out/hieragglo_time.log
