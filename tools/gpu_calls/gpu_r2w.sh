#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hieragglo.py -q -m gpu --timeout 600 -x > gpurun_out/pytest_hier.log 2>&1; echo "exit $?" >> gpurun_out/pytest_hier.log
tail -8 gpurun_out/pytest_hier.log
B200_HA_DEBUG=1 timeout 600 python tools/hieragglo_traj.py 1500 300 2>&1 | tail -8
B200_HA_DEBUG=1 timeout 600 python tools/hieragglo_traj.py 10000 1000 2>&1 | tail -8
B200_HA_DEBUG=1 timeout 900 python tools/hieragglo_time.py 30000 2>&1 | grep -v "n=3 " | grep "team=None\|team=8\|team=16 merges" 
