#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rmsavgcorr.py tests/test_hieragglo.py -q -m gpu --timeout 600 > gpurun_out/pytest_r2z.log 2>&1; echo "exit $?" >> gpurun_out/pytest_r2z.log
tail -6 gpurun_out/pytest_r2z.log
timeout 900 python -m pytest tests/test_gpu_cpptraj_e2e.py -q -m gpu --timeout 600 -k "rmsavgcorr" 2>&1 | tail -2
timeout 600 python tools/rmsavgcorr_time.py 10000 1000 > gpurun_out/rmsavgcorr_time_r2b.log 2>&1; cat gpurun_out/rmsavgcorr_time_r2b.log
(B200_HA_DEBUG=1 timeout 600 python tools/hieragglo_traj.py 10000 1000; B200_HA_DEBUG=1 timeout 900 python tools/hieragglo_time.py 30000 50000) 2>&1 | grep -v "n=3 " > gpurun_out/hieragglo_timing_r2b.log; grep -v "^hieragglo" gpurun_out/hieragglo_timing_r2b.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:avgcorr_kernel -s 1 -c 1 -o gpurun_out/prof_avgcorr_r2c -f python tools/rmsavgcorr_time.py 3000 1000 > gpurun_out/ncu_avgcorr_r2c.log 2>&1
