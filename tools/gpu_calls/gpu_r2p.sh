#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rmsavgcorr.py -q -m gpu --timeout 600 > gpurun_out/pytest_r2p.log 2>&1; echo "exit $?" >> gpurun_out/pytest_r2p.log
tail -15 gpurun_out/pytest_r2p.log
for nw in 1 2 4; do
  echo "== NWIN $nw"
  B200_AVGCORR_NWIN=$nw timeout 900 python tools/rmsavgcorr_time.py 10000 1000 2>&1 | grep -v "^CPU"
done > gpurun_out/rmsavgcorr_time.log 2>&1
cat gpurun_out/rmsavgcorr_time.log
