#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/pytest_gpu_r2t.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu_r2t.log
tail -5 gpurun_out/pytest_gpu_r2t.log
for v in o_old default; do
  if [ $v == default ]; then unset B200_RMSD_LIB; else export B200_RMSD_LIB=$PWD/variants/$v.so; fi
  echo "== $v"; timeout 300 python tools/onevn_check.py 2>&1 | tail -2
  echo "== $v 100000 frames"; timeout 300 python tools/onevn_check.py 100000 5000 2>&1 | tail -1
done
