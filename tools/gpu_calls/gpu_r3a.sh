#!/bin/bash
for v in o_s6p6 default o_s8p6 o_s7p7 default; do
  if [ $v == default ]; then unset B200_RMSD_LIB; else export B200_RMSD_LIB=$PWD/variants/$v.so; fi
  echo "== $v"; timeout 300 python tools/onevn_check.py 100000 5000 2>&1 | tail -2
done
unset B200_RMSD_LIB
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_r2.py -q -m gpu --timeout 600 -x 2>&1 | tail -2
