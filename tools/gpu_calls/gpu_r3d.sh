#!/bin/bash
for v in default p_early default p_early; do
  if [ $v == default ]; then unset B200_RMSD_LIB; else export B200_RMSD_LIB=$PWD/variants/$v.so; fi
  echo "== $v"; timeout 300 python tools/variant_check.py 2>&1 | tail -2
done
export B200_RMSD_LIB=$PWD/variants/p_early.so
timeout 600 python tools/variant_check.py 50000 2000 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_i8.py tests/test_gpu_parity.py -q -m gpu --timeout 600 -x 2>&1 | tail -2
unset B200_RMSD_LIB
timeout 600 python tools/variant_check.py 50000 2000 2>&1 | tail -1
