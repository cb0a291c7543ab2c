#!/bin/bash
for v in default p_t1s3 p_t2s3 default; do
  if [ $v == default ]; then unset B200_RMSD_LIB; else export B200_RMSD_LIB=$PWD/variants/$v.so; fi
  echo "== $v"; timeout 300 python tools/variant_check.py 2>&1 | tail -3
done
