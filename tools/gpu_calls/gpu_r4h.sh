#!/bin/bash
mkdir -p gpurun_out; L=gpurun_out/variants_r4h.log; : > $L
for rep in 1 2; do
for v in default pack_u4 pack_u8; do
  if [ $v == default ]; then unset B200_RMSD_LIB; else export B200_RMSD_LIB=/root/repo/variants/$v.so; fi
  timeout 200 python tools/variant_check.py 2>&1 | grep -v "rep " >> $L
done
done
cat $L
