#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu_r2d.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2d.log
tail -30 gpurun_out/pytest_gpu_r2d.log
: > gpurun_out/variants_r2d.log
for v in default tpw2; do
  if [ $v == default ]; then unset B200_RMSD_LIB; else export B200_RMSD_LIB=$PWD/variants/$v.so; fi
  timeout 200 python tools/variant_check.py 10000 1000 >> gpurun_out/variants_r2d.log 2>&1
done
unset B200_RMSD_LIB
grep "parity\|BEST\|Error\|error" gpurun_out/variants_r2d.log
timeout 600 python tools/e2e_host_modes.py > gpurun_out/e2e_host_modes_r2d.txt 2>&1
cat gpurun_out/e2e_host_modes_r2d.txt
