#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu_r2c.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2c.log
tail -30 gpurun_out/pytest_gpu_r2c.log
: > gpurun_out/variants_r2c.log
for v in default tpw2 nopol; do
  if [ $v == default ]; then unset B200_RMSD_LIB; else export B200_RMSD_LIB=$PWD/variants/$v.so; fi
  timeout 200 python tools/variant_check.py 10000 1000 >> gpurun_out/variants_r2c.log 2>&1
done
unset B200_RMSD_LIB
grep "parity\|BEST" gpurun_out/variants_r2c.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2c.log 2>&1
tail -c 3000 gpurun_out/bench_r2c.log
