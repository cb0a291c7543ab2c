#!/bin/bash
mkdir -p gpurun_out
timeout 300 tools/microbench/host_path_probe 1024 16 2>&1 | grep -A30 "madvise\|^---" > gpurun_out/host_path_probe_r2b.txt
cat gpurun_out/host_path_probe_r2b.txt
: > gpurun_out/variants_r2b.log
for v in default nopol nopol2 nowin nofp; do
  if [ $v == default ]; then unset B200_RMSD_LIB; else export B200_RMSD_LIB=$PWD/variants/$v.so; fi
  timeout 200 python tools/variant_check.py 10000 1000 >> gpurun_out/variants_r2b.log 2>&1
done
grep "parity\|BEST" gpurun_out/variants_r2b.log
export B200_RMSD_LIB=$PWD/variants/nopol.so
timeout 600 python -m pytest tests/test_gpu_i8.py tests/test_gpu_parity.py -q -m gpu --timeout 600 > gpurun_out/pytest_nopol_r2b.log 2>&1
tail -15 gpurun_out/pytest_nopol_r2b.log
