#!/bin/bash
# round 2, call a: GPU tests after the hygiene changes, host-path probe, pair-kernel variants (same box)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi_r2a.txt 2>&1
( nproc; lscpu | head -25; free -g; numactl -H 2>/dev/null ) > gpurun_out/host_r2a.txt 2>&1
timeout 600 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu_r2a.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2a.log
tail -4 gpurun_out/pytest_gpu_r2a.log
timeout 300 tools/microbench/host_path_probe 1024 16 > gpurun_out/host_path_probe_r2a.txt 2>&1
cat gpurun_out/host_path_probe_r2a.txt
: > gpurun_out/variants_r2a.log
for v in default keep defer nopol x48; do
  if [ $v == default ]; then unset B200_RMSD_LIB; else export B200_RMSD_LIB=$PWD/variants/$v.so; fi
  timeout 200 python tools/variant_check.py 10000 1000 --clocks >> gpurun_out/variants_r2a.log 2>&1
done
unset B200_RMSD_LIB
grep "parity\|BEST\|fpDone\|solve_fp64\|mma_wait_full\|mma_total" gpurun_out/variants_r2a.log
