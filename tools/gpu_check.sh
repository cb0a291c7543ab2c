#!/bin/bash
# Runs on the GPU box via gpurun: GPU tests, peak probes, short bench.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 120 python - > gpurun_out/peaks.log 2>&1 <<'PY'
import cpptraj_b200 as b
b.init(1)
for v,name in ((0,'m8n8k4'),(1,'m16n8k4'),(2,'m16n8k8'),(3,'m16n8k16'),(4,'dfma')):
    print(name, '%.2f TFLOP/s' % b.measure_fp64_mma_peak(v))
PY
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/peaks.log; tail -2 gpurun_out/bench.log
