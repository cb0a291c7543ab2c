#!/bin/bash
# One gpurun call: GPU tests, FP64 peak probes, bench (N=1), reference arm, ncu launch list + full capture.
# usage: tools/gpu_round.sh TAG [kernel-regex]
TAG=${1:-r1}
KRE=${2:-pair_kernel}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi_$TAG.txt 2>&1
( nproc; lscpu | head -20; free -g ) > gpurun_out/host_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 120 python - > gpurun_out/peaks_$TAG.log 2>&1 <<'PY'
import cpptraj_b200 as b
b.init(1)
for v,name in ((0,'m8n8k4'),(1,'m16n8k4'),(2,'m16n8k8'),(3,'m16n8k16'),(4,'dfma')):
    print(name, '%.2f TFLOP/s' % b.measure_fp64_mma_peak(v))
PY
cat gpurun_out/peaks_$TAG.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_$TAG.log
tail -2 gpurun_out/bench_$TAG.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1
tail -1 gpurun_out/bench_ref_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 1 -c 1 \
    -o gpurun_out/prof_pair_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out
