#!/bin/bash
# CTA-pair (tcgen05.mma.cta_group::2) bring-up: exactness first under short timeouts, then parity, bench, timing modes.
# usage: tools/gpu_cg2.sh TAG ["modes"]
TAG=${1:-cg2}; MODES=${2:-"0 2 3"}
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_i8.py -x -q --timeout 100 -k "integer_cov" > gpurun_out/cg2_stage1_$TAG.log 2>&1
rc=$?; echo "stage1 exit $rc" >> gpurun_out/cg2_stage1_$TAG.log
tail -15 gpurun_out/cg2_stage1_$TAG.log
if [ $rc -ne 0 ]; then echo "stage1 failed: stopping"; exit 0; fi
timeout 600 python -m pytest tests/test_gpu_i8.py -q --timeout 200 > gpurun_out/cg2_stage2_$TAG.log 2>&1
echo "stage2 exit $?" >> gpurun_out/cg2_stage2_$TAG.log
tail -8 gpurun_out/cg2_stage2_$TAG.log
for cg in 2 1; do
  B200_I8_CTA_GROUP=$cg timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_cg$cg.log 2>&1
  echo "bench exit $?" >> gpurun_out/bench_${TAG}_cg$cg.log
done
python - <<PY
import json
for cg in (2, 1):
    for l in open('gpurun_out/bench_${TAG}_cg%d.log' % cg):
        if l.startswith('{'):
            d=json.loads(l); r=d['roofline']
            print('cg%d value %.3e e2e %.3e ms/step %.3f frac_exec %s launch_ms %.3f share %.3f clocks %s' % (cg, d['value'], d['e2e']['value'], d['ms_per_step'], r.get('frac_executed'), r['avg_launch_ms'], r['kernel_share_of_step'], d['clocks']))
        elif 'exit' in l or 'Error' in l: print(cg, l.strip())
PY
for m in $MODES; do
  B200_I8_DEBUG_MODE=$m B200_I8_CLOCKS=1 timeout 200 python tools/i8_modes.py >> gpurun_out/modes_$TAG.log 2>&1
done
cat gpurun_out/modes_$TAG.log
