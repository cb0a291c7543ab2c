#!/bin/bash
# quick loop: i8 tests + bench.  usage: tools/gpu_quick.sh TAG [extra bench args]
TAG=${1:-q}; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_i8.py -x -q --timeout 200 > gpurun_out/i8_tests_$TAG.log 2>&1
echo "tests exit $?" >> gpurun_out/i8_tests_$TAG.log
tail -4 gpurun_out/i8_tests_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$TAG.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_$TAG.log
python - <<PY
import json
for l in open('gpurun_out/bench_$TAG.log'):
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']
        print('value %.3e e2e %.3e ms/step %.3f frac_exec %s launch_ms %.3f share %.3f clocks %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], r.get('frac_executed'), r['avg_launch_ms'], r['kernel_share_of_step'], d['clocks']))
    elif 'exit' in l or 'Error' in l: print(l.strip())
PY
