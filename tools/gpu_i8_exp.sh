#!/bin/bash
# tcgen05 pair kernel experiments: i8 tests, timing modes with cycle counters, MMA probe variants.
# usage: tools/gpu_i8_exp.sh TAG "modes" [frames]
TAG=${1:-x}; MODES=${2:-"0 1 2 3 6"}; NF=${3:-10000}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_i8.py -x -q --timeout 120 > gpurun_out/i8_tests_$TAG.log 2>&1
echo "tests exit $?" >> gpurun_out/i8_tests_$TAG.log
tail -5 gpurun_out/i8_tests_$TAG.log
: > gpurun_out/i8_modes_$TAG.log
for m in $MODES; do
  B200_I8_DEBUG_MODE=$m B200_I8_CLOCKS=1 timeout 120 python tools/i8_modes.py $NF >> gpurun_out/i8_modes_$TAG.log 2>&1 || echo "mode $m failed/timeout" >> gpurun_out/i8_modes_$TAG.log
done
timeout 120 python - >> gpurun_out/i8_modes_$TAG.log 2>&1 <<'PY'
import cpptraj_b200 as b
b.init(1)
for v in (0, 1, 2):
    print("i8 mma probe variant %d: %.1f TOP/s" % (v, b.measure_i8_mma_peak(v)))
PY
cat gpurun_out/i8_modes_$TAG.log
