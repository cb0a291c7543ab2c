#!/bin/bash
# validation of the tree: smoke, GPU tests, default bench line, reference arm
mkdir -p gpurun_out
T0=$(date +%s)
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r4y.log 2>&1; echo "smoke exit $? ($(( $(date +%s) - T0 )) s)" | tee -a gpurun_out/smoke_r4y.log
T0=$(date +%s)
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/pytest_gpu_r4y.log 2>&1; echo "pytest exit $? ($(( $(date +%s) - T0 )) s)" >> gpurun_out/pytest_gpu_r4y.log
tail -3 gpurun_out/pytest_gpu_r4y.log
T0=$(date +%s)
timeout 900 python bench.py > gpurun_out/bench_r4y.log 2> gpurun_out/bench_r4y.err; echo "bench exit $? ($(( $(date +%s) - T0 )) s)" | tee -a gpurun_out/bench_r4y.err
tail -c 3000 gpurun_out/bench_r4y.log
T0=$(date +%s)
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref_r4y.log 2>&1; echo "reference arm exit $? ($(( $(date +%s) - T0 )) s)" | tee -a gpurun_out/bench_ref_r4y.log
tail -c 1500 gpurun_out/bench_ref_r4y.log
