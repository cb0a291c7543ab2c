#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n2.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_r2.py -q -m gpu --timeout 600 -k "multi_device" > gpurun_out/pytest_gpu_r2g.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2g.log
tail -25 gpurun_out/pytest_gpu_r2g.log
timeout 300 tools/microbench/host_path_probe 1024 16 2>&1 | head -12 > gpurun_out/host_path_probe_n2.txt
cat gpurun_out/host_path_probe_n2.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r2g_n2.log 2>&1
tail -c 1500 gpurun_out/bench_r2g_n2.log
