#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/power_probe.py > gpurun_out/power_probe_r4b.log 2>&1
cat gpurun_out/power_probe_r4b.log
nvidia-smi -q -d POWER | head -40 > gpurun_out/smi_power_r4b.txt
