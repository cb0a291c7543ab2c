#!/bin/bash
# hieragglo re-scan scheduling: old library vs CTA-per-row with 8 / 12 / 16 loads in flight; then the GPU tests of the new default
mkdir -p gpurun_out; L=gpurun_out/hieragglo_rescan_r4v.log; : > $L
for v in ha_old ha_u8 ha_u12 ha_u16; do
  echo "== $v (10,000 and 30,000 frames, bench-style trajectory cache)" >> $L
  B200_RMSD_LIB=/root/repo/variants/$v.so B200_HA_DEBUG=1 timeout 300 python tools/hieragglo_traj.py 10000 1000 >> $L 2>&1
  B200_RMSD_LIB=/root/repo/variants/$v.so timeout 300 python tools/hieragglo_traj.py 30000 300 2>&1 | grep frames >> $L
done
cat $L | grep -v "^hieragglo n="
timeout 600 python -m pytest tests/test_hieragglo.py -q -m gpu 2>&1 | tail -2
