#!/bin/bash
# Same-box A/B of two builds of libb200rmsd.so on the cfg2 device-resident pass (box-to-box variation is ~3 %, more than
# most kernel tweaks).  Build the variant with  B200_RMSD_LIB_OUT=/root/repo/variant.so python cpptraj_b200/build.py --force
# usage (under gpurun): tools/gpu_ab.sh /root/repo/variant.so [reps]
VAR=${1:-/root/repo/variant.so}; REPS=${2:-3}
mkdir -p gpurun_out; : > gpurun_out/ab.log
for rep in $(seq $REPS); do
  for v in default variant; do
    if [ $v == default ]; then unset B200_RMSD_LIB; else export B200_RMSD_LIB=$VAR; fi
    echo -n "$v: " >> gpurun_out/ab.log
    timeout 200 python tools/i8_modes.py 2>&1 | head -1 >> gpurun_out/ab.log
  done
done
cat gpurun_out/ab.log
