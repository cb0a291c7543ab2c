#!/bin/bash
# CUDA set-up time inside cpptraj.B200 (tiny deck) against OMP_NUM_THREADS and against the bare library (tools/init_time.py)
mkdir -p gpurun_out; L=$PWD/gpurun_out/setup_time_r4m.log; : > $L
D=$PWD/oracle/_ref/cpptraj_b200
W=$(mktemp -d); cd $W
printf "noprogress\nparm $D/tz2.parm7\ntrajin $D/tz2.crd\n2drms @CA R2D\nrun\n" > in
for t in 16 1 16 1; do
  for i in 1 2 3; do
    echo -n "OMP_NUM_THREADS=$t: " >> $L
    OMP_NUM_THREADS=$t $D/cpptraj.B200 -i in 2>&1 | grep -o "set-up [0-9.]* s" >> $L
  done
  python $OLDPWD/tools/init_time.py >> $L 2>&1
done
cat $L
