#!/bin/bash
# Golden outputs for tests/test_gpu_cpptraj_e2e.py::test_cpptraj_cluster_sieve_restore_on_b200: the UNMODIFIED reference
# (plain OpenMP build, recipe of SURVEY.md 8c) runs the sieve deck on the CPU; cluster number vs time and the summary go to
# tests/golden/.  usage: tools/make_golden_cluster_sieve.sh [reference dir] [scratch dir]   (~5 min)
set -e
REF=${1:-/root/reference}
OUT=${2:-/tmp/cpptraj_plain_build}
HERE=$(cd "$(dirname "$0")/.." && pwd)
if [ ! -x "$OUT/bin/cpptraj.OMP" ]; then
  rm -rf "$OUT" && mkdir -p "$OUT" && cp -r "$REF"/. "$OUT"/ && chmod -R u+w "$OUT"
  cd "$OUT"
  CXX=/usr/bin/g++ CC=/usr/bin/gcc ./configure -openmp -nonetcdf -nobzlib -nozlib -nomathlib -noarpack -nofftw3 -pubfft \
      -noreadline -nosanderlib -notng --nobuildlibs gnu > configure.log 2>&1
  printf '#include <stdlib.h>\nvoid pubfft_init_(int*n,double*w,int*i){}\nvoid pubfft_forward_(int*n,double*a,double*w,int*i){abort();}\nvoid pubfft_back_(int*n,double*a,double*w,int*i){abort();}\n' > pub_fft_stub.c
  gcc -O2 -c pub_fft_stub.c -o src/pub_fft.o
  sed -i 's/^READLINE_LIB=-lreadline/READLINE_LIB=/' config.h
  make -j"$(nproc)" install > make.log 2>&1 || { tail -30 make.log; exit 1; }
fi
W=$(mktemp -d)
cd "$W"
cat > cluster.in <<EOD
noprogress
parm $REF/test/tz2.parm7
trajin $REF/test/tz2.crd
cluster crd1 @CA clusters 5 rms out sieve5.out summary sieve5.summary.dat sieve 5 bestrep cumulative includesieveincalc
EOD
OMP_NUM_THREADS=4 "$OUT/bin/cpptraj.OMP" -i cluster.in > run.log 2>&1 || { tail -20 run.log; exit 1; }
cp sieve5.out "$HERE/tests/golden/cluster_sieve5.out"
cp sieve5.summary.dat "$HERE/tests/golden/cluster_sieve5.summary.dat"
ls -la "$HERE/tests/golden/" | grep sieve
