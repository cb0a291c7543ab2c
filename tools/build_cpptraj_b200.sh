#!/bin/bash
# Applies the B200 host glue to a SCRATCH COPY of the cpptraj tree and (optionally) builds cpptraj.B200.
#   tools/build_cpptraj_b200.sh [--check | --build] [reference dir] [scratch dir]
# --check (default): copy, patch, place src/cuda_b200/, and compile the glue and the patched reference files with
#                    -fsyntax-only -DCUDA_B200 (seconds; needs only g++).
# --build          : also run the reference's configure (recipe of SURVEY.md 8c: OpenMP, no external libs) and make,
#                    linking libb200rmsd.so; result: <scratch>/bin/cpptraj.OMP with the B200 branches compiled in.
# Nothing of the reference is copied into this repository; the patch is cpptraj_host/reference.patch.
set -e
MODE=--check
if [ "$1" == "--check" ] || [ "$1" == "--build" ]; then MODE=$1; shift; fi
REF=${1:-/root/reference}
if [ "$MODE" == "--check" ]; then OUT=${2:-/tmp/cpptraj_b200_check}; else OUT=${2:-/tmp/cpptraj_b200_build}; fi
HERE=$(cd "$(dirname "$0")/.." && pwd)
rm -rf "$OUT" && mkdir -p "$OUT"
if [ "$MODE" == "--check" ]; then cp -r "$REF"/src "$OUT"/src; else cp -r "$REF"/. "$OUT"/; fi
chmod -R u+w "$OUT"
( cd "$OUT" && patch -p1 -s < "$HERE/cpptraj_host/reference.patch" )
mkdir -p "$OUT/src/cuda_b200"
cp "$HERE"/cpptraj_host/src/cuda_b200/B200_Rmsd.h "$HERE"/cpptraj_host/src/cuda_b200/B200_Rmsd.cpp "$HERE"/include/b200_rmsd.h "$OUT/src/cuda_b200/"
cd "$OUT/src"
for f in cuda_b200/B200_Rmsd.cpp Analysis_Rms2d.cpp Cluster/MetricArray.cpp Action_Rmsd.cpp Cluster/List.cpp; do
  /usr/bin/g++ -std=c++11 -fsyntax-only -fopenmp -DCUDA_B200 -DNO_MATHLIB -DNONETCDF -I. -Icuda_b200 "$f"
  echo "syntax ok: $f"
done
[ "$MODE" == "--check" ] && exit 0
# ---- full build (SURVEY.md 8c recipe) + B200 objects
cd "$OUT"
CXX=/usr/bin/g++ CC=/usr/bin/gcc ./configure -openmp -nonetcdf -nobzlib -nozlib -nomathlib -noarpack -nofftw3 -pubfft \
    -noreadline -nosanderlib -notng --nobuildlibs gnu > configure.log 2>&1
cat > pub_fft_stub.c <<'EOC'
#include <stdlib.h>
void pubfft_init_(int*n,double*w,int*i){}
void pubfft_forward_(int*n,double*a,double*w,int*i){abort();}
void pubfft_back_(int*n,double*a,double*w,int*i){abort();}
EOC
gcc -O2 -c pub_fft_stub.c -o src/pub_fft.o
sed -i 's/^READLINE_LIB=-lreadline/READLINE_LIB=/' config.h
# compile flags: add -DCUDA_B200; link: the glue object + libb200rmsd.so (rpath to this repository)
sed -i "s|^DIRECTIVES=|DIRECTIVES=-DCUDA_B200 |" config.h
sed -i "s|^LDFLAGS=|LDFLAGS=$OUT/src/cuda_b200/B200_Rmsd.o -L$HERE/cpptraj_b200 -lb200rmsd -Wl,-rpath,$HERE/cpptraj_b200 |" config.h
/usr/bin/g++ -std=c++11 -O2 -fopenmp -DCUDA_B200 -Isrc -Isrc/cuda_b200 -c src/cuda_b200/B200_Rmsd.cpp -o src/cuda_b200/B200_Rmsd.o
make -j"$(nproc)" install > make.log 2>&1 || { tail -30 make.log; exit 1; }
ls -la bin/
# stage the binary and the reference's own Test_2DRMS inputs / golden outputs where they travel to the GPU box
# (oracle/_ref/ is git-ignored: nothing of the reference enters the history)
STAGE="$HERE/oracle/_ref/cpptraj_b200"
mkdir -p "$STAGE"
cp bin/cpptraj.OMP "$STAGE"/cpptraj.B200
cp "$REF"/test/tz2.parm7 "$REF"/test/tz2.crd "$STAGE"/
cp "$REF"/test/Test_2DRMS/rmsd.dat.save "$REF"/test/Test_2DRMS/rmsd.mass.dat.save "$REF"/test/Test_2DRMS/trp.dat.save \
   "$REF"/test/Test_2DRMS/nofit.dat.save "$REF"/test/Test_Cluster/cnumvtime.dat.save "$REF"/test/Test_Cluster/summary.dat.save "$REF"/test/Test_RMSD/NoMod.dat.save "$STAGE"/
chmod u+w "$STAGE"/*
ls -la "$STAGE"
