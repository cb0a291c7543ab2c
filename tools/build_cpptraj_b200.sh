#!/bin/bash
# Applies the B200 host glue to a SCRATCH COPY of the cpptraj tree and (optionally) builds cpptraj with it.
#   tools/build_cpptraj_b200.sh [--check | --build | --build-intree | --rebuild] [reference dir] [scratch dir]
# --check (default): copy src/, patch, place src/cuda_b200/, and compile the glue and the patched reference files with
#                    -fsyntax-only -DCUDA_B200 (seconds; needs only g++).
# --build          : full copy, patch, the reference's own `configure ... -cuda_b200` (recipe of SURVEY.md 8c: OpenMP, no
#                    external libs) and `make install`; links the prebuilt cpptraj_b200/libb200rmsd.so of this repository
#                    (B200_RMSD_LIB=...: kernel changes need no relink).  Result: <scratch>/bin/cpptraj.OMP.b200.
# --build-intree   : same, but the kernels are compiled inside the cpptraj tree (src/cuda_b200/b200_rmsd.cu, nvcc sm_100a)
#                    and the CUDA runtime is linked statically: the self-contained build INTEGRATION.md describes.
# --rebuild        : keep an existing scratch tree (already patched), refresh src/cuda_b200/ and run make again.
# Nothing of the reference is copied into this repository; the patch is cpptraj_host/reference.patch.
set -e
MODE=--check
case "$1" in --check|--build|--build-intree|--rebuild) MODE=$1; shift;; esac
REF=${1:-/root/reference}
if [ "$MODE" == "--check" ]; then OUT=${2:-/tmp/cpptraj_b200_check}; else OUT=${2:-/tmp/cpptraj_b200_build}; fi
HERE=$(cd "$(dirname "$0")/.." && pwd)
export CUDA_HOME=${CUDA_HOME:-/usr/local/cuda}
place_glue() {
  mkdir -p "$OUT/src/cuda_b200"
  cp "$HERE"/cpptraj_host/src/cuda_b200/* "$HERE"/include/b200_rmsd.h "$HERE"/include/b200_rmsd_debug.h "$OUT/src/cuda_b200/"
  cp "$HERE"/cpptraj_b200/csrc/b200_rmsd.cu "$HERE"/cpptraj_b200/csrc/*.cuh "$HERE"/cpptraj_b200/csrc/host_util.h "$OUT/src/cuda_b200/"
}
if [ "$MODE" != "--rebuild" ]; then
  rm -rf "$OUT" && mkdir -p "$OUT"
  if [ "$MODE" == "--check" ]; then
    cp -r "$REF"/src "$OUT"/src; cp "$REF"/configure "$OUT"/; mkdir -p "$OUT"/cmake-cpptraj; cp "$REF"/cmake-cpptraj/CudaConfig.cmake "$OUT"/cmake-cpptraj/
  else
    cp -r "$REF"/. "$OUT"/
  fi
  chmod -R u+w "$OUT"
  ( cd "$OUT" && patch -p1 -s < "$HERE/cpptraj_host/reference.patch" )
fi
place_glue
if [ "$MODE" == "--check" ]; then
  cd "$OUT/src"
  for f in cuda_b200/B200_Rmsd.cpp Analysis_Rms2d.cpp Analysis_RmsAvgCorr.cpp Action_Rmsd.cpp Action_Align.cpp Exec_CrdTransform.cpp Exec_CrdAction.cpp Cluster/MetricArray.cpp Cluster/List.cpp \
           Cluster/Node.cpp Cluster/BestReps.cpp Cluster/Algorithm_Kmeans.cpp Cluster/Algorithm_HierAgglo.cpp Cluster/Results_Coords.cpp Cluster/Output.cpp Cluster/Control.cpp Cluster/PseudoF.cpp; do
    /usr/bin/g++ -std=c++11 -fsyntax-only -fopenmp -DCUDA_B200 -DNO_MATHLIB -DNONETCDF -I. -Icuda_b200 "$f"
    echo "syntax ok: $f"
  done
  bash -n "$OUT/configure" && echo "syntax ok: configure"
  exit 0
fi
# ---- full build (SURVEY.md 8c recipe) with the B200 target
cd "$OUT"
EXTRA="B200_RMSD_LIB=$HERE/cpptraj_b200/libb200rmsd.so"
[ "$MODE" == "--build-intree" ] && EXTRA=""
if [ "$MODE" != "--rebuild" ] || [ ! -f config.h ] || ! grep -q '^CUDA_B200_TARGET=cuda_b200' config.h; then
  CXX=/usr/bin/g++ CC=/usr/bin/gcc ./configure -openmp -nonetcdf -nobzlib -nozlib -nomathlib -noarpack -nofftw3 -pubfft \
      -noreadline -nosanderlib -notng --nobuildlibs -cuda_b200 $EXTRA gnu > configure.log 2>&1 || { tail -20 configure.log; exit 1; }
  cat > pub_fft_stub.c <<'EOC'
#include <stdlib.h>
void pubfft_init_(int*n,double*w,int*i){}
void pubfft_forward_(int*n,double*a,double*w,int*i){abort();}
void pubfft_back_(int*n,double*a,double*w,int*i){abort();}
EOC
  gcc -O2 -c pub_fft_stub.c -o src/pub_fft.o
  sed -i 's/^READLINE_LIB=-lreadline/READLINE_LIB=/' config.h      # (configure bug with -noreadline)
fi
touch src/pub_fft.o
make -j"$(nproc)" install > make.log 2>&1 || { tail -30 make.log; exit 1; }
ls -la bin/
# stage the binary and the reference's own test inputs / golden outputs where they travel to the GPU box
# (oracle/_ref/ is git-ignored: nothing of the reference enters the history)
STAGE="$HERE/oracle/_ref/cpptraj_b200"
mkdir -p "$STAGE"
cp bin/cpptraj.OMP.b200 "$STAGE"/cpptraj.B200
cp "$REF"/test/tz2.parm7 "$REF"/test/tz2.crd "$REF"/test/tz2.truncoct.parm7 "$REF"/test/tz2.truncoct.crd "$STAGE"/ 2>/dev/null || true
cp "$REF"/test/Test_2DRMS/rmsd.dat.save "$REF"/test/Test_2DRMS/rmsd.mass.dat.save "$REF"/test/Test_2DRMS/trp.dat.save \
   "$REF"/test/Test_2DRMS/nofit.dat.save "$REF"/test/Test_Cluster/cnumvtime.dat.save "$REF"/test/Test_Cluster/summary.dat.save \
   "$REF"/test/Test_RMSD/NoMod.dat.save "$STAGE"/
chmod u+w "$STAGE"/*
ls -la "$STAGE"
