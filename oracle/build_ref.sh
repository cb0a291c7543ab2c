#!/bin/bash
# Build oracle/_ref/libcpptraj_ref_rmsd.so: the reference's OWN source files
# (compiled where they lie under /root/reference/src, nothing copied) plus our
# thin driver oracle/ref_driver.cpp.  Does not run the reference's build system.
# Only possible where /root/reference exists (the build container); the GPU box
# uses the prebuilt .so that travels with the snapshot.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${CPPTRAJ_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: $REF/src not present; keeping any prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
# NB: the image exports CXX=/opt/gcc/bin/g++ whose libgomp.spec is missing; use the system g++.
CXX="${REF_CXX:-/usr/bin/g++}"
# Same optimisation level cpptraj's configure picks for gnu (-O3) + OpenMP.
FLAGS="-O3 -fopenmp -fPIC -std=c++11 -w -I$REF/src"
SRCS="Frame Matrix_3x3 Vec3 Box CoordinateInfo CompactFrameArray AtomMask MaskToken ArgList StringRoutines DistRoutines RangeToken Range Atom Residue Molecule CpptrajStdio NameType ReplicaDimArray Unit Segment"
OBJS=""
for s in $SRCS; do
  if [ -f "$REF/src/$s.cpp" ]; then
    $CXX $FLAGS -c "$REF/src/$s.cpp" -o "$OUT/obj/$s.o"
    OBJS="$OBJS $OUT/obj/$s.o"
  fi
done
# hierarchical clustering bookkeeping (closest-index rules of the cluster distance matrix)
$CXX $FLAGS -c "$REF/src/Cluster/DynamicMatrix.cpp" -o "$OUT/obj/Cluster_DynamicMatrix.o"
OBJS="$OBJS $OUT/obj/Cluster_DynamicMatrix.o"
$CXX $FLAGS -c "$HERE/ref_driver.cpp" -o "$OUT/obj/ref_driver.o"
$CXX -shared -fopenmp -o "$OUT/libcpptraj_ref_rmsd.so" "$OUT/obj/ref_driver.o" $OBJS -Wl,--no-undefined
echo "built $OUT/libcpptraj_ref_rmsd.so"
