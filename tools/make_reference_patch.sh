#!/bin/bash
# Regenerates cpptraj_host/reference.patch from a patched scratch copy of the cpptraj tree: unified diffs (2 lines of
# context) of exactly the files the B200 host glue touches.  New files (src/cuda_b200/*) are not in the patch: they
# live under cpptraj_host/src/cuda_b200/ and are copied by tools/build_cpptraj_b200.sh.
#   tools/make_reference_patch.sh [scratch dir] [reference dir]
set -e
SCR=${1:-/tmp/cpptraj_b200_build}; REF=${2:-/root/reference}
HERE=$(cd "$(dirname "$0")/.." && pwd)
FILES="configure cmake-cpptraj/CudaConfig.cmake src/CMakeLists.txt src/Makefile src/Matrix.h src/Action_Align.cpp src/Action_Align.h src/Exec_CrdTransform.cpp src/Action_Rmsd.cpp src/Action_Rmsd.h src/Analysis_Rms2d.cpp src/Analysis_RmsAvgCorr.cpp
       src/Exec_CrdAction.cpp src/DataSet_Coords_CRD.h src/Cluster/Algorithm_HierAgglo.cpp src/Cluster/Algorithm_HierAgglo.h src/Cluster/Algorithm_Kmeans.cpp src/Cluster/BestReps.cpp src/Cluster/BestReps.h src/Cluster/Output.cpp src/Cluster/Control.cpp src/Cluster/Control.h
       src/Cluster/List.cpp src/Cluster/MetricArray.cpp src/Cluster/MetricArray.h src/Cluster/Results_Coords.cpp src/Cluster/Metric_RMS.h src/Cluster/Node.cpp src/Cluster/Node.h src/Cluster/PseudoF.cpp"
OUT="$HERE/cpptraj_host/reference.patch"
: > "$OUT"
for f in $FILES; do
  if ! diff -q "$REF/$f" "$SCR/$f" > /dev/null; then
    diff -U2 --label "a/$f" --label "b/$f" "$REF/$f" "$SCR/$f" >> "$OUT" || true
  fi
done
echo "$(grep -c '^--- a/' "$OUT") files, $(wc -l < "$OUT") lines -> $OUT"
