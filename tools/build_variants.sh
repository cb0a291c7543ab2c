#!/bin/bash
# Builds kernel-experiment variants of libb200rmsd.so under variants/ (git-ignored, travels to the GPU box) for
# same-box A/B runs (tools/variant_check.py).  usage: tools/build_variants.sh name:"-DFLAG=1 ..." ...
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$HERE/variants"
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  echo "== $name: $flags"
  B200_RMSD_LIB_OUT="$HERE/variants/$name.so" B200_NVCC_EXTRA="$flags" B200_NO_PTXAS_LOG=1 python "$HERE/cpptraj_b200/build.py" --force 2>&1 \
    | grep -A3 "pair_i8_kernelILb1ELi2ELb0" | grep -i "spill\|registers" || true
done
ls -la "$HERE/variants"
