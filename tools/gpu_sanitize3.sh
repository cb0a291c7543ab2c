#!/bin/bash
# compute-sanitizer on the kernels changed late in round 2: hieragglo re-scans (CTA per row / warp per row, long lists from a
# tight cluster, ties) with the full 16-CTA team, and the centroid kernel with its running sum in shared memory
mkdir -p gpurun_out
cat > /tmp/san3.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import cpptraj_b200 as b
from cpptraj_b200.synth import make_trajectory
b.init(1)
rng = np.random.default_rng(3)
n = 900
x = rng.standard_normal((n, 3)) * 3
x[:450] = x[0] + rng.standard_normal((450, 3)) * 0.01      # a tight cluster: the closest of almost every other cluster
x[5] = x[9]; x[700] = x[7]
d = np.sqrt(((x[:, None] - x[None]) ** 2).sum(-1)); d = np.round(d * 64) / 64
tri = d[np.triu_indices(n, 1)].astype(np.float32)
for linkage in (0, 1, 2):
    for team in ("16", "4"):
        os.environ["B200_HA_TEAM"] = team
        b.hieragglo(tri, n, linkage, 3, None)
crd, mass = make_trajectory(5, 160, 300)
sel = np.arange(300, dtype=np.int32)
lists = [np.arange(0, 90, dtype=np.int32), np.arange(90, 91, dtype=np.int32), np.arange(91, 160, dtype=np.int32)]
for fit in (True, False):
    for m in (None, mass[sel]):
        b.build_centroids(crd, sel, lists, mass=m, fit=fit)
print("sanitizer workload 3 done")
P
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san3.py > gpurun_out/sanitize_${tool}3.log 2>&1; echo "$tool: $(grep -c 'sanitizer workload 3 done' gpurun_out/sanitize_${tool}3.log) done; $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}3.log | tail -1)"
done
