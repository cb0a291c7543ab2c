#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/san2.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import cpptraj_b200 as b
from cpptraj_b200.synth import make_trajectory
b.init(1)
rng = np.random.default_rng(3)
n = 400
x = rng.standard_normal((n, 3)) * 3
x[5] = x[9]; x[100] = x[7]
d = np.sqrt(((x[:, None] - x[None]) ** 2).sum(-1)); d = np.round(d * 8) / 8
tri = d[np.triu_indices(n, 1)].astype(np.float32)
for linkage in (0, 1, 2):
    for team in ("1", "4"):
        os.environ["B200_HA_TEAM"] = team
        b.hieragglo(tri, n, linkage, 3, None)
crd, mass = make_trajectory(5, 120, 300)
sel = np.arange(300, dtype=np.int32)
b.rmsavgcorr(crd, sel, np.arange(1, 120, dtype=np.int32), mass=mass[sel])
print("sanitizer workload 2 done")
P
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/san2.py > gpurun_out/sanitize_racecheck2.log 2>&1; tail -4 gpurun_out/sanitize_racecheck2.log
timeout 1500 compute-sanitizer --tool synccheck --print-limit 20 python /tmp/san2.py > gpurun_out/sanitize_synccheck2.log 2>&1; tail -3 gpurun_out/sanitize_synccheck2.log
