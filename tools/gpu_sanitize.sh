#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on shared memory) over the kernels added / changed in round 2, on small inputs
mkdir -p gpurun_out
cat > /tmp/san.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import cpptraj_b200 as b
from cpptraj_b200.synth import make_trajectory
b.init(1)
crd, mass = make_trajectory(5, 300, 1100)
sel = np.arange(0, 1100, dtype=np.int32)
tri = b.rms2d_tri(crd, sel)
for linkage in (0, 1, 2):
    for team in ("1", "4"):
        os.environ["B200_HA_TEAM"] = team
        b.hieragglo(tri, 300, linkage, 3, None)
b.rmsavgcorr(crd[:120], sel, np.arange(1, 120, dtype=np.int32), mass=mass[sel])
ref = crd[3].reshape(-1, 3)[sel].astype(np.float64); ref -= ref.mean(0)
b.rmsd_1vN(crd, sel, ref, want_rot=True)
b.rmsd_1vN(crd, sel[::3].copy(), ref[::3] - ref[::3].mean(0), want_rot=True)
print("sanitizer workload done")
P
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san.py > gpurun_out/sanitize_memcheck.log 2>&1; tail -4 gpurun_out/sanitize_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 5 python /tmp/san.py > gpurun_out/sanitize_racecheck.log 2>&1; tail -4 gpurun_out/sanitize_racecheck.log
