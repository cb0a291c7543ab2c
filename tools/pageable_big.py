"""The host path as cpptraj drives it at BASELINE config 5 size: pageable COORDS in, a touched pageable 20 GB triangle out
(cpptraj's Matrix<float> after its zero fill).  B200_PIPE_TRACE=1 for the time line.  usage: python tools/pageable_big.py [frames] [atoms] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cpptraj_b200 as b
import bench
nF = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
nA = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
b.init(1)
crd = np.empty((nF, 3 * nA), np.float32)
bench.gen_trajectory(20261020, nF, nA, crd)
sel = np.arange(nA, dtype=np.int32)
nT = nF * (nF - 1) // 2
out = np.zeros(nT, np.float32)          # touched
for r in range(reps):
    t0 = time.perf_counter()
    b.rms2d_tri(crd, sel, out=out)
    dt = time.perf_counter() - t0
    print("call %d: %.3f s  %.3e pairs/s  %.1f GB/s of results" % (r, dt, nT / dt, nT * 4 / dt / 1e9), flush=True)
