"""b200_hieragglo on the pairwise cache of a bench-style trajectory (a quarter of the frames are rigid copies of the base
conformation: a tight cluster that is the closest of almost every other cluster -- long re-scan lists).
usage: B200_HA_DEBUG=1 python tools/hieragglo_traj.py [frames atoms]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpptraj_b200 as b
from cpptraj_b200.synth import make_trajectory
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
na = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
crd, _ = make_trajectory(20261018, nf, na)
b.init(1)
tri = b.rms2d_tri(crd, np.arange(na, dtype=np.int32))
for linkage in (1, 0, 2):
    t0 = time.time()
    into, frm, fmin = b.hieragglo(tri, nf, linkage, 10, None)
    print("frames %d linkage %d: %d merges in %.3f s, last min %.4f" % (nf, linkage, len(into), time.time() - t0, fmin[-1]), flush=True)
if nf <= 1500:
    from oracle.pyoracle import Oracle
    t0 = time.time()
    want = Oracle().hieragglo(tri, nf, 1, 10, None)
    got = b.hieragglo(tri, nf, 1, 10, None)
    print("CPU restatement %.2f s; merges equal %s" % (time.time() - t0, np.array_equal(want[0], got[0]) and np.array_equal(want[1], got[1])))
