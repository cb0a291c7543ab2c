"""Times b200_rmsavgcorr (all window sizes, 'first' mode) on a synthetic trajectory and, on a prefix of the window sizes,
the CPU restatement / the reference's own Frame arithmetic (OpenMP over window sizes, as the reference).
usage: python tools/rmsavgcorr_time.py [nframes natoms]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpptraj_b200 as b
from cpptraj_b200.synth import make_trajectory
from oracle.pyoracle import Oracle, Reference, have_reference

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
na = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
crd, mass = make_trajectory(7, nf, na)
sel = np.arange(na, dtype=np.int32)
b.init(1)
b.rmsavgcorr(crd[:50], sel, np.arange(1, 50))
win = np.arange(1, nf, dtype=np.int32)
items = float(np.sum(nf - win + 1))
for rep in range(2):
    t0 = time.time()
    avg, sd = b.rmsavgcorr(crd, sel, win)
    dt = time.time() - t0
    print("B200: %d frames x %d atoms, %d window sizes, %.3e averaged-frame RMSDs in %.3f s = %.3e /s; %.1f GB/s of prefix rows"
          % (nf, na, len(win), items, dt, items / dt, items * na * 48 / dt / 1e9), flush=True)
sub = np.unique(np.linspace(2, nf - 1, 64).astype(np.int32))
chk = Reference() if have_reference() else Oracle()
t0 = time.time()
a2, s2 = chk.rmsavgcorr(crd, sel, sub)
dt = time.time() - t0
it2 = float(np.sum(nf - sub + 1))
print("CPU (%s, %d threads): %d window sizes, %.3e RMSDs in %.2f s = %.3e /s" % (type(chk).__name__, chk.threads(), len(sub), it2, dt, it2 / dt))
print("max |avg diff| %.2e, max |sd diff| %.2e" % (np.abs(avg[sub - 1] - a2).max(), np.abs(sd[sub - 1] - s2).max()))
