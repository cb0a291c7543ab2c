"""End-to-end time of b200_rms2d_tri (pinned COORDS in, pinned triangle out, cfg2 unless [frames] [atoms]) for A/B runs of
the host pipeline's knobs (environment, one process per setting).  usage: [B200_...=..] python tools/e2e_pipe_ab.py [frames] [atoms] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpptraj_b200 as b
from cpptraj_b200.synth import make_trajectory
nF = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nA = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
b.init(1)
crd, _ = make_trajectory(20261017, nF, nA)
sel = np.arange(nA, dtype=np.int32)
nT = nF * (nF - 1) // 2
pin_in = torch.from_numpy(crd).pin_memory().numpy()
pin_out = torch.empty(nT, dtype=torch.float32).pin_memory().numpy()
ts = []
for r in range(reps + 3):
    t0 = time.perf_counter()
    b.rms2d_tri(pin_in, sel, out=pin_out)
    ts.append(time.perf_counter() - t0)
ts = sorted(ts[3:])
knobs = " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("B200_") and k != "B200_PIPE_TRACE")
print("%-50s best %.3f ms  median %.3f ms  -> %.3e pairs/s  checksum %.6f" % (knobs or "(defaults)", ts[0] * 1e3, ts[len(ts) // 2] * 1e3,
      nT / ts[len(ts) // 2], float(pin_out[::997].astype(np.float64).sum())), flush=True)
