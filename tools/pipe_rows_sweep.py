"""e2e time of the pipelined host path (pinned in / pinned out, cfg2) against B200_PIPE_ROWS."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpptraj_b200 as b
from cpptraj_b200.synth import make_trajectory
nF, nA = 10000, 1000
crd, _ = make_trajectory(20261017, nF, nA)
sel = np.arange(nA, dtype=np.int32)
pin_in = torch.from_numpy(crd).pin_memory()
pin_out = torch.empty(nF * (nF - 1) // 2, dtype=torch.float32).pin_memory()
for rows in sys.argv[1:] or ["0"]:
    if rows == "0":
        os.environ.pop("B200_PIPE_ROWS", None)
    else:
        os.environ["B200_PIPE_ROWS"] = rows
    b.init(1)
    ts = []
    for r in range(8):
        t0 = time.perf_counter()
        b.rms2d_tri(pin_in.numpy(), sel, out=pin_out.numpy())
        ts.append(time.perf_counter() - t0)
    ts = sorted(ts[2:])
    print("B200_PIPE_ROWS=%s: best %.2f ms median %.2f ms" % (rows, ts[0] * 1e3, ts[len(ts) // 2] * 1e3), flush=True)
    b.shutdown()
