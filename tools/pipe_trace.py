"""Time line of the pipelined host path (B200_PIPE_TRACE): pinned COORDS in, pinned triangle out, cfg2."""
import os, sys, time
os.environ["B200_PIPE_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpptraj_b200 as b
from cpptraj_b200.synth import make_trajectory
nF, nA = 10000, 1000
b.init(1)
crd, _ = make_trajectory(20261017, nF, nA)
sel = np.arange(nA, dtype=np.int32)
pin_in = torch.from_numpy(crd).pin_memory()
pin_out = torch.empty(nF * (nF - 1) // 2, dtype=torch.float32).pin_memory()
for r in range(4):
    t0 = time.perf_counter()
    b.rms2d_tri(pin_in.numpy(), sel, out=pin_out.numpy())
    print("call %d: %.3f ms" % (r, 1e3 * (time.perf_counter() - t0)), file=sys.stderr, flush=True)
