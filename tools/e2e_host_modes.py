"""End-to-end time of b200_rms2d_tri on cfg2 (or [frames] [atoms]) for the kinds of host buffers a caller may pass:
pinned (bench.py), pageable COORDS (cpptraj's std::vector<float>), pageable result -- touched, or fresh as cpptraj's
new float[] is.  usage: python tools/e2e_host_modes.py [frames] [atoms] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpptraj_b200 as b
from cpptraj_b200.synth import make_trajectory
nF = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nA = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
b.init(1)
crd, _ = make_trajectory(20261017, nF, nA)
sel = np.arange(nA, dtype=np.int32)
nT = nF * (nF - 1) // 2
pin_in = torch.from_numpy(crd).pin_memory()
pin_out = torch.empty(nT, dtype=torch.float32).pin_memory()
page_out = np.empty(nT, np.float32); page_out[:] = 0      # touched
ref = None
def run(name, cin, mk_out):
    global ref
    ts = []
    for r in range(reps + 2):
        out = mk_out()
        t0 = time.perf_counter()
        b.rms2d_tri(cin, sel, out=out)
        ts.append(time.perf_counter() - t0)
        if ref is None: ref = out.copy()
        elif r == 0: assert np.array_equal(out, ref), name
        del out
    ts = sorted(ts[2:])
    print("%-46s best %.2f ms  median %.2f ms  -> %.3e pairs/s" % (name, ts[0] * 1e3, ts[len(ts) // 2] * 1e3, nT / ts[len(ts) // 2]), flush=True)
run("pinned COORDS   -> pinned triangle", pin_in.numpy(), lambda: pin_out.numpy())
run("pageable COORDS -> pinned triangle", crd, lambda: pin_out.numpy())
run("pageable COORDS -> pageable, touched", crd, lambda: page_out)
run("pageable COORDS -> pageable, fresh (cpptraj)", crd, lambda: np.empty(nT, np.float32))
for t in (1, 4, 8):
    os.environ["B200_HOST_THREADS"] = str(t)
    b.shutdown(); b.init(1)
    run("  same, B200_HOST_THREADS=%d" % t, crd, lambda: np.empty(nT, np.float32))
