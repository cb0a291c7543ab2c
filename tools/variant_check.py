"""Same-box comparison of libb200rmsd.so builds (B200_RMSD_LIB selects the build): parity of the tcgen05 engine against
the FP64 engine on a cfg2 prefix, device-resident cfg2 timing, per-CTA cycle counters.
usage: [B200_RMSD_LIB=variants/x.so] python tools/variant_check.py [frames] [atoms] [--clocks]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpptraj_b200 as b
from cpptraj_b200.synth import make_trajectory
args = [a for a in sys.argv[1:] if not a.startswith("--")]
nF = int(args[0]) if len(args) > 0 else 10000
nA = int(args[1]) if len(args) > 1 else 1000
tag = os.path.basename(os.environ.get("B200_RMSD_LIB", "default"))
b.init(1)
crd, _ = make_trajectory(20261017, nF, nA)
d_crd = torch.from_numpy(crd).cuda(); d_sel = torch.arange(nA, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
# ---- parity: tcgen05 engine vs FP64 engine on the first 2800 frames
nP = min(nF, 2800)
outs = []
for eng in ("auto", "fp64"):
    b.set_pair_engine(eng)
    o = torch.zeros(nP * (nP - 1) // 2, dtype=torch.float32, device="cuda")
    b.dev_rms2d_tri(d_crd, 3 * nA, nP, d_sel, nA, o, fit=True, stream=st)
    torch.cuda.synchronize()
    outs.append(o)
b.set_pair_engine("auto")
diff = (outs[0].double() - outs[1].double()).abs()
print("%s parity vs fp64 engine on %d frames: max %.3e mean %.3e (engine %s)" % (tag, nP, diff.max().item(), diff.mean().item(), b.last_pair_engine()))
del outs, diff
# ---- timing
d_out = torch.empty(nF * (nF - 1) // 2, dtype=torch.float32, device="cuda")
def step(): b.dev_rms2d_tri(d_crd, 3 * nA, nF, d_sel, nA, d_out, fit=True, stream=st)
for _ in range(3): step()
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    b.set_profiling(True); b.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): step()
    e1.record(); torch.cuda.synchronize()
    s = b.get_stats(); b.set_profiling(False)
    ms = e0.elapsed_time(e1) / 5
    best = min(best, ms)
    print("%s rep %d: %.3f ms/step  %.3e pairs/s  pair kernels %.3f ms/step  pack %.3f ms/step" % (
        tag, rep, ms, nF * (nF - 1) / 2 / ms * 1e3, s["pair_ms"] / 5, s["pack_ms"] / 5))
print("%s BEST %.3f ms/step %.3e pairs/s" % (tag, best, nF * (nF - 1) / 2 / best * 1e3))
if "--clocks" in sys.argv:
    import ctypes as C
    L = b.lib(); L.b200_debug_i8_clocks.argtypes = [C.c_void_p, C.c_int]
    L.b200_debug_i8_clocks(None, 0)
    step(); torch.cuda.synchronize()
    buf = np.zeros((148, 16), np.int64)
    L.b200_debug_i8_clocks(buf.ctypes.data_as(C.c_void_p), 148)
    names = ["prod_wait_empty", "mma_wait_accEmpty", "mma_wait_full", "mma_total", "mma_tiles",
             "drain_wait_accFull", "drain_wait_xEmpty", "drain_total", "solve_wait_xFull", "solve_total", "drain_tmem_ld", "drain_fold", "solve_window", "solve_fp64", "mma_wait_fpDone", "solve_root"]
    tiles = buf[:, 4][buf[:, 4] != 0].mean()
    for k, nm in enumerate(names):
        col = buf[:, k]; nz = col[col != 0]
        print("  %-18s mean %12.0f  per tile %8.0f  (%d CTAs)" % (nm, nz.mean() if len(nz) else 0, (nz.mean() / tiles) if len(nz) else 0, len(nz)))
