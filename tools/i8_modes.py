"""Timing experiments for the tcgen05 pair kernel: device-resident cfg2, kernel time by CUDA events.
usage: B200_I8_DEBUG_MODE=k python tools/i8_modes.py [frames] [atoms]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpptraj_b200 as b
from cpptraj_b200.synth import make_trajectory
nF = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nA = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
b.init(1)
crd, _ = make_trajectory(20261017, nF, nA)
d_crd = torch.from_numpy(crd).cuda(); d_sel = torch.arange(nA, dtype=torch.int32, device="cuda")
d_out = torch.empty(nF * (nF - 1) // 2, dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def step(): b.dev_rms2d_tri(d_crd, 3 * nA, nF, d_sel, nA, d_out, fit=True, stream=st)
for _ in range(3): step()
torch.cuda.synchronize()
b.set_profiling(True); b.reset_stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
s = b.get_stats()
ms = e0.elapsed_time(e1) / 5
print("mode %s engine %s: %.3f ms/step  %.3e pairs/s  pair kernels %.3f ms/step  pack %.3f ms/step" % (
    os.environ.get("B200_I8_DEBUG_MODE", "0"), b.last_pair_engine(), ms, nF * (nF - 1) / 2 / ms * 1e3, s["pair_ms"] / 5, s["pack_ms"] / 5))
if os.environ.get("B200_I8_CLOCKS"):
    import ctypes as C
    L = b.lib(); L.b200_debug_i8_clocks.argtypes = [C.c_void_p, C.c_int]
    b.set_profiling(False)
    L.b200_debug_i8_clocks(None, 0)
    step(); torch.cuda.synchronize()
    buf = np.zeros((148, 16), np.int64)
    L.b200_debug_i8_clocks(buf.ctypes.data_as(C.c_void_p), 148)
    names = ["prod_wait_empty", "mma_wait_accEmpty", "mma_wait_full", "mma_total", "mma_tiles",
             "drain_wait_accFull", "drain_wait_xEmpty", "drain_total", "solve_wait_xFull", "solve_total", "drain_tmem_ld", "drain_fold", "solve_window", "solve_fp64", "mma_wait_fpDone", "solve_root"]
    # counters accumulate over the 3 band launches of one step (last launch overwrites: use sums per CTA of last launch only)
    # with CTA pairs the MMA counters exist on the leader CTAs only: average over the CTAs that counted
    for k, nm in enumerate(names):
        col = buf[:, k]; nz = col[col != 0]
        print("  %-18s mean %12.0f  max %12.0f  (%d CTAs)" % (nm, nz.mean() if len(nz) else 0, col.max(), len(nz)))
