"""Is pair_i8_kernel clock-for-clock or power bound?  Runs the device-resident cfg2 pass back to back for ~2 s per
timing mode of the kernel (B200_I8_DEBUG_MODE: 0 product, 7 tensor SMs only export the covariance -- no FP64 window, no
per-pair solve --, 2 MMAs + operand stream alone) while NVML samples SM clock and board power every 10 ms, and reads the
kernel's own cycle counters (clock64 in the MMA warp) for one step next to its CUDA-event time: cycles / time = the SM
clock the kernel actually ran at.  Last: the tcgen05 kind::i8 issue probe (operands fixed in shared memory) sustained.
usage (GPU box): [B200_PROBE_MODES=0,7,2] python tools/power_probe.py [frames] [atoms] [seconds]"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpptraj_b200 as b
from cpptraj_b200.synth import make_trajectory
from bench import ClockSampler
nF = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nA = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
b.init(1)
crd, _ = make_trajectory(20261017, nF, nA)
d_crd = torch.from_numpy(crd).cuda(); d_sel = torch.arange(nA, dtype=torch.int32, device="cuda")
d_out = torch.empty(nF * (nF - 1) // 2, dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
pairs = nF * (nF - 1) / 2
def step(): b.dev_rms2d_tri(d_crd, 3 * nA, nF, d_sel, nA, d_out, fit=True, stream=st)
L = b.lib(); L.b200_debug_i8_clocks.argtypes = [C.c_void_p, C.c_int]
def timed(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for mode in [int(x) for x in os.environ.get("B200_PROBE_MODES", "0,7,2").split(",")]:
    os.environ["B200_I8_DEBUG_MODE"] = str(mode)
    for _ in range(3): step()
    torch.cuda.synchronize(); time.sleep(0.5)           # let the board cool to idle clocks / power
    burst = timed(5)
    time.sleep(0.5)
    n = max(20, int(secs * 1e3 / burst))
    smp = ClockSampler(0); smp.start()
    sustained = timed(n)
    tail = timed(20)
    r = smp.result()
    sm = np.array(smp.sm, float); pw = np.array(smp.power, float)
    busy = pw > 0.5 * pw.max() if len(pw) else np.zeros(0, bool)
    # cycle counters of one step, right after the sustained loop (same thermal / power state)
    b.set_profiling(True); b.reset_stats()
    L.b200_debug_i8_clocks(None, 0)
    step(); torch.cuda.synchronize()
    buf = np.zeros((148, 16), np.int64)
    L.b200_debug_i8_clocks(buf.ctypes.data_as(C.c_void_p), 148)
    s = b.get_stats(); b.set_profiling(False)
    mma = buf[:, 3][buf[:, 3] != 0]; tiles = buf[:, 4][buf[:, 4] != 0]
    cyc = float(mma.max()) if len(mma) else 0.0
    print("mode %d: burst(5 steps) %.3f ms/step  sustained(%d steps) %.3f ms/step  last 20: %.3f ms/step = %.3e pairs/s" % (
        mode, burst, n, sustained, tail, pairs / tail * 1e3))
    print("   NVML under load: sm clock median %.0f min %.0f MHz (max %s), power mean %.0f max %.0f W, reasons %s, %d samples" % (
        np.median(sm[busy]) if busy.any() else -1, sm[busy].min() if busy.any() else -1, r.get("sm_max_mhz"),
        pw[busy].mean() if busy.any() else -1, pw.max() if len(pw) else -1, r.get("reasons"), len(sm)))
    print("   one step with cycle counters: pair kernels %.3f ms, MMA warp %.0f cycles (%.0f per tile, %.0f tiles) -> %.0f MHz effective in the kernel" % (
        s["pair_ms"], cyc, cyc / tiles.mean() if len(tiles) else 0, tiles.mean() if len(tiles) else 0, cyc / s["pair_ms"] / 1e3 if s["pair_ms"] else 0))
os.environ["B200_I8_DEBUG_MODE"] = "0"
time.sleep(0.5)
print("tcgen05 kind::i8 issue probe (5 x 1.3 ms bursts per call, best): first call %.0f TOP/s" % b.measure_i8_mma_peak(0))
smp = ClockSampler(0); smp.start()
t0 = time.time(); vals = []
while time.time() - t0 < secs: vals.append(b.measure_i8_mma_peak(0))
r = smp.result()
sm = np.array(smp.sm, float); pw = np.array(smp.power, float)
print("   sustained for %.1f s: last call %.0f TOP/s, min %.0f; sm clock median %.0f MHz, power max %.0f W, reasons %s" % (
    secs, vals[-1], min(vals), np.median(sm), pw.max(), r.get("reasons")))
