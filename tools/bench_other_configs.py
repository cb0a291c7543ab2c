"""Measurements of the BASELINE configs that are not the bench.py line (one GPU): prints one JSON line per config.

  cfg3  rmsd one-vs-many, 1,000,000 frames x 5,000 atoms to a reference, streaming: a 20,000-frame block of
        synthetic COORDS (1.2 GB, pinned) is pushed 50 times through b200_rmsd_1vN_push_f32 (60 GB over PCIe),
        and the same block is processed device-resident (b200_dev_rmsd_1vN) for the HBM roofline
        (12*N bytes per frame).
  cfg4  cluster pairwise cache, Metric_RMS mass-weighted, 50,000 frames x 2,000 atoms (1,249,975,000 pairs),
        device-resident, one GPU (the 8-GPU figure is this rate times the shard count: no collective).
usage: python tools/bench_other_configs.py [cfg3] [cfg4] [--small]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpptraj_b200 as b
from cpptraj_b200.synth import make_trajectory, masses

small = "--small" in sys.argv
which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["cfg3", "cfg4"]
b.init(1)
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass


def events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


if "cfg3" in which:
    nA, nBlock, reps = 5000, (2000 if small else 20000), (5 if small else 50)
    h = torch.empty((nBlock, 3 * nA), dtype=torch.float32, pin_memory=True)
    make_trajectory(20261018, nBlock, nA, out=h.numpy())
    sel = np.arange(nA, dtype=np.int32)
    ref = h.numpy()[0].reshape(-1, 3).astype(np.float64)
    ref -= ref.mean(0)
    # ---- device-resident kernel rate
    d = h.cuda(); d_sel = torch.from_numpy(sel).cuda(); d_ref = torch.from_numpy(ref.reshape(-1)).cuda()
    d_out = torch.empty(nBlock, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        b.dev_rmsd_1vN(d, 3 * nA, nBlock, d_sel, nA, d_ref, d_out, stream=st)
    e0, e1 = events(); e0.record()
    K = 10
    for _ in range(K):
        b.dev_rmsd_1vN(d, 3 * nA, nBlock, d_sel, nA, d_ref, d_out, stream=st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    dev_fps = nBlock / ms * 1e3
    dev_gbs = dev_fps * 12.0 * nA / 1e9
    r_dev = d_out.cpu().numpy()
    del d, d_out
    # ---- streaming through the C ABI from pinned host memory
    hd = b.Rmsd1vN(ref, sel, fit=True, want_rot=False)
    with b.Rmsd1vN(ref, sel, fit=True, want_rot=False) as wu:   # warm-up (allocations, first launch) on its own handle
        wu.push(h.numpy()[: min(nBlock, 512)]); wu.flush()
    t0 = time.perf_counter()
    rs = []
    best, bestv, base = -1, np.inf, 512 if nBlock >= 512 else nBlock
    for _ in range(reps):
        hd.push(h.numpy())
        r_k, _, _, best_k = hd.flush()                          # flush per block: the result array stays block-sized
        rs.append(r_k)
        best = best_k
    hd.close()
    r = rs[0]
    pass
    dt = time.perf_counter() - t0
    frames = nBlock * reps
    line = {"config": "cfg3: rmsd one-vs-many fit, %d frames x %d atoms, streaming (block of %d frames pushed %d times)" % (frames, nA, nBlock, reps),
            "metric": "frames/s", "e2e": {"value": frames / dt, "unit": "frames/s", "GB/s_h2d": frames * 12.0 * nA / dt / 1e9, "seconds": dt,
                                           "api": "b200_rmsd_1vN_push_f32 / flush, pinned host COORDS"},
            "value": dev_fps, "unit": "frames/s (device-resident)", "ms_per_block": ms,
            "roofline": {"bound": "hbm", "achieved": dev_gbs, "unit": "GB/s", "peak": peaks.get("hbm_gbs"),
                         "frac": dev_gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None, "bytes_per_frame": 12 * nA,
                         "kernel": "onevn_stream_kernel<float> + onevn_finish_kernel"},
            "argmin_frame": int(best), "argmin_rmsd": float(np.min(r)), "checksum": float(np.sum(r[:nBlock])),
            "device_vs_stream_max_abs_diff": float(np.abs(r[:nBlock] - r_dev).max())}
    print(json.dumps(line), flush=True)

if "cfg4" in which:
    nF, nA = (8000 if small else 50000), 2000
    h = torch.empty((nF, 3 * nA), dtype=torch.float32, pin_memory=True)
    t0 = time.perf_counter()
    make_trajectory(20261019, nF, nA, out=h.numpy())
    tgen = time.perf_counter() - t0
    m = masses(nA)
    d = h.cuda(); d_sel = torch.arange(nA, dtype=torch.int32, device="cuda"); d_m = torch.from_numpy(m).cuda()
    npairs = nF * (nF - 1) // 2
    d_out = torch.empty(npairs, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        b.dev_rms2d_tri(d, 3 * nA, nF, d_sel, nA, d_out, d_mass=d_m, fit=True, stream=st)
    torch.cuda.synchronize()
    b.set_profiling(True); b.reset_stats()
    e0, e1 = events(); e0.record()
    K = 3
    for _ in range(K):
        b.dev_rms2d_tri(d, 3 * nA, nF, d_sel, nA, d_out, d_mass=d_m, fit=True, stream=st)
    e1.record(); torch.cuda.synchronize()
    s = b.get_stats(); b.set_profiling(False)
    ms = e0.elapsed_time(e1) / K
    eng, qs = b.last_pair_engine()
    # (parity of this configuration: tests/test_gpu_i8.py::test_i8_config4_like_mass_weighted -- the oracle is test-only)
    line = {"config": "cfg4: cluster pairwise cache (Metric_RMS, mass-weighted), %d frames x %d atoms, %d pairs, 1 GPU, device-resident" % (nF, nA, npairs),
            "metric": "pair-RMSDs/s", "value": npairs / ms * 1e3, "unit": "pair-RMSDs/s", "ms_per_step": ms,
            "engine": {1: "fp64-dmma", 2: "tcgen05-int8"}.get(eng), "fixed_point_fraction_bits": qs,
            "pair_kernel_ms": s["pair_ms"] / K, "pack_ms": s["pack_ms"] / K,
            "algorithmic_TFLOP/s": 18.0 * nA * npairs / (s["pair_ms"] / K * 1e-3) / 1e12,
            "checksum": float(d_out[: 1 << 20].double().sum().item()), "gen_seconds": round(tgen, 1)}
    print(json.dumps(line), flush=True)
b.shutdown()
