#!/usr/bin/env python
"""bench.py -- the B200 best-fit RMSD path on BASELINE.json's configs (metric: fitted pair-RMSDs/s of rms2d FxN).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2|cfg3|cfg4|cfg5] [--impl reference] [--no-extras]
    torchrun --nproc-per-node N bench.py --gpus N ...           (one rank per GPU)

Workloads (synthetic, cpptraj_b200.synth, seeds from SURVEY.md 8d):
    cfg2  rms2d fit, 10,000 frames x 1,000 atoms (49,995,000 pairs)                         default at N == 1
    cfg5  rms2d fit, 100,000 frames x 1,000 atoms, the upper triangle sharded into          default at N  > 1
          contiguous, pair-balanced row bands over the ranks (strong scaling, no data-path
          collective: every rank packs the frames it needs itself)
    cfg4  cluster pairwise cache: mass-weighted fit, 50,000 frames x 2,000 atoms, sharded like cfg5
    cfg3  rmsd one-vs-many: 1,000,000 frames x 5,000 atoms to a reference (a 20,000-frame block repeated 50 times),
          frames sharded over the ranks; metric frames/s (and GB/s against HBM)
A "step" is one pass over all pairs (frames) of the rank's shard.

One JSON line is printed by rank 0:
    value      whole job, inputs (raw float32 COORDS) resident in HBM, CUDA events on the launching stream
    e2e        same metric through the C ABI with pinned HOST buffers: H2D of the COORDS, kernels, D2H of the result inside
               the timed region; `host_ceiling` = what N ranks copying device -> pinned host at once reach on this box
               (measured here), `frac_of_d2h_ceiling` = the run's own D2H rate against it; `host_ceiling.duplex_ms` = the time
               the box needs to move one step's H2D and D2H bytes at the same time with nothing else running (the two
               directions of the link slow each other down) and `copy_floor_frac` = that floor / the step's e2e time;
               `pageable` = the same call with the pageable / never-touched buffers cpptraj passes (N == 1)
    roofline   dominant kernel (tcgen05 int8 pair kernel when eligible, else FP64 DMMA; one-vs-many streaming kernel for
               cfg3): algorithmic work / CUDA-event kernel time against a peak measured live by a probe kernel
    parity     every rank checks sampled rows of ITS OWN band (24 rows x up to 2,048 columns) of the e2e result against the
               reference's own Frame::RMSD_CenteredRef (oracle/_ref); max over ranks; the ranks agree on one fixed-point
               grid (min over ranks of the fractional bits, pinned with b200_set_fixed_point_bits)
    cpu_baseline  the reference's own pair loop (oracle/_ref, OpenMP, all host threads) on a bounded prefix (rank 0)
    configs    (default run only) the other BASELINE configs as short legs: cfg5 on this GPU count when the main line is
               cfg2 (one workload for the 1 -> 8 GPU curve), cfg4, cfg3, and `cpptraj`: cpptraj.B200 and the unmodified
               cpptraj.OMP timed by their own "TIME: Analyses took" lines on the same binpos file

--impl reference times the reference's CPU implementation alone (rank 0 only under torchrun).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fitted pair-RMSDs/sec (rms2d FxN)"
UNIT = "pair-RMSDs/s"
CONFIGS = {
    "cfg2": dict(kind="tri", frames=10000, atoms=1000, mass=False, seed=20261017,
                 what="rms2d fit, 10000 frames x 1000 atoms"),
    "cfg4": dict(kind="tri", frames=50000, atoms=2000, mass=True, seed=20261019,
                 what="cluster pairwise cache (Metric_RMS, mass-weighted fit), 50000 frames x 2000 atoms"),
    "cfg5": dict(kind="tri", frames=100000, atoms=1000, mass=False, seed=20261020,
                 what="rms2d fit, 100000 frames x 1000 atoms"),
    "cfg3": dict(kind="1vn", frames=1000000, atoms=5000, block=20000, seed=20261018,
                 what="rmsd one-vs-many, 1000000 frames x 5000 atoms (20000-frame block x 50)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="auto", choices=["auto"] + sorted(CONFIGS))
    ap.add_argument("--frames", type=int, default=0, help="override frame count (custom workload)")
    ap.add_argument("--atoms", type=int, default=0)
    ap.add_argument("--no-extras", action="store_true", help="main line only: no legs for the other BASELINE configs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU seconds for the bounded baseline sample")
    return ap.parse_args()


def main_config(args):
    if args.config != "auto":
        name = args.config
    else:
        name = "cfg5" if args.gpus > 1 else "cfg2"
    cfg = dict(CONFIGS[name])
    if args.frames:
        name, cfg["frames"] = "custom", args.frames
    if args.atoms:
        cfg["atoms"] = args.atoms
    return name, cfg


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_evt = threading.Event()
        self.sm, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
             0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
             0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self.stop_evt.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.NAMES.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self.stop_evt.wait(0.01)

    def result(self):
        self.stop_evt.set()
        if self.is_alive():
            self.join(2.0)
        if not self.ok or not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "nvml unavailable"}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "power_w_max": max(self.power) if self.power else None}


# ----------------------------------------------------------------------------- synthetic data
def gen_trajectory(seed, nF, nA, out):
    """cpptraj_b200.synth's generator, chunks on a thread pool (numpy releases the GIL): frame f is a random rigid motion
    of a 3.8 A random-walk chain + N(0, sigma_f^2) noise, sigma_f cycling over {0, 0.05, 0.5, 2} A, every 64th frame a
    bit-exact duplicate of its predecessor.  Deterministic for (seed, nF, nA)."""
    from concurrent.futures import ThreadPoolExecutor
    from cpptraj_b200.synth import SIGMAS, _rotations, base_chain
    base = base_chain(np.random.default_rng(seed), nA)
    chunk = 2048

    def job(f0):
        f1 = min(nF, f0 + chunk)
        n = f1 - f0
        rng = np.random.default_rng([seed, f0])
        R = _rotations(rng, n)
        T = rng.uniform(-20.0, 20.0, (n, 1, 3))
        sig = np.array([SIGMAS[f % len(SIGMAS)] for f in range(f0, f1)], np.float32)[:, None, None]
        xyz = (np.einsum("fij,aj->fai", R, base) + T).astype(np.float32)
        xyz += sig * rng.standard_normal((n, nA, 3), dtype=np.float32)
        out[f0:f1, :3 * nA] = xyz.reshape(n, -1)

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        list(ex.map(job, range(0, nF, chunk)))
    dup = np.arange(63, nF, 64)
    out[dup] = out[dup - 1]


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_impl(threads=None):
    from oracle.pyoracle import Oracle, Reference, have_reference
    impl, kind = (Reference(), "reference") if have_reference() else (Oracle(), "port")
    cores = threads or host_cores()
    impl.set_threads(cores)
    return impl, kind, cores


def cpu_tri_baseline(crd, sel, mass, target_seconds):
    """The reference's own pair loop on a bounded prefix of crd: dict(value, cores, kind, sample) + the prefix size."""
    impl, kind, cores = cpu_impl()

    def run(nf):
        t0 = time.perf_counter()
        impl.rms2d_tri(crd[:nf], sel, mass=mass)
        dt = time.perf_counter() - t0
        return impl.last_loop_seconds() if kind == "reference" else dt   # pair loop only, what cpptraj's TIME: line covers

    nf = min(crd.shape[0], 400)
    rate = nf * (nf - 1) / 2 / max(run(nf), 1e-4)
    nf2 = int(min(crd.shape[0], max(nf, (2 * rate * target_seconds) ** 0.5)))
    dt = run(nf2)
    return {"value": nf2 * (nf2 - 1) / 2 / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "rms2d %sfit on the first %d of %d frames x %d atoms (%d pairs), %.1f s" % (
                "mass-weighted " if mass is not None else "", nf2, crd.shape[0], len(sel), nf2 * (nf2 - 1) // 2, dt)}


# ----------------------------------------------------------------------------- CPU reference arm
def reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from cpptraj_b200.synth import masses
    name, cfg = main_config(args)
    nF, nA = cfg["frames"], cfg["atoms"]
    steps_total = max(1, args.steps + args.warmup)
    per_step = max(2.0, min(20.0, 150.0 / steps_total))
    impl, kind, cores = cpu_impl()
    if cfg["kind"] == "1vn":
        nGen = min(cfg["block"], 4000)
        crd = np.empty((nGen, 3 * nA), np.float32)
        gen_trajectory(cfg["seed"], nGen, nA, crd)
        sel = np.arange(nA, dtype=np.int32)
        ref = crd[0].reshape(-1, 3).astype(np.float64)
        t0 = time.perf_counter(); impl.rmsd_1vN(crd[:200], sel, ref); rate = 200 / max(time.perf_counter() - t0, 1e-4)
        nf = int(min(nGen, max(200, rate * per_step)))

        def step():
            t0 = time.perf_counter()
            impl.rmsd_1vN(crd[:nf], sel, ref)
            return time.perf_counter() - t0
        units, unit, metric = nf, "frames/s", "one-vs-many fitted RMSD frames/sec (rmsd action)"
        sample = "rmsd to a fixed reference, %d frames x %d atoms per step, %d OpenMP threads" % (nf, nA, cores)
    else:
        nGen = min(nF, 12000 if nA <= 1000 else 6000)
        crd = np.empty((nGen, 3 * nA), np.float32)
        gen_trajectory(cfg["seed"], nGen, nA, crd)
        sel = np.arange(nA, dtype=np.int32)
        mass = masses(nA) if cfg["mass"] else None
        t0 = time.perf_counter(); impl.rms2d_tri(crd[:300], sel, mass=mass)
        rate = 300 * 299 / 2 / max(time.perf_counter() - t0, 1e-4)
        nf = int(min(nGen, max(300, (2 * rate * per_step) ** 0.5)))

        def step():
            t0 = time.perf_counter()
            impl.rms2d_tri(crd[:nf], sel, mass=mass)
            dt = time.perf_counter() - t0
            return impl.last_loop_seconds() if kind == "reference" else dt
        units, unit, metric = nf * (nf - 1) / 2, UNIT, METRIC
        sample = "rms2d fit on the first %d of %d frames x %d atoms (%d pairs/step), %d OpenMP threads" % (nf, nF, nA, units, cores)
    for _ in range(args.warmup):
        step()
    tot = sum(step() for _ in range(args.steps))
    value = units * args.steps / tot
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s (the CPU arm times a bounded sample per step: %s)" % (name, cfg["what"], sample)},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- B200 arm
class Ctx:
    """Process-wide state of the B200 arm: ranks, torch.distributed plumbing, the library."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import cpptraj_b200 as b
        self.torch, self.dist, self.b, self.args = torch, dist, b, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit("WORLD_SIZE %d != --gpus %d" % (self.world, args.gpus))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback "
                             "(use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        b.init(devices=[self.local])
        self.peaks = {}
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.stream = torch.cuda.current_stream().cuda_stream

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x, op="max"):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        ops = {"max": self.dist.ReduceOp.MAX, "min": self.dist.ReduceOp.MIN, "sum": self.dist.ReduceOp.SUM}
        self.dist.all_reduce(t, op=ops[op])
        return float(t.item())


def host_ceiling(ctx, nbytes, h2d_bytes=0, d2h_bytes=0):
    """All ranks copy device -> pinned host (then host -> device) at the same time: aggregate GB/s of this box."""
    torch = ctx.torch
    n = int(min(nbytes, 2 << 30))
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    pin = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    out = {}
    for name, fn in (("d2h", lambda: pin.copy_(dev, non_blocking=True)), ("h2d", lambda: dev.copy_(pin, non_blocking=True))):
        best = 1e9
        for _ in range(3):
            ctx.barrier()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            best = min(best, ctx.reduce(time.perf_counter() - t0))
        out[name] = ctx.reduce(float(n), "sum") / best / 1e9
    del dev, pin
    torch.cuda.empty_cache()
    # both directions at once, with this rank's own byte counts of one step: the time the box needs merely to move a
    # step's inputs in and its results out over a full-duplex link (the two directions slow each other down)
    if h2d_bytes and d2h_bytes:
        hn, dn = int(min(h2d_bytes, 4 << 30)), int(min(d2h_bytes, 4 << 30))
        d_in = torch.empty(hn, dtype=torch.uint8, device="cuda"); p_in = torch.empty(hn, dtype=torch.uint8, pin_memory=True)
        d_out = torch.empty(dn, dtype=torch.uint8, device="cuda"); p_out = torch.empty(dn, dtype=torch.uint8, pin_memory=True)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        best = 1e9
        for _ in range(3):
            ctx.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            with torch.cuda.stream(s1):
                d_in.copy_(p_in, non_blocking=True)
            with torch.cuda.stream(s2):
                p_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, ctx.reduce(time.perf_counter() - t0))
        out["duplex_ms"] = 1e3 * best * max(h2d_bytes / hn, d2h_bytes / dn)   # (probe buffers are capped at 4 GB each)
        del d_in, p_in, d_out, p_out
        torch.cuda.empty_cache()
    return out


def tri_bench(ctx, name, cfg, steps, warmup, cpu_seconds, light=False):
    """One rms2d / pairwise-cache triangle workload on all ranks.  light: a short leg (no FP64-engine comparison, no
    pageable leg, shorter CPU sample)."""
    import ctypes as C
    from cpptraj_b200.synth import masses
    torch, b = ctx.torch, ctx.b
    world, rank, local = ctx.world, ctx.rank, ctx.local
    nF, nA = cfg["frames"], cfg["atoms"]
    stride = 3 * nA
    mass = masses(nA) if cfg["mass"] else None
    t_gen = time.perf_counter()
    h_crd = torch.empty((nF, stride), dtype=torch.float32, pin_memory=True)
    gen_trajectory(cfg["seed"], nF, nA, h_crd.numpy())     # every rank generates the same trajectory
    t_gen = time.perf_counter() - t_gen
    sel = np.arange(nA, dtype=np.int32)
    r0, r1 = b.shard_rows(nF, rank, world)
    first = nF * r0 - r0 * (r0 + 1) // 2
    nelt = (nF * r1 - r1 * (r1 + 1) // 2) - first
    total_pairs = nF * (nF - 1) // 2

    # =================== value: device-resident ===================
    d_crd = h_crd.cuda(non_blocking=False)
    d_sel = torch.from_numpy(sel).cuda()
    d_mass = torch.from_numpy(mass).cuda() if mass is not None else None
    d_out = torch.empty(max(nelt, 1), dtype=torch.float32, device="cuda")
    d_out_base = d_out.data_ptr() - 4 * first      # the ABI indexes the whole triangle: base[first] is d_out[0]

    def dev_step():
        b.dev_rms2d_tri(d_crd, stride, nF, d_sel, nA, d_out_base, d_mass=d_mass, fit=True, rank=rank, count=world, stream=ctx.stream)

    def timed_device_run(k, w):
        b.set_profiling(False)
        for _ in range(w):
            dev_step()
        ctx.barrier()
        b.set_profiling(True)
        b.reset_stats()
        sampler = ClockSampler(local)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        e0.record()
        for _ in range(k):
            dev_step()
        e1.record()
        ctx.barrier()
        own = e0.elapsed_time(e1)
        ms = ctx.reduce(own)
        clk = sampler.result()
        stt = b.get_stats()
        b.set_profiling(False)
        return ms, own, stt, clk

    def roofline_of(stt, own_ms, engine):
        launches = max(1, stt["pair_launches"])
        ms_per_launch = stt["pair_ms"] / launches
        flop_per_launch = 18.0 * nA * (stt["pairs"] / launches)       # algorithmic: SURVEY 8(d), 18*N per fitted pair
        achieved = flop_per_launch / (ms_per_launch * 1e-3) / 1e12 if ms_per_launch > 0 else 0.0
        r = {"bound": "tensor", "achieved": achieved, "unit": "TFLOP/s", "launches": int(stt["pair_launches"]),
             "avg_launch_ms": ms_per_launch, "kernel_share_of_step": stt["pair_ms"] / own_ms if own_ms > 0 else None}
        # dram__bytes_read + dram__bytes_write of one launch, from the committed ncu --set full capture of cfg2 at N = 1
        # (profiles/r2_pair_i8_kernel_ncu_raw.csv); not re-measured per run, null for other shapes
        if engine == 2 and name == "cfg2" and world == 1:
            r["traffic"] = 101.7e6
            r["traffic_source"] = "from_profile: profiles/r2_pair_i8_kernel_ncu_raw.csv (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"
        else:
            r["traffic"] = None
        bf16 = ctx.peaks.get("bf16_tflops_sustained")
        bf16_burst = ctx.peaks.get("bf16_tflops")
        if engine == 2:
            kpad = (nA + 63) // 64 * 64
            executed = achieved * 9.0 * kpad / nA * (128.0 * 256.0) / (126.0 * 252.0)   # 3 x 3 digit products + tile padding
            peak = b.measure_i8_mma_peak()
            r.update({"peak": peak, "frac": achieved / peak if peak > 0 else None,
                      "executed": executed, "frac_executed": executed / peak if peak > 0 else None,
                      "kernel": "pair_i8_kernel<tri, CG=%d> (tcgen05.mma kind::i8, int32 accumulators in TMEM, exact int64 digit fold, "
                                "fused per-pair solve)" % b.get_i8_cta_group(),
                      "peak_source": "tcgen05 kind::i8 issue peak (int8 TOP/s) measured live by b200_measure_i8_mma_peak (operands resident "
                                     "in smem); MEASURED_PEAKS.json holds bf16 only (int8 nominal = 2x bf16).  achieved/frac are ALGORITHMIC "
                                     "(18N flop per pair); executed/frac_executed count the 81 int8 dot products per pair incl. tile padding, "
                                     "i.e. tensor-pipe utilisation (the figure SURVEY.md 8(d) ties north_star's >= 50 % target to)",
                      "frac_executed_of_2x_bf16_measured": executed / (2 * bf16) if bf16 else None,
                      # the denominators the task prescribes (MEASURED_PEAKS.json: cuBLAS bf16, burst for a kernel timed alone,
                      # sustained inside a long step), doubled because kind::i8 issues at twice the bf16 rate
                      "measured_peaks": {"unit": "TOP/s int8-equivalent = 2 x cuBLAS bf16 TFLOP/s of MEASURED_PEAKS.json",
                                         "burst": 2 * bf16_burst if bf16_burst else None, "sustained": 2 * bf16 if bf16 else None,
                                         "frac_executed_of_burst": executed / (2 * bf16_burst) if bf16_burst else None,
                                         "frac_executed_of_sustained": executed / (2 * bf16) if bf16 else None},
                      "power_note": "profiles/r2c_pair_i8_power_probe.txt: the issue probe behind `peak` (constant operands resident in "
                                    "shared memory, no operand stream) holds 1965 MHz at ~640 W; any data-fed tensor kernel on this board -- "
                                    "this one, its MMA + operand-stream skeleton, cuBLAS bf16 -- reaches the 1000 W cap and settles at "
                                    "1450-1530 MHz within a second (sw_power_cap), where removing the whole FP64 window and per-pair solve "
                                    "(-20 % cycles) buys 2 % of time: the attainable ceiling is `measured_peaks`, not the issue probe"})
        else:
            peak = max(b.measure_fp64_mma_peak(0), b.measure_fp64_mma_peak(3))
            r.update({"peak": peak, "frac": achieved / peak if peak > 0 else None,
                      "kernel": "pair_kernel<fit,tri> (FP64 DMMA.8x8x4, 32x32 frame tiles)",
                      "peak_source": "FP64 tensor (DMMA) issue peak measured live by b200_measure_fp64_mma_peak; "
                                     "MEASURED_PEAKS.json holds no FP64 figure",
                      "frac_of_bf16_measured": achieved / bf16 if bf16 else None})
        return r

    # ---- one fixed-point grid for all shards: every rank's own choice, min over ranks, pinned
    b.set_fixed_point_bits(0)
    dev_step()
    torch.cuda.synchronize()
    engine, qbits = b.last_pair_engine()
    grid = {"engine": engine, "own_bits": qbits}
    if world > 1:
        all_i8 = ctx.reduce(1.0 if engine == 2 else 0.0, "min") > 0.5
        if all_i8:
            qmin = int(ctx.reduce(float(qbits), "min"))
            b.set_fixed_point_bits(qmin)
            grid["pinned_bits"] = qmin
        else:
            b.set_pair_engine("fp64")     # (one rank ineligible: everybody takes the FP64 engine)
    dev_ms, own_ms, st, clocks = timed_device_run(steps, warmup)
    engine, qbits = b.last_pair_engine()
    value = total_pairs * steps / (dev_ms * 1e-3)
    chk = float(d_out[: min(nelt, 1 << 20)].double().sum().item()) if nelt else 0.0   # the timed work is the real work
    roofline = roofline_of(st, own_ms, engine)
    roofline["engine"] = {1: "fp64-dmma", 2: "tcgen05-int8"}.get(engine, "?")
    if engine == 2:
        roofline["fixed_point_fraction_bits"] = qbits
    if engine == 2 and not light:
        # the always-available FP64 engine on the same inputs, for context (and as a parity cross-check)
        b.set_pair_engine("fp64")
        k2 = max(1, min(steps, 2))
        ms2, own2, st2_, _ = timed_device_run(k2, 1)
        chk_fp64 = float(d_out[: min(nelt, 1 << 20)].double().sum().item()) if nelt else 0.0
        r2 = roofline_of(st2_, own2, 1)
        roofline["fp64_engine"] = {"value": total_pairs * k2 / (ms2 * 1e-3), "unit": UNIT, "achieved": r2["achieved"],
                                   "peak": r2["peak"], "frac": r2["frac"], "checksum": chk_fp64}
        b.set_pair_engine("auto")
        if abs(chk - chk_fp64) > 1e-3 * max(1.0, abs(chk)):
            raise RuntimeError("tcgen05 and FP64 engines disagree: %r vs %r" % (chk, chk_fp64))
    gpu_launches = int(st["kernel_launches"])
    del d_out, d_crd
    torch.cuda.empty_cache()

    # =================== e2e: host buffers through the C ABI ===================
    h_out = torch.empty(max(nelt, 1), dtype=torch.float32, pin_memory=True)
    h_out_base = h_out.data_ptr() - 4 * first
    L = b.lib()
    p_crd = C.c_void_p(h_crd.data_ptr())
    p_sel = sel.ctypes.data_as(C.c_void_p)
    p_mass = mass.ctypes.data_as(C.c_void_p) if mass is not None else None
    fe, ne = C.c_size_t(0), C.c_size_t(0)

    def e2e_step():
        rc = L.b200_rms2d_tri_shard(p_crd, stride, nF, None, nF, p_sel, nA, p_mass, 1, rank, world,
                                    C.c_void_p(h_out_base), C.byref(fe), C.byref(ne))
        if rc:
            raise RuntimeError(L.b200_last_error().decode())

    for _ in range(max(1, min(warmup, 3))):
        e2e_step()
    b.reset_stats()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    torch.cuda.synchronize()
    own_e2e = time.perf_counter() - t0
    e2e_s = ctx.reduce(own_e2e)
    st2 = b.get_stats()
    e2e_value = total_pairs * steps / e2e_s
    h2d = ctx.reduce(st2["h2d_bytes"] / steps)
    d2h = ctx.reduce(st2["d2h_bytes"] / steps)
    d2h_total = ctx.reduce(st2["d2h_bytes"] / steps, "sum")
    chk2 = float(h_out[: min(nelt, 1 << 20)].double().sum().item()) if nelt else 0.0
    if abs(chk - chk2) > 1e-3 * max(1.0, abs(chk)):
        raise RuntimeError("device-resident and host-path results disagree: %r vs %r" % (chk, chk2))
    ceil = host_ceiling(ctx, nelt * 4, st2["h2d_bytes"] / steps, st2["d2h_bytes"] / steps)
    d2h_rate = d2h_total / (e2e_s / steps) / 1e9
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "ms_per_step": 1e3 * e2e_s / steps,
           "api": "b200_rms2d_tri_shard, pinned host COORDS in, pinned host triangle out; bytes are per rank (max over ranks)",
           "host_ceiling": {"d2h_gbs": ceil["d2h"], "h2d_gbs": ceil["h2d"],
                            "how": "all %d rank(s) copy device <-> pinned host at the same time, best of 3" % world},
           "d2h_gbs_all_ranks": d2h_rate, "frac_of_d2h_ceiling": d2h_rate / ceil["d2h"] if ceil["d2h"] > 0 else None}
    if ceil.get("duplex_ms"):
        e2e["host_ceiling"]["duplex_ms"] = ceil["duplex_ms"]
        e2e["host_ceiling"]["duplex_how"] = ("every rank copies one step's own H2D and D2H bytes at the same time, on two streams, nothing else "
                                             "running: the floor of a step's end-to-end time on this box")
        e2e["copy_floor_frac"] = ceil["duplex_ms"] / (1e3 * e2e_s / steps)

    # =================== parity: every rank, rows of its own band, against the reference's own arithmetic ===================
    parity = None
    try:
        impl, kind, _ = cpu_impl(threads=max(1, host_cores() // world))
        rows = np.unique(np.linspace(r0, max(r0, r1 - 1), 24).astype(np.int64)) if r1 > r0 else np.zeros(0, np.int64)
        worst, npairs = 0.0, 0
        hn, cn = h_out.numpy(), h_crd.numpy()
        for i in rows:
            i = int(i)
            if i >= nF - 1:
                continue
            cols = np.arange(i + 1, min(nF, i + 1 + 2048))
            want = impl.rms2d_full(cn[i:i + 1], sel, cn[cols[0]:cols[-1] + 1], sel, massTgt=mass, massRef=mass)[0]
            got = hn[nF * i - (i + 1) * i // 2 + cols - i - 1 - first]
            worst = max(worst, float(np.abs(got.astype(np.float64) - want).max()))
            npairs += len(cols)
        parity = {"max_abs_diff_A": ctx.reduce(worst), "pairs_checked": int(ctx.reduce(float(npairs), "sum")),
                  "checker": kind, "how": "every rank: 24 rows of its own band x up to 2048 columns of the e2e result"}
    except Exception as e:  # a missing checker must not lose the bench line
        parity = {"max_abs_diff_A": None, "error": repr(e)}

    # =================== pageable host buffers, as cpptraj passes them (N == 1, main line only) ===================
    pageable = None
    if world == 1 and not light:
        crd_np = np.array(h_crd.numpy())           # pageable copy, touched (cpptraj filled it while reading the trajectory)
        ts = []
        for _ in range(5):
            out_np = np.empty(total_pairs, np.float32)   # fresh: never touched, like Matrix<float>'s new float[]
            t0 = time.perf_counter()
            b.rms2d_tri(crd_np, sel, mass=mass, out=out_np)
            ts.append(time.perf_counter() - t0)
        if not np.array_equal(out_np[: 1 << 20], h_out.numpy()[: 1 << 20]):
            raise RuntimeError("pageable and pinned host paths disagree")
        med = sorted(ts[1:])[len(ts[1:]) // 2]
        pageable = {"value": total_pairs / med, "unit": UNIT, "ms_per_step": 1e3 * med,
                    "api": "b200_rms2d_tri, pageable COORDS in, fresh pageable triangle out (what cpptraj's glue passes); "
                           "staged through pinned ring slots by the library's copy pools"}
        del crd_np, out_np
    e2e["pageable"] = pageable

    # =================== CPU baseline (rank 0) ===================
    cpu = None
    if rank == 0 and not ctx.args.no_cpu_baseline:
        try:
            cpu = cpu_tri_baseline(h_crd.numpy(), sel, mass, cpu_seconds)
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
    b.set_fixed_point_bits(0)
    b.set_pair_engine("auto")
    del h_out, h_crd
    res = {
        "value": value, "ms_per_step": dev_ms / steps, "steps": steps, "warmup": warmup,
        "dtype": "f64" if engine != 2 else "s8 x s8 -> s32 exact covariance (24-bit fixed point) + f64 solve",
        "config": {"workload": "%s: %s, %d pairs" % (name, cfg["what"], total_pairs),
                   "sharding": "upper-triangle row bands, %d rank(s), no collective" % world,
                   "engine": roofline["engine"], "fixed_point_grid": grid,
                   "l2": "no explicit flush: every step streams %.0f MB raw COORDS + %.0f MB packed operands + a %.0f MB "
                         "result triangle per rank through the 126 MB L2" % (
                             nF * stride * 4 / 1e6,
                             (nF * ((nA + 63) // 64 * 64) * 9 if engine == 2 else nF * ((nA + 15) // 16 * 16) * 24) / 1e6, nelt * 4 / 1e6),
                   "seed": cfg["seed"], "gen_seconds": round(t_gen, 2)},
        "e2e": e2e, "roofline": roofline, "parity": parity, "cpu_baseline": cpu, "clocks": clocks,
        "gpu_launches": gpu_launches, "checksum": chk,
    }
    if cpu and parity:
        cpu["parity_max_abs_diff_A"] = parity.get("max_abs_diff_A")
    return res


def onevn_bench(ctx, name, cfg, steps, warmup, cpu_seconds):
    """cfg3: one-vs-many rmsd, frames sharded over the ranks.  value = frames/s device-resident (HBM roofline),
    e2e = streamed from pinned host (PCIe)."""
    torch, b = ctx.torch, ctx.b
    world, rank = ctx.world, ctx.rank
    nA, blockF, nF = cfg["atoms"], cfg["block"], cfg["frames"]
    reps = max(1, nF // blockF // world)             # this rank's share: `reps` passes over the block
    myF = reps * blockF
    stride = 3 * nA
    h_crd = torch.empty((blockF, stride), dtype=torch.float32, pin_memory=True)
    gen_trajectory(cfg["seed"], blockF, nA, h_crd.numpy())
    sel = np.arange(nA, dtype=np.int32)
    ref_raw = h_crd.numpy()[7].reshape(-1, 3).astype(np.float64)
    ref = ref_raw - ref_raw.mean(0)
    torch.cuda.empty_cache()
    d_block = h_crd.cuda()
    d_sel = torch.from_numpy(sel).cuda()
    d_ref = torch.from_numpy(ref).cuda()
    # the rank's whole share resident in HBM (the block replicated on the device: 60 GB for the million frames on one
    # GPU) and ONE call per step over all of it; where that does not fit, one call per block
    try:
        d_crd = d_block.repeat(reps, 1) if reps > 1 else d_block
        callF, calls = myF, 1
    except RuntimeError:
        torch.cuda.empty_cache()
        d_crd, callF, calls = d_block, blockF, reps
    d_rms = torch.empty(callF, dtype=torch.float64, device="cuda")

    def dev_pass():
        for _ in range(calls):
            b.dev_rmsd_1vN(d_crd, stride, callF, d_sel, nA, d_ref, d_rms, stream=ctx.stream)

    for _ in range(max(1, warmup)):
        dev_pass()
    ctx.barrier()
    b.set_profiling(True)
    b.reset_stats()
    sampler = ClockSampler(ctx.local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record()
    for _ in range(steps):
        dev_pass()
    e1.record()
    ctx.barrier()
    own = e0.elapsed_time(e1)
    ms = ctx.reduce(own)
    clocks = sampler.result()
    st = b.get_stats()
    b.set_profiling(False)
    total_frames = myF * world
    value = total_frames * steps / (ms * 1e-3)
    # parity and argmin against the reference on the block (the million frames are the block repeated)
    got_all = d_rms.cpu().numpy()
    got = got_all[:blockF]
    replicas_equal = bool(np.array_equal(got_all.reshape(-1, blockF), np.broadcast_to(got, (callF // blockF, blockF))))
    impl, kind, cores = cpu_impl()
    t0 = time.perf_counter()
    want = impl.rmsd_1vN(h_crd.numpy(), sel, ref_raw)
    cpu_dt = time.perf_counter() - t0
    parity = {"max_abs_diff_A": ctx.reduce(float(np.abs(got - want).max())), "argmin_ref": int(np.argmin(want)),
              "checker": kind, "how": "all %d frames of the block against the reference's RMSD_CenteredRef; the other "
                                      "replicas of the block bit-equal to the first: %s" % (blockF, replicas_equal)}
    parity["replicas_bit_equal"] = replicas_equal
    # e2e: streamed from pinned host through the streaming handle (push / flush), argmin on the fly
    with b.Rmsd1vN(ref, sel, None, True, False) as h:
        def e2e_pass():
            for _ in range(reps):
                h.push(h_crd.numpy())
            return h.flush()
        e2e_pass()
        b.reset_stats()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(max(1, min(steps, 3))):
            r, _, _, best = e2e_pass()
        e2e_s = ctx.reduce(time.perf_counter() - t0) / max(1, min(steps, 3))
        st2 = b.get_stats()
    parity["argmin_exact"] = bool(best % blockF == parity["argmin_ref"] or abs(want[best % blockF] - want.min()) <= 1e-4)
    parity["e2e_max_abs_diff_A"] = float(np.abs(r[:blockF] - want).max())
    hbm = ctx.peaks.get("hbm_gbs")
    launches = max(1, st["onevn_launches"])
    # the dominant kernel alone (CUDA events around onevn_stream2_kernel); the whole pass (chunk table, part sums, streaming
    # kernel, per-frame finish) is reported beside it
    k_launches = max(1, st["onevn_stream_launches"])
    k_ms = st["onevn_stream_ms"] / k_launches if st["onevn_stream_ms"] > 0 else st["onevn_ms"] / launches
    gbs = 12.0 * nA * callF / (k_ms * 1e-3) / 1e9
    gbs_pass = 12.0 * nA * callF / (st["onevn_ms"] / launches * 1e-3) / 1e9
    res = {
        "metric": "one-vs-many fitted RMSD frames/sec (rmsd action)", "unit": "frames/s",
        "value": value, "ms_per_step": ms / steps, "steps": steps, "warmup": warmup, "dtype": "f64",
        "config": {"workload": "%s: %s; %d frames per rank and step, %d call(s) of %d frames resident in HBM" % (name, cfg["what"], myF, calls, callF),
                   "sharding": "frames over %d rank(s), argmin = min over ranks" % world, "seed": cfg["seed"]},
        "e2e": {"value": total_frames / e2e_s, "unit": "frames/s", "ms_per_step": 1e3 * e2e_s,
                "h2d_bytes_per_step": ctx.reduce(st2["h2d_bytes"] / max(1, min(steps, 3))),
                "d2h_bytes_per_step": ctx.reduce(st2["d2h_bytes"] / max(1, min(steps, 3))),
                "h2d_gbs_per_rank": 12.0 * nA * myF / e2e_s / 1e9,
                "api": "b200_rmsd_1vN_push_f32 from pinned host + flush (PCIe-bound)"},
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm if hbm else None,
                     "traffic": 1.214e9 * (callF / blockF) if (nA == 5000 and blockF == 20000) else None,
                     "traffic_source": "from_profile: profiles/r2_onevn_stream2_kernel_ncu_raw.csv (dram__bytes_read.sum + dram__bytes_write.sum "
                                       "of one launch over one 20,000-frame block, scaled to the frames of a launch)",
                     "frames_per_launch": callF,
                     "launches": int(st["onevn_stream_launches"]), "avg_launch_ms": k_ms,
                     "whole_pass": {"achieved": gbs_pass, "frac": gbs_pass / hbm if hbm else None, "avg_ms": st["onevn_ms"] / launches,
                                    "what": "chunk table + part sums + streaming kernel + per-frame finish kernel"},
                     "kernel": "onevn_stream2_kernel<float> (chunk-major: reference chunk resident in smem, frames through a 6-stage TMA ring, "
                               "13 FP64 sums per frame and chunk)",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy bandwidth); achieved = 12*N bytes per frame (SURVEY 8d) / "
                                    "CUDA-event time of the streaming kernel",
                     "kernel_share_of_step": st["onevn_stream_ms"] / own if own > 0 else None},
        "parity": parity,
        "cpu_baseline": {"value": blockF / cpu_dt, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": "the %d-frame block, %.2f s" % (blockF, cpu_dt)} if rank == 0 else None,
        "clocks": clocks, "gpu_launches": int(st["kernel_launches"]),
    }
    return res


def cpptraj_leg(cfg, cpu_seconds):
    """the same binpos file to cpptraj.B200 (full cfg2) and to the unmodified cpptraj.OMP (a prefix:
    the cost is linear in pairs), each timed by its own 'TIME: Analyses took' line."""
    import re
    b200_bin = os.path.join(ROOT, "oracle", "_ref", "cpptraj_b200", "cpptraj.B200")
    omp_bin = os.path.join(ROOT, "oracle", "_ref", "cpptraj_plain", "cpptraj.OMP")
    if not os.path.exists(b200_bin):
        return {"unavailable": "cpptraj.B200 not staged (tools/build_cpptraj_b200.sh --build needs the reference tree)"}
    nF, nA = cfg["frames"], cfg["atoms"]
    crd = np.empty((nF, 3 * nA), np.float32)
    gen_trajectory(cfg["seed"], nF, nA, crd)
    w = tempfile.mkdtemp(prefix="b200_cpptraj_")
    rec = np.zeros(nF, dtype=[("n", "<i4"), ("xyz", "<f4", (3 * nA,))])
    rec["n"] = nA
    rec["xyz"] = crd
    with open(os.path.join(w, "t.binpos"), "wb") as f:      # "fxyz" + per frame: int natom, 3N float (src/Traj_Binpos.cpp:58-151)
        f.write(b"fxyz")
        rec.tofile(f)
    with open(os.path.join(w, "t.pdb"), "w") as f:
        for i in range(nA):
            x, y, z = crd[0, 3 * i:3 * i + 3]
            f.write("ATOM  %5d  CA  ALA A%4d    %8.3f%8.3f%8.3f  1.00  0.00           C\n" % ((i + 1) % 100000, (i % 9999) + 1, x, y, z))
        f.write("END\n")

    def run(binary, nframes, env):
        # the usual shape of a deck: the analysis is set up before the trajectory is read (the B200 build starts its
        # device set-up then, in the background), frames go to the default COORDS set during `run`, then the analysis runs
        open(os.path.join(w, "in"), "w").write(
            "noprogress\nparm t.pdb\ntrajin t.binpos 1 %d\n2drms @CA R2D\nrun\n" % nframes)
        t0 = time.perf_counter()
        r = subprocess.run([binary, "-i", "in"], cwd=w, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
        wall = time.perf_counter() - t0
        if os.environ.get("B200_BENCH_CPPTRAJ_LOG"):
            open(os.environ["B200_BENCH_CPPTRAJ_LOG"], "a").write("==== %s %d frames\n%s\n" % (binary, nframes, r.stdout))
        m = re.search(r"TIME: Analyses took ([0-9.]+) seconds", r.stdout)
        if r.returncode != 0 or not m:
            return None, wall, r.stdout[-600:]
        extra = {}
        m2 = re.search(r"B200 RMSD path: \d+ device\(s\) \(set-up ([0-9.]+) s", r.stdout)
        m3 = re.search(r"atoms: ([0-9.]+) s in the library", r.stdout)
        m4 = re.search(r"TIME: Total execution time: ([0-9.]+) seconds", r.stdout)
        if m2: extra["device_setup_s"] = float(m2.group(1))
        if m3: extra["library_call_s"] = float(m3.group(1))
        if m4: extra["total_execution_s"] = float(m4.group(1))
        return float(m.group(1)), wall, extra

    out = {"file": "binpos, %d frames x %d atoms (%.0f MB), topology from a PDB" % (nF, nA, nF * (12 * nA + 4) / 1e6)}
    run(b200_bin, min(nF, 2000), dict(os.environ))     # warm-up: CUDA context, pinned ring slots
    t, wall, err = run(b200_bin, nF, dict(os.environ))
    pairs = nF * (nF - 1) // 2
    if t:
        out["cpptraj_b200"] = {"analyses_s": t, "wall_s": wall, "value": pairs / t, "unit": UNIT, "frames": nF,
                               "note": "value = pairs / 'TIME: Analyses took' of a fresh process: waits for the rest of the device set-up "
                                       "(CUDA context, ~1 s, started in the background when the command is parsed), then the library "
                                       "call (first call of the process: buffers and pinned ring slots are allocated in it)"}
        out["cpptraj_b200"].update(err or {})
        if out["cpptraj_b200"].get("library_call_s"):
            out["cpptraj_b200"]["value_library_call"] = pairs / out["cpptraj_b200"]["library_call_s"]
    else:
        out["cpptraj_b200"] = {"error": err}
    if os.path.exists(omp_bin):
        cores = host_cores()
        nf = int(min(nF, max(1000, (2 * 2.0e5 * cores * cpu_seconds) ** 0.5)))
        t, wall, err = run(omp_bin, nf, dict(os.environ, OMP_NUM_THREADS=str(cores)))
        out["cpptraj_omp"] = ({"analyses_s": t, "wall_s": wall, "value": nf * (nf - 1) / 2 / t, "unit": UNIT, "frames": nf, "threads": cores,
                               "note": "the UNMODIFIED reference binary on the first %d frames of the same file" % nf}
                              if t else {"error": err})
        if t and isinstance(err, dict):
            out["cpptraj_omp"].update(err)
        if out["cpptraj_b200"].get("value") and out["cpptraj_omp"].get("value"):
            out["ratio"] = out["cpptraj_b200"]["value"] / out["cpptraj_omp"]["value"]
    else:
        out["cpptraj_omp"] = {"unavailable": "oracle/_ref/cpptraj_plain/cpptraj.OMP not staged"}
    # ---- the two other cpptraj commands of the path that are dominated by device work: hierarchical clustering (cache
    #      fill + every merge on the device) and rmsavgcorr (all window sizes); the unmodified binary gets a prefix
    def run_deck(binary, text, env, pats):
        open(os.path.join(w, "in"), "w").write(text)
        t0 = time.perf_counter()
        r = subprocess.run([binary, "-i", "in"], cwd=w, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1200)
        res = {"wall_s": time.perf_counter() - t0}
        if os.environ.get("B200_BENCH_CPPTRAJ_LOG"):
            open(os.environ["B200_BENCH_CPPTRAJ_LOG"], "a").write("==== %s\n%s\n%s\n" % (binary, text, r.stdout))
        if r.returncode != 0 or "Error" in r.stdout:
            return {"error": r.stdout[-600:]}
        for k, pat in pats.items():
            m = re.search(pat, r.stdout)
            if m:
                res[k] = float(m.group(1))
        return res

    cores = host_cores()
    omp_env = dict(os.environ, OMP_NUM_THREADS=str(cores))
    cl_pats = {"pairwise_s": r"Pairwise Calc\.\s*:\s*([0-9.]+) s", "clustering_s": r"TIME:\s+Clustering\s*:\s*([0-9.]+) s",
               "analyses_s": r"TIME: Analyses took ([0-9.]+) seconds", "device_merges_s": r"initial clusters on the device in ([0-9.]+) s",
               "cluster_init_s": r"Cluster Init\.\s*:\s*([0-9.]+) s", "cluster_post_s": r"Cluster Post\.\s*:\s*([0-9.]+) s",
               "best_rep_s": r"Find best rep\.\s*:\s*([0-9.]+) s", "summary_s": r"Summary calc\s*:\s*([0-9.]+) s"}
    cl_deck = "noprogress\nparm t.pdb\ntrajin t.binpos 1 %d\ncluster C1 @CA hieragglo clusters 10 averagelinkage rms summary cl.summary.dat\nrun\n"
    nB, nO = nF, min(nF, 1500)
    out["cluster_hieragglo"] = {"deck": "cluster C1 @CA hieragglo clusters 10 averagelinkage rms summary cl.summary.dat (in-memory pairwise cache)",
                                "cpptraj_b200": dict(run_deck(b200_bin, cl_deck % nB, dict(os.environ), cl_pats), frames=nB)}
    if os.path.exists(omp_bin):
        out["cluster_hieragglo"]["cpptraj_omp"] = dict(run_deck(omp_bin, cl_deck % nO, omp_env, cl_pats), frames=nO, threads=cores,
                                                       note="the UNMODIFIED reference binary on the first %d frames; its merge loop "
                                                            "costs O(cluster size x frames) cache reads per merge" % nO)
    ra_pats = {"analyses_s": r"TIME: Analyses took ([0-9.]+) seconds", "device_s": r"frames on the device in ([0-9.]+) s"}
    ra_deck = "noprogress\nparm t.pdb\ntrajin t.binpos 1 %d\nrmsavgcorr RA @CA first\nrun\n"
    nB, nO = nF, min(nF, 2500)
    rmsds = lambda n: n * (n + 1) / 2.0 - 1.0      # window sizes 1 .. n-1: sum of (n - W + 1)
    out["rmsavgcorr"] = {"deck": "rmsavgcorr RA @CA first (every window size)", "unit": "averaged-frame RMSDs/s",
                         "cpptraj_b200": dict(run_deck(b200_bin, ra_deck % nB, dict(os.environ), ra_pats), frames=nB)}
    if out["rmsavgcorr"]["cpptraj_b200"].get("analyses_s"):
        out["rmsavgcorr"]["cpptraj_b200"]["value"] = rmsds(nB) / out["rmsavgcorr"]["cpptraj_b200"]["analyses_s"]
    if os.path.exists(omp_bin):
        o = dict(run_deck(omp_bin, ra_deck % nO, omp_env, ra_pats), frames=nO, threads=cores,
                 note="the UNMODIFIED reference binary on the first %d frames (cost ~ frames^2 x atoms)" % nO)
        if o.get("analyses_s"):
            o["value"] = rmsds(nO) / o["analyses_s"]
            if out["rmsavgcorr"]["cpptraj_b200"].get("value"):
                out["rmsavgcorr"]["ratio"] = out["rmsavgcorr"]["cpptraj_b200"]["value"] / o["value"]
        out["rmsavgcorr"]["cpptraj_omp"] = o
    # ---- BASELINE config 5 as a user runs it: 2drms of 100,000 x 1,000 (a 20 GB Matrix<float>) through the B200 binary, one
    #      process from the trajectory file to the filled matrix; the unmodified binary's rate is the prefix run above
    if not os.environ.get("B200_BENCH_NO_BIG_CPPTRAJ"):
        try:
            c5 = CONFIGS["cfg5"]
            nF5, nA5 = c5["frames"], c5["atoms"]
            del rec, crd
            big = np.zeros(nF5, dtype=[("n", "<i4"), ("xyz", "<f4", (3 * nA5,))])
            big["n"] = nA5
            gen_trajectory(c5["seed"], nF5, nA5, big["xyz"])
            with open(os.path.join(w, "t5.binpos"), "wb") as f:
                f.write(b"fxyz")
                big.tofile(f)
            del big
            pats = {"analyses_s": r"TIME: Analyses took ([0-9.]+) seconds", "library_call_s": r"atoms: ([0-9.]+) s in the library",
                    "device_setup_s": r"\(set-up ([0-9.]+) s", "trajectory_s": r"Trajectory Process : ([0-9.]+) s",
                    "total_execution_s": r"TIME: Total execution time: ([0-9.]+) seconds"}
            r5 = run_deck(b200_bin, "noprogress\nparm t.pdb\ntrajin t5.binpos\n2drms @CA R2D\nrun\n", dict(os.environ), pats)
            pairs5 = nF5 * (nF5 - 1) // 2
            leg5 = {"deck": "2drms @CA R2D on %d frames x %d atoms (binpos, %.1f GB), result: a %.0f GB Matrix<float> in pageable memory" % (
                        nF5, nA5, nF5 * (12 * nA5 + 4) / 1e9, pairs5 * 4 / 1e9), "cpptraj_b200": dict(r5, frames=nF5), "unit": UNIT}
            if r5.get("analyses_s"):
                leg5["cpptraj_b200"]["value"] = pairs5 / r5["analyses_s"]
                if r5.get("library_call_s"):
                    leg5["cpptraj_b200"]["value_library_call"] = pairs5 / r5["library_call_s"]
                if out.get("cpptraj_omp", {}).get("value"):
                    leg5["ratio"] = leg5["cpptraj_b200"]["value"] / out["cpptraj_omp"]["value"]
                    leg5["ratio_how"] = ("pairs / 'TIME: Analyses took' of cpptraj.B200 on the whole matrix over the rate of the UNMODIFIED "
                                         "cpptraj.OMP on a prefix of the cfg2 file (its cost is linear in pairs)")
            out["cfg5_size"] = leg5
        except Exception as e:       # (disk or memory of the box: the leg is optional)
            out["cfg5_size"] = {"error": repr(e)}
    for f in os.listdir(w):
        os.remove(os.path.join(w, f))
    os.rmdir(w)
    return out


def summary(res):
    """A leg of the `configs` object: the headline numbers of a full result."""
    keep = {k: res.get(k) for k in ("metric", "unit", "value", "ms_per_step", "steps", "dtype", "parity", "cpu_baseline", "clocks")}
    keep["metric"], keep["unit"] = keep["metric"] or METRIC, keep["unit"] or UNIT
    keep["workload"] = res["config"]["workload"]
    e = res["e2e"]
    keep["e2e"] = {k: e.get(k) for k in ("value", "ms_per_step", "h2d_bytes_per_step", "d2h_bytes_per_step", "frac_of_d2h_ceiling", "copy_floor_frac",
                                         "host_ceiling", "h2d_gbs_per_rank") if k in e}
    r = res["roofline"]
    keep["roofline"] = {k: r.get(k) for k in ("bound", "achieved", "peak", "unit", "frac", "executed", "frac_executed", "engine") if k in r}
    return keep


def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)
    ctx = Ctx(args)
    name, cfg = main_config(args)
    if cfg["kind"] == "1vn":
        res = onevn_bench(ctx, name, cfg, args.steps, args.warmup, args.cpu_seconds)
    else:
        res = tri_bench(ctx, name, cfg, args.steps, args.warmup, args.cpu_seconds)
    extras = {}
    if args.config == "auto" and not args.frames and not args.no_extras:
        legs = ["cfg5", "cfg4", "cfg3"] if name == "cfg2" else ["cfg4", "cfg3"]
        for leg in legs:
            try:
                c = CONFIGS[leg]
                if c["kind"] == "1vn":
                    extras[leg] = summary(onevn_bench(ctx, leg, c, 3, 1, 3.0))
                else:
                    extras[leg] = summary(tri_bench(ctx, leg, c, 2, 1, 4.0, light=True))
            except Exception as e:       # a leg must not lose the main line
                extras[leg] = {"error": repr(e)}
        if ctx.world == 1 and ctx.rank == 0:
            try:
                extras["cpptraj"] = cpptraj_leg(CONFIGS["cfg2"], 3.0)
            except Exception as e:
                extras["cpptraj"] = {"error": repr(e)}
    if ctx.rank == 0:
        line = {
            "metric": res.get("metric", METRIC), "value": res["value"], "unit": res.get("unit", UNIT), "n_gpus": ctx.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": res["dtype"], "data": "synthetic", "config": res["config"],
            "e2e": res["e2e"], "roofline": res["roofline"], "parity": res["parity"], "cpu_baseline": res["cpu_baseline"],
            "clocks": res["clocks"], "gpu_launches": res["gpu_launches"],
        }
        if "checksum" in res:
            line["checksum"] = res["checksum"]
        if extras:
            line["configs"] = extras
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()
    ctx.b.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
