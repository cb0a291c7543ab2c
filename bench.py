#!/usr/bin/env python
"""bench.py -- fitted pair-RMSDs/s of the rms2d hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nproc-per-node N bench.py --gpus N ...           (one rank per GPU)

Workload (synthetic, cpptraj_b200.synth, seeds from SURVEY.md 8d):
    N == 1 : BASELINE configs[1]  rms2d fit, 10,000 frames x 1,000 atoms (49,995,000 pairs)
    N  > 1 : BASELINE configs[4]  rms2d fit, 100,000 frames x 1,000 atoms, the upper triangle
             sharded into contiguous, pair-balanced row bands over the ranks (strong scaling,
             no data-path collective: every rank packs the frames it needs itself).
A "step" is one pass over all pairs of the rank's shard.

One JSON line is printed by rank 0:
    value      pairs/s, whole job, inputs (raw float32 COORDS) resident in HBM, timed with CUDA
               events on the stream the kernels are launched on: pack + pair-tile kernels,
               result triangle left in HBM.
    e2e        same metric through the C ABI (b200_rms2d_tri_shard) with pinned HOST buffers:
               H2D of the COORDS, kernels, D2H of the float triangle all inside the timed region.
    roofline   pair-tile kernel of the engine the library chose (tcgen05 int8 when eligible, else FP64 DMMA):
               algorithmic flop (18*N per fitted pair) / CUDA-event kernel time; `executed` counts the 81 int8
               digit products per pair the tensor pipe actually runs; peak = this device's tcgen05 kind::i8
               (or FP64 DMMA) issue peak measured live by a probe kernel (MEASURED_PEAKS.json holds bf16 only;
               the fraction of 2x the measured bf16 figure is reported beside it).
    cpu_baseline  the reference's own Frame::RMSD_CenteredRef loop (oracle/_ref, OpenMP, all host
               threads) on a bounded prefix of the same trajectory (N == 1, rank 0 only).

--impl reference times that CPU implementation alone (rank 0 only under torchrun).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fitted pair-RMSDs/sec (rms2d FxN)"
UNIT = "pair-RMSDs/s"
SEEDS = {"cfg2": 20261017, "cfg5": 20261020}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=0, help="override frame count")
    ap.add_argument("--atoms", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU seconds for the bounded baseline sample")
    return ap.parse_args()


def workload(args):
    if args.frames:
        return "custom", args.frames, args.atoms, 20261017
    if args.gpus > 1:
        return "cfg5", 100000, args.atoms, SEEDS["cfg5"]
    return "cfg2", 10000, args.atoms, SEEDS["cfg2"]


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


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_pairs_rate(crd, sel, target_seconds, threads=None):
    """Time the reference's own pair loop (oracle/_ref) on a bounded prefix of crd.
    Returns dict(value, cores, kind, sample, seconds)."""
    from oracle.pyoracle import Oracle, Reference, have_reference
    if have_reference():
        impl, kind = Reference(), "reference"
    else:
        impl, kind = Oracle(), "port"
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    if threads:
        cores = threads
    impl.set_threads(cores)

    def run(nf):
        t0 = time.perf_counter()
        impl.rms2d_tri(crd[:nf], sel)
        dt = time.perf_counter() - t0
        if kind == "reference":
            dt = impl.last_loop_seconds()   # pair loop only, what cpptraj's TIME: line covers
        return dt

    nf = min(crd.shape[0], 400)
    dt = max(run(nf), 1e-4)
    rate = nf * (nf - 1) / 2 / dt
    nf2 = int(min(crd.shape[0], max(nf, (2 * rate * target_seconds) ** 0.5)))
    return impl, kind, cores, nf2, rate


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from cpptraj_b200.synth import make_trajectory
    name, nF, nA, seed = workload(args)
    # a bounded prefix is all the CPU can do in minutes; cost is exactly linear in pairs
    nGen = min(nF, 12000)
    crd, _ = make_trajectory(seed, nGen, nA)
    sel = np.arange(nA, dtype=np.int32)
    steps_total = max(1, args.steps + args.warmup)
    per_step = max(2.0, min(20.0, 150.0 / steps_total))
    impl, kind, cores, nf, _ = cpu_pairs_rate(crd, sel, per_step)
    pairs = nf * (nf - 1) / 2

    def step():
        t0 = time.perf_counter()
        impl.rms2d_tri(crd[:nf], sel)
        dt = time.perf_counter() - t0
        return impl.last_loop_seconds() if kind == "reference" else dt

    for _ in range(args.warmup):
        step()
    tot = sum(step() for _ in range(args.steps))
    value = pairs * args.steps / tot
    sample = "rms2d fit on the first %d of %d frames x %d atoms (%d pairs/step), %d OpenMP threads" % (nf, nF, nA, pairs, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: rms2d fit, %d frames x %d atoms (CPU arm times a %d-frame prefix per step)" % (name, nF, nA, nf)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------- B200 arm
def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)

    import ctypes as C
    import torch
    import torch.distributed as dist
    import cpptraj_b200 as b
    from cpptraj_b200.synth import make_trajectory

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE %d != --gpus %d" % (world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    b.init(devices=[local])

    name, nF, nA, seed = workload(args)
    stride = 3 * nA
    # ---- synthetic COORDS in pinned host memory (every rank generates the same trajectory)
    t_gen = time.perf_counter()
    h_crd = torch.empty((nF, stride), dtype=torch.float32, pin_memory=True)
    make_trajectory(seed, nF, nA, out=h_crd.numpy())
    t_gen = time.perf_counter() - t_gen
    sel = np.arange(nA, dtype=np.int32)
    r0, r1 = b.shard_rows(nF, rank, world)
    first = nF * r0 - r0 * (r0 + 1) // 2
    nelt = (nF * r1 - r1 * (r1 + 1) // 2) - first
    total_pairs = nF * (nF - 1) // 2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # =================== value: device-resident ===================
    d_crd = h_crd.cuda(non_blocking=False)
    d_sel = torch.from_numpy(sel).cuda()
    d_out = torch.empty(max(nelt, 1), dtype=torch.float32, device="cuda")
    # the ABI indexes the whole triangle; hand it a base pointer such that base[first] is d_out[0]
    d_out_base = d_out.data_ptr() - 4 * first
    stream = torch.cuda.current_stream().cuda_stream

    def dev_step():
        b.dev_rms2d_tri(d_crd, stride, nF, d_sel, nA, d_out_base, fit=True, rank=rank, count=world, stream=stream)

    peaks_file = {}
    try:
        peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    def timed_device_run(steps, warmup):
        """warmup + `steps` timed passes; returns (ms max over ranks, this rank's ms, stats, clocks)."""
        b.set_profiling(False)
        for _ in range(warmup):
            dev_step()
        barrier()
        b.set_profiling(True)
        b.reset_stats()
        sampler = ClockSampler(local)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            dev_step()
        e1.record()
        barrier()
        own = e0.elapsed_time(e1)
        ms = max_over_ranks(own)
        clk = sampler.result()
        stt = b.get_stats()
        b.set_profiling(False)
        return ms, own, stt, clk

    def roofline_of(stt, own_ms, engine):
        """Roofline of the dominant (pair-tile) kernel on this rank, from CUDA events on its stream."""
        launches = max(1, stt["pair_launches"])
        ms_per_launch = stt["pair_ms"] / launches
        flop_per_launch = 18.0 * nA * (stt["pairs"] / launches)       # algorithmic: SURVEY 8(d), 18*N per fitted pair
        achieved = flop_per_launch / (ms_per_launch * 1e-3) / 1e12 if ms_per_launch > 0 else 0.0
        # dram__bytes_read + dram__bytes_write of one launch of the dominant kernel, from the committed ncu --set full
        # capture of this exact workload (profiles/r1d_pair_i8_kernel.md: 56.0 MB + 45.7 MB); null for other shapes
        traffic = 101.7e6 if (engine == 2 and name == "cfg2" and world == 1) else None
        r = {"bound": "tensor", "achieved": achieved, "unit": "TFLOP/s", "traffic": traffic,
             "launches": int(stt["pair_launches"]), "avg_launch_ms": ms_per_launch,
             "kernel_share_of_step": stt["pair_ms"] / own_ms if own_ms > 0 else None}
        bf16 = peaks_file.get("bf16_tflops_sustained")
        if engine == 2:
            # 3 x 3 int8 digit products per covariance entry: the tensor pipe executes 9x the algorithmic flop
            kpad = (nA + 63) // 64 * 64
            executed = achieved * 9.0 * kpad / nA * (128.0 * 256.0) / (126.0 * 252.0)
            peak = b.measure_i8_mma_peak()
            r.update({"peak": peak, "frac": achieved / peak if peak > 0 else None,
                      "executed": executed, "frac_executed": executed / peak if peak > 0 else None,
                      "kernel": ("pair_i8_kernel<tri, CG=2> (tcgen05.mma.cta_group::2 kind::i8 M256 N256 K32 on CTA pairs, int32 "
                                 "accumulators in TMEM, 28x28 frame-pair tiles, int64 digit fold + fused per-pair solve)" if b.get_i8_cta_group() == 2 else
                                 "pair_i8_kernel<tri, CG=1> (tcgen05.mma kind::i8 M128 N256 K32, int32 accumulators in TMEM, "
                                 "14x28 frame-pair tiles, fused FP64 solve)"),
                      "cta_group": b.get_i8_cta_group(),
                      "peak_source": "tcgen05 kind::i8 issue peak (int8 TOP/s) measured live by b200_measure_i8_mma_peak "
                                     "(operands resident in smem); MEASURED_PEAKS.json holds bf16 only (int8 nominal = 2x bf16). "
                                     "achieved/frac are ALGORITHMIC (18N flop per pair); executed/frac_executed count the 81 "
                                     "int8 dot products per pair incl. tile padding, i.e. tensor-pipe utilisation",
                      "frac_executed_of_2x_bf16_measured": executed / (2 * bf16) if bf16 else None,
                      "note": "frac = algorithmic / peak as the bench contract defines it; the exact-integer method executes 9x the "
                              "algorithmic flop by design (3 x 3 digits), so tensor-pipe utilisation -- the figure SURVEY.md 8(d) ties "
                              "north_star's >= 50 % target to -- is frac_executed"})
        else:
            peak = max(b.measure_fp64_mma_peak(0), b.measure_fp64_mma_peak(3))
            r.update({"peak": peak, "frac": achieved / peak if peak > 0 else None,
                      "kernel": "pair_kernel<fit,tri> (FP64 DMMA.8x8x4, 32x32 frame tiles)",
                      "peak_source": "FP64 tensor (DMMA) issue peak measured live by b200_measure_fp64_mma_peak; "
                                     "MEASURED_PEAKS.json holds no FP64 figure",
                      "frac_of_bf16_measured": achieved / bf16 if bf16 else None})
        return r

    dev_ms, own_ms, st, clocks = timed_device_run(args.steps, args.warmup)
    engine, qbits = b.last_pair_engine()
    value = total_pairs * args.steps / (dev_ms * 1e-3)
    # checksum so that the timed work is demonstrably the real work
    chk = float(d_out[: min(nelt, 1 << 20)].double().sum().item()) if nelt else 0.0
    roofline = roofline_of(st, own_ms, engine)
    roofline["engine"] = {1: "fp64-dmma", 2: "tcgen05-int8"}.get(engine, "?")
    if engine == 2:
        roofline["fixed_point_fraction_bits"] = qbits
        # the always-available FP64 engine on the same inputs, for context (and as a parity cross-check)
        b.set_pair_engine("fp64")
        k2 = max(1, min(args.steps, 2))
        ms2, own2, st2_, _ = timed_device_run(k2, 1)
        chk_fp64 = float(d_out[: min(nelt, 1 << 20)].double().sum().item()) if nelt else 0.0
        r2 = roofline_of(st2_, own2, 1)
        roofline["fp64_engine"] = {"value": total_pairs * k2 / (ms2 * 1e-3), "unit": UNIT, "achieved": r2["achieved"],
                                   "peak": r2["peak"], "frac": r2["frac"], "checksum": chk_fp64}
        b.set_pair_engine("auto")
        if abs(chk - chk_fp64) > 1e-3 * max(1.0, abs(chk)):
            raise RuntimeError("tcgen05 and FP64 engines disagree: %r vs %r" % (chk, chk_fp64))
        dev_step()   # leave d_out as the primary engine wrote it
        torch.cuda.synchronize()
    gpu_launches = int(st["kernel_launches"])
    del d_out
    torch.cuda.empty_cache()

    # =================== e2e: host buffers through the C ABI ===================
    h_out = torch.empty(max(nelt, 1), dtype=torch.float32, pin_memory=True)
    h_out_base = h_out.data_ptr() - 4 * first
    L = b.lib()
    p_crd = C.c_void_p(h_crd.data_ptr())
    p_sel = sel.ctypes.data_as(C.c_void_p)
    fe, ne = C.c_size_t(0), C.c_size_t(0)

    def e2e_step():
        rc = L.b200_rms2d_tri_shard(p_crd, stride, nF, None, nF, p_sel, nA, None, 1, rank, world,
                                    C.c_void_p(h_out_base), C.byref(fe), C.byref(ne))
        if rc:
            raise RuntimeError(L.b200_last_error().decode())

    for _ in range(max(1, min(args.warmup, 3))):
        e2e_step()
    b.reset_stats()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    st2 = b.get_stats()
    e2e_value = total_pairs * args.steps / e2e_s
    h2d = max_over_ranks(st2["h2d_bytes"] / args.steps)
    d2h = max_over_ranks(st2["d2h_bytes"] / args.steps)
    chk2 = float(h_out[: min(nelt, 1 << 20)].double().sum().item()) if nelt else 0.0
    if abs(chk - chk2) > 1e-3 * max(1.0, abs(chk)):
        raise RuntimeError("device-resident and host-path results disagree: %r vs %r" % (chk, chk2))

    # =================== CPU baseline (rank 0, N == 1 only) ===================
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            impl, kind, cores, nf, _ = cpu_pairs_rate(h_crd.numpy(), sel, args.cpu_seconds)
            t0 = time.perf_counter()
            ref_out = impl.rms2d_tri(h_crd.numpy()[:nf], sel)
            dt = time.perf_counter() - t0
            if kind == "reference":
                dt = impl.last_loop_seconds()
            cpu = {"value": nf * (nf - 1) / 2 / dt, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": "rms2d fit on the first %d of %d frames x %d atoms (%d pairs), %.1f s" % (nf, nF, nA, nf * (nf - 1) // 2, dt)}
            # the bounded sample doubles as a parity spot check (rows < nf of shard 0)
            n_chk = min(nf, r1)
            if n_chk > 1:
                idx = [nF * i - (i + 1) * i // 2 + np.arange(i + 1, nf) - i - 1 for i in range(min(n_chk, 64))]
                got = np.concatenate([h_out.numpy()[ix - first] for ix in idx])
                want = np.concatenate([ref_out[nf * i - (i + 1) * i // 2 + np.arange(i + 1, nf) - i - 1] for i in range(min(n_chk, 64))])
                cpu["parity_max_abs_diff_A"] = float(np.abs(got.astype(np.float64) - want).max())
        except Exception as e:  # the baseline is a reported number, never a reason to lose the bench line
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64" if engine != 2 else "s8 x s8 -> s32 exact covariance (24-bit fixed point) + f64 solve",
            "data": "synthetic",
            "config": {"workload": "%s: rms2d fit, %d frames x %d atoms, %d pairs" % (name, nF, nA, total_pairs),
                       "sharding": "upper-triangle row bands, %d rank(s), no collective" % world,
                       "engine": roofline["engine"],
                       "l2": "no explicit flush: every step streams %.0f MB raw COORDS + %.0f MB packed operands + a %.0f MB "
                             "result triangle per rank through the 126 MB L2" % (
                           nF * stride * 4 / 1e6,
                           (nF * ((nA + 63) // 64 * 64) * 9 if engine == 2 else nF * ((nA + 15) // 16 * 16) * 24) / 1e6,
                           nelt * 4 / 1e6),
                       "seed": seed, "gen_seconds": round(t_gen, 2)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "api": "b200_rms2d_tri_shard, pinned host COORDS in, pinned host triangle out"},
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "gpu_launches": gpu_launches,
            "checksum": chk,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    b.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
