"""Device time of every shard of an N-way split of the rms2d triangle, measured one after the other on ONE GPU
(load balance of b200_shard_rows).  usage: python tools/shard_times.py [frames] [atoms] [shards]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpptraj_b200 as b
from cpptraj_b200.synth import make_trajectory
nF = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
nA = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
nS = int(sys.argv[3]) if len(sys.argv) > 3 else 8
only = [int(x) for x in sys.argv[4].split(',')] if len(sys.argv) > 4 else list(range(nS))
b.init(1)
h = torch.empty((nF, 3 * nA), dtype=torch.float32, pin_memory=True)
make_trajectory(20261020, nF, nA, out=h.numpy())
d_crd = h.cuda(); d_sel = torch.arange(nA, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for r in only:
    r0, r1 = b.shard_rows(nF, r, nS)
    first = nF * r0 - r0 * (r0 + 1) // 2
    nelt = (nF * r1 - r1 * (r1 + 1) // 2) - first
    d_out = torch.empty(nelt, dtype=torch.float32, device="cuda")
    base = d_out.data_ptr() - 4 * first
    for _ in range(2):
        b.dev_rms2d_tri(d_crd, 3 * nA, nF, d_sel, nA, base, fit=True, rank=r, count=nS, stream=st)
    torch.cuda.synchronize()
    b.set_profiling(True); b.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        b.dev_rms2d_tri(d_crd, 3 * nA, nF, d_sel, nA, base, fit=True, rank=r, count=nS, stream=st)
    e1.record(); torch.cuda.synchronize()
    s = b.get_stats(); b.set_profiling(False)
    print("shard %d/%d rows [%6d,%6d) pairs %.3e  %.2f ms/step  pair kernels %.2f ms (%d launches)  pack %.2f ms  -> %.3e pairs/s" % (
        r, nS, r0, r1, nelt, e0.elapsed_time(e1) / 3, s["pair_ms"] / 3, s["pair_launches"] // 3, s["pack_ms"] / 3, nelt / (e0.elapsed_time(e1) / 3) * 1e3))
    del d_out
