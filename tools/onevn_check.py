"""One-vs-many kernels on cfg3's block (device-resident): parity against the oracle and rate; B200_1VN_V2=0/1 selects
the streaming variant.  usage: python tools/onevn_check.py [frames] [atoms]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cpptraj_b200 as b
from cpptraj_b200.synth import make_trajectory
from oracle.pyoracle import Oracle
nF = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
nA = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
b.init(1)
crd, _ = make_trajectory(20261018, nF, nA)
sel = np.arange(nA, dtype=np.int32)
ref_raw = crd[7].reshape(-1, 3).astype(np.float64); ref = ref_raw - ref_raw.mean(0)
d_crd = torch.from_numpy(crd).cuda(); d_sel = torch.from_numpy(sel).cuda(); d_ref = torch.from_numpy(ref).cuda()
d_rms = torch.empty(nF, dtype=torch.float64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def step(): b.dev_rmsd_1vN(d_crd, 3 * nA, nF, d_sel, nA, d_ref, d_rms, stream=st)
step(); torch.cuda.synchronize()
o = Oracle(); o.set_threads(16)
want = o.rmsd_1vN(crd[:2000], sel, ref_raw)
print("v2=%s parity on 2000 frames: %.3e" % (os.environ.get("B200_1VN_V2", "auto"), np.abs(d_rms.cpu().numpy()[:2000] - want).max()))
for _ in range(3): step()
torch.cuda.synchronize()
b.set_profiling(True); b.reset_stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
s = b.get_stats(); ms = e0.elapsed_time(e1) / 10
print("  %.3f ms/pass  %.2f M frames/s  %.0f GB/s  (kernels %.3f ms, streaming kernel alone %.3f ms = %.0f GB/s)"
      % (ms, nF / ms / 1e3, 12.0 * nA * nF / ms / 1e6, s["onevn_ms"] / 10, s["onevn_stream_ms"] / max(1, s["onevn_stream_launches"]),
         12.0 * nA * nF / (s["onevn_stream_ms"] / max(1, s["onevn_stream_launches"])) / 1e6))
