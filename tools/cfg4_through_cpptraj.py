"""BASELINE config 4 as a user would run it: `cluster hieragglo ... rms mass` (in-memory pairwise cache) through the cpptraj.B200
binary on 50,000 frames x 2,000 atoms of the synthetic trajectory (binpos + a PDB topology with C/N/O/S/H masses);
CPPTRAJ_B200_NGPU selects the devices of the cache fill.  Prints cpptraj's own TIME lines.
usage: python tools/cfg4_through_cpptraj.py [frames] [atoms] [clusters]     (clusters = 0: `2drms * R2D` instead, BASELINE config 5
as a user would run it with 100000 1000 0; seed of that config)"""
import os, re, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
nF = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
nA = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
nC = int(sys.argv[3]) if len(sys.argv) > 3 else 10
binary = os.path.join(ROOT, "oracle", "_ref", "cpptraj_b200", "cpptraj.B200")
w = tempfile.mkdtemp(prefix="b200_cfg4_")
t0 = time.perf_counter()
crd = np.empty((nF, 3 * nA), np.float32)
bench.gen_trajectory(20261019 if nC > 0 else 20261020, nF, nA, crd)
rec = np.zeros(nF, dtype=[("n", "<i4"), ("xyz", "<f4", (3 * nA,))])
rec["n"] = nA; rec["xyz"] = crd
with open(os.path.join(w, "t.binpos"), "wb") as f:
    f.write(b"fxyz"); rec.tofile(f)
el = ["C", "N", "O", "S", "H"]
with open(os.path.join(w, "t.pdb"), "w") as f:
    for i in range(nA):
        x, y, z = crd[0, 3 * i:3 * i + 3]
        e = el[i % 5]
        f.write("ATOM  %5d  %-3s ALA A%4d    %8.3f%8.3f%8.3f  1.00  0.00          %2s\n" % ((i + 1) % 100000, e, (i % 9999) + 1, x, y, z, e))
    f.write("END\n")
del rec, crd
print("trajectory written in %.1f s (%.2f GB)" % (time.perf_counter() - t0, nF * (12 * nA + 4) / 1e9), flush=True)
if nC > 0:
    open(os.path.join(w, "in"), "w").write(
        "noprogress\nparm t.pdb\ntrajin t.binpos\ncluster C1 * hieragglo clusters %d averagelinkage rms mass summary s.dat out cn.dat\nrun\n" % nC)
else:
    open(os.path.join(w, "in"), "w").write("noprogress\nparm t.pdb\ntrajin t.binpos\n2drms * R2D\nrun\n")
t0 = time.perf_counter()
r = subprocess.run([binary, "-i", "in"], cwd=w, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=3000)
print("cpptraj.B200 exit %d, wall %.1f s" % (r.returncode, time.perf_counter() - t0))
for line in r.stdout.splitlines():
    if re.search(r"TIME|B200|Error|Warning|clusters|Caching|Estimated|hieragglo", line): print(line)
if os.path.exists(os.path.join(w, "s.dat")): print(open(os.path.join(w, "s.dat")).read())
