"""Times b200_hieragglo on synthetic caches (random points in 8-D: clustered structure, no exact ties) and checks
the first merges against the CPU restatement on a prefix-sized problem.  usage: python tools/hieragglo_time.py [N ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cpptraj_b200 as b

def cache(rng, n, dim=8):
    c = rng.standard_normal((max(4, n // 200), dim)) * 6.0
    x = c[rng.integers(len(c), size=n)] + rng.standard_normal((n, dim))
    x = x.astype(np.float32)
    out = np.empty(n * (n - 1) // 2, np.float32)
    g = (x * x).sum(1)
    pos = 0
    for i in range(n - 1):
        d2 = g[i] + g[i + 1:] - 2.0 * (x[i + 1:] @ x[i])
        out[pos:pos + n - 1 - i] = np.sqrt(np.maximum(d2, 0))
        pos += n - 1 - i
    return out

b.init(1)
rng = np.random.default_rng(5)
sizes = [int(a) for a in sys.argv[1:]] or [2000, 10000, 30000]
for n in sizes:
    tri = cache(rng, n)
    for linkage in (0, 1, 2):
        for team in ([None] if n < 8000 else [None, 1, 4, 16]):
            if team is None:
                os.environ.pop("B200_HA_TEAM", None)
            else:
                os.environ["B200_HA_TEAM"] = str(team)
            b.hieragglo(tri[:3], 3, linkage)
            t0 = time.time()
            into, frm, fmin = b.hieragglo(tri, n, linkage, 10, None)
            dt = time.time() - t0
            print("N=%d linkage=%d team=%s: %d merges in %.3f s (%.2f us/merge incl. %.2f GB upload), last min %.4f"
                  % (n, linkage, team, len(into), dt, 1e6 * dt / max(1, len(into)), tri.nbytes / 1e9, fmin[-1]), flush=True)
    if n <= 3000:
        from oracle.pyoracle import Oracle
        t0 = time.time()
        want = Oracle().hieragglo(tri, n, 1, 10, None)
        print("  CPU restatement (frame-pair loops, 1 thread): %.2f s; merges equal: %s"
              % (time.time() - t0, np.array_equal(want[0], b.hieragglo(tri, n, 1, 10, None)[0])), flush=True)
