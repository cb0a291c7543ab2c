"""Deterministic synthetic coordinate trajectories.

Base conformation: 3-D random walk with 3.8 A steps (CA-like chain).  Frame f is
a random rigid rotation + translation (+-20 A) of the base plus N(0, sigma_f^2)
noise with sigma_f cycling over {0, 0.05, 0.5, 2} A, so pair RMSDs span 0..3 A;
every 64th frame is a bit-exact duplicate of its predecessor.  Stored float32
in cpptraj's COORDS layout: one row per frame, xyz interleaved, optional extra
floats (velocities / box) after the positions so that stride != 3*natom.
"""
import numpy as np

SIGMAS = (0.0, 0.05, 0.5, 2.0)
MASS_CYCLE = (12.01, 14.01, 16.00, 32.06, 1.008)


def _rotations(rng, n):
    q = rng.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def base_chain(rng, natom):
    steps = rng.standard_normal((natom, 3))
    steps /= np.linalg.norm(steps, axis=1, keepdims=True)
    return np.cumsum(3.8 * steps, axis=0)


def masses(natom):
    return np.array([MASS_CYCLE[i % len(MASS_CYCLE)] for i in range(natom)], np.float64)


def make_trajectory(seed, nframes, natoms, natom_total=None, stride_extra=0, chunk=4096, out=None):
    """Return (crd float32 [nframes, 3*natom_total + stride_extra], mass float64 [natom_total])."""
    nt = natom_total or natoms
    rng = np.random.default_rng(seed)
    base = base_chain(rng, nt)
    stride = 3 * nt + stride_extra
    crd = out if out is not None else np.empty((nframes, stride), np.float32)
    assert crd.shape == (nframes, stride) and crd.dtype == np.float32
    for f0 in range(0, nframes, chunk):
        f1 = min(nframes, f0 + chunk)
        n = f1 - f0
        R = _rotations(rng, n)
        T = rng.uniform(-20.0, 20.0, (n, 1, 3))
        sig = np.array([SIGMAS[f % len(SIGMAS)] for f in range(f0, f1)])[:, None, None]
        noise = rng.standard_normal((n, nt, 3), dtype=np.float32)
        xyz = np.einsum("fij,aj->fai", R, base) + T
        xyz += sig * noise
        crd[f0:f1, :3 * nt] = xyz.reshape(n, -1).astype(np.float32)
        if stride_extra:
            crd[f0:f1, 3 * nt:] = rng.standard_normal((n, stride_extra), dtype=np.float32)
    dup = np.arange(63, nframes, 64)
    crd[dup] = crd[dup - 1]
    return crd, masses(nt)
