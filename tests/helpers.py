import numpy as np


def tri_to_square(tri, n):
    m = np.zeros((n, n), np.float64)
    iu = np.triu_indices(n, 1)
    m[iu] = tri            # calcTriIndex order == row-major upper triangle
    m.T[iu] = tri
    return m


def synth_case(seed, nf, na, ntot=None, extra=0):
    from cpptraj_b200.synth import make_trajectory
    ntot = ntot or na
    crd, mass = make_trajectory(seed, nf, na, natom_total=ntot, stride_extra=extra)
    if ntot != na:
        sel = np.arange(0, ntot, max(1, ntot // na), dtype=np.int32)[:na]
    else:
        sel = np.arange(na, dtype=np.int32)
    return crd, mass, sel


TOL = 1e-4  # Angstrom, absolute per pair (BASELINE.json north_star)


def check_argmin(best, want, tol=TOL):
    """north_star: the best-fit frame index must match exactly.  Exact when the minimum is
    separated from the runner-up by more than the contract tolerance; for numerical ties
    (e.g. the reference frame itself vs a rotated bit-copy, both ~1e-7 A) any tied frame."""
    want = np.asarray(want)
    order = np.argsort(want, kind="stable")
    if len(want) > 1 and want[order[1]] - want[order[0]] > tol:
        assert best == int(order[0]), (best, int(order[0]))
    else:
        assert want[best] - want[order[0]] <= tol, (best, int(order[0]))


def exact_fit_rmsd(x, y, w=None):
    """float64 SVD evaluation of the best-fit RMSD (no squaring of the covariance): used only
    where the reference's own Jacobi-on-RR^T arithmetic is unstable (collinear selections)."""
    x = np.asarray(x, np.float64); y = np.asarray(y, np.float64)
    w = np.ones(len(x)) if w is None else np.asarray(w, np.float64)
    x = x - (w[:, None] * x).sum(0) / w.sum()
    y = y - (w[:, None] * y).sum(0) / w.sum()
    S = (w[:, None] * x).T @ y
    sv = np.linalg.svd(S, compute_uv=False)
    lam = sv[0] + sv[1] + (sv[2] if np.linalg.det(S) >= 0 else -sv[2])
    e0 = 0.5 * ((w[:, None] * x * x).sum() + (w[:, None] * y * y).sum())
    return np.sqrt(max(0.0, 2.0 * (e0 - lam) / w.sum()))
