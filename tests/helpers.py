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
