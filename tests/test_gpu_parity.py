"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle
(bit-pinned to the reference, tests/test_oracle.py).  Tolerance: 1e-4 A absolute per pair
(BASELINE.json north_star); exact best-frame index for one-vs-many."""
import numpy as np
import pytest

from helpers import TOL, synth_case, tri_to_square, check_argmin, exact_fit_rmsd

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def fp64_engine(b200):
    """This module pins the FP64 DMMA engine (some asserts are far tighter than the 1e-4 A contract);
    the tcgen05 int8 engine and the automatic choice are covered by tests/test_gpu_i8.py."""
    b200.set_pair_engine("fp64")
    yield
    b200.set_pair_engine("auto")


def maxdiff(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max()) if np.size(a) else 0.0


# ---------------------------------------------------------------- reference decks (tz2)
def test_tz2_2drms_goldens(b200, tz2, saves, live):
    crd, mass = tz2["crd"], tz2["mass"]
    s37 = tz2["res"](3, 7)
    out = b200.rms2d_tri(crd[:10], s37)
    assert np.array_equal(np.round(tri_to_square(out, 10), 3), saves["rmsd"])
    assert maxdiff(out, live["tz2_3to7_fit"]) <= 1e-6
    out = b200.rms2d_tri(crd[:10], s37, mass=mass[s37])
    assert np.array_equal(np.round(tri_to_square(out, 10), 3), saves["rmsd_mass"])
    assert maxdiff(out, live["tz2_3to7_mass"]) <= 1e-6
    full = b200.rms2d_full(crd[:10], tz2["res"](2, 2), crd[:10], tz2["res"](11, 11))
    assert maxdiff(np.round(full.astype(np.float64), 3), saves["trp"]) < 1e-6
    assert maxdiff(full, live["tz2_trp_full"]) <= 1e-6


def test_tz2_config0_all_ca(b200, tz2, live, oracle):
    """BASELINE configs[0]: rms2d @CA on the tz2 trajectory (101 frames x 12 atoms)."""
    allca = np.nonzero(tz2["names"] == "CA")[0].astype(np.int32)
    out = b200.rms2d_tri(tz2["crd"], allca)
    assert maxdiff(out, live["tz2_allca_fit"]) <= 1e-6
    m = tz2["mass"][allca]
    out = b200.rms2d_tri(tz2["crd"], allca, mass=m)
    assert maxdiff(out, live["tz2_allca_cluster_mass"]) <= 2e-6   # cluster flavour: same number, roles swapped
    assert maxdiff(out, oracle.rms2d_tri(tz2["crd"], allca, mass=m)) <= 1e-6


def test_tz2_rms_nomod_and_argmin(b200, tz2, saves, live):
    ca = np.nonzero((tz2["names"] == "CA") & (tz2["resnum"] >= 2) & (tz2["resnum"] <= 12))[0].astype(np.int32)
    ref = tz2["crd"][0].reshape(-1, 3)[ca].astype(np.float64)
    ref -= ref.mean(0)                       # ReferenceAction hands over the centred reference
    r, _, _, best = b200.rmsd_1vN(tz2["crd"], ca, ref)
    assert np.array_equal(np.round(r, 4), saves["nomod"])
    assert maxdiff(r, live["tz2_ca_1vN"]) <= 1e-6   # frame 0 vs itself: reference 1.6e-7, exact 0
    check_argmin(best, live["tz2_ca_1vN"])


# ---------------------------------------------------------------- synthetic vs oracle
@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_fit_all_mma_variants(b200, oracle, variant):
    b200.set_mma_variant(variant)
    try:
        c, m, sel = synth_case(21, 200, 1000)
        assert maxdiff(b200.rms2d_tri(c, sel), oracle.rms2d_tri(c, sel)) <= TOL
    finally:
        b200.set_mma_variant(3)


@pytest.mark.parametrize("nf,na,ntot,extra", [
    (2, 3, 3, 0), (3, 1, 5, 0), (33, 7, 20, 0), (31, 16, 16, 0), (32, 17, 40, 7), (65, 15, 15, 0),
    (64, 64, 64, 0), (97, 130, 400, 0), (130, 333, 333, 1000), (300, 1000, 1000, 0), (257, 2000, 2100, 0),
])
def test_fit_mass_nofit_ragged_shapes(b200, oracle, nf, na, ntot, extra):
    c, m, sel = synth_case(100 + nf + na, nf, na, ntot, extra)
    assert maxdiff(b200.rms2d_tri(c, sel), oracle.rms2d_tri(c, sel)) <= TOL
    assert maxdiff(b200.rms2d_tri(c, sel, mass=m[sel]), oracle.rms2d_tri(c, sel, mass=m[sel])) <= TOL
    assert maxdiff(b200.rms2d_tri(c, sel, fit=False), oracle.rms2d_tri(c, sel, fit=False)) <= TOL
    assert maxdiff(b200.rms2d_tri(c, sel, mass=m[sel], fit=False), oracle.rms2d_tri(c, sel, mass=m[sel], fit=False)) <= TOL


def test_live_reference_fixtures(b200, live):
    for tag, seed, nf, na, ntot in (("s1", 11, 48, 100, 100), ("s2", 12, 40, 257, 300), ("s3", 13, 33, 7, 20)):
        c, m, sel = synth_case(seed, nf, na, ntot)
        assert maxdiff(b200.rms2d_tri(c, sel), live[tag + "_fit"]) <= TOL
        assert maxdiff(b200.rms2d_tri(c, sel, mass=m[sel]), live[tag + "_mass"]) <= TOL
        assert maxdiff(b200.rms2d_tri(c, sel, fit=False), live[tag + "_nofit"]) <= TOL
        ref = c[0].reshape(-1, 3)[sel].astype(np.float64)
        ref -= (m[sel, None] * ref).sum(0) / m[sel].sum()
        r, rot, tr, best = b200.rmsd_1vN(c, sel, ref, mass=m[sel], want_rot=True)
        assert maxdiff(r, live[tag + "_1vN_rms"]) <= 1e-5   # zero-RMSD pairs carry sqrt(eps*E0/M) noise on both sides
        ok = live[tag + "_1vN_rms"] > 1e-3      # rotation of a zero-RMSD (identical) pair is ill-defined only in sign conventions
        assert maxdiff(rot[ok], live[tag + "_1vN_rot"][ok]) <= 1e-6
        assert maxdiff(tr, live[tag + "_1vN_tr"]) <= 1e-9
        check_argmin(best, live[tag + "_1vN_rms"])


def test_single_frame_and_empty(b200):
    c, m, sel = synth_case(3, 4, 10)
    assert b200.rms2d_tri(c[:1], sel).size == 0
    out = b200.rms2d_tri(np.vstack([c[:1], c[:1]]), sel)
    assert out.shape == (1,) and out[0] < 1e-5


def test_exact_duplicates_are_zero(b200, oracle):
    c, m, sel = synth_case(8, 130, 500)          # frames 63 and 127 duplicate their predecessors
    out = tri_to_square(b200.rms2d_tri(c, sel), 130)
    assert out[62, 63] < 2e-5 and out[126, 127] < 2e-5
    assert maxdiff(b200.rms2d_tri(c, sel), oracle.rms2d_tri(c, sel)) <= TOL


def test_sieve_frame_index(b200, oracle):
    """Cluster path: framesToCache is an arbitrary (sieved, even shuffled) frame list."""
    c, m, sel = synth_case(31, 150, 120, 200)
    fidx = np.arange(3, 150, 5, dtype=np.int32)
    assert maxdiff(b200.rms2d_tri(c, sel, mass=m[sel], frame_idx=fidx),
                   oracle.cluster_tri(c, sel, mass=m[sel], frame_idx=fidx)) <= TOL
    rng = np.random.default_rng(0)
    fidx = rng.permutation(150).astype(np.int32)[:77]
    assert maxdiff(b200.rms2d_tri(c, sel, frame_idx=fidx), oracle.rms2d_tri(c, sel, frame_idx=fidx)) <= TOL


def test_full_matrix_masks_and_masses(b200, oracle):
    c, m, sel = synth_case(41, 70, 90, 250)
    sel2 = (sel + 3).astype(np.int32)
    c2, _, _ = synth_case(42, 45, 90, 250)
    for fit in (True, False):
        got = b200.rms2d_full(c, sel, c2, sel2, fit=fit)
        assert got.shape == (70, 45)
        assert maxdiff(got, oracle.rms2d_full(c, sel, c2, sel2, fit=fit)) <= TOL
        got = b200.rms2d_full(c, sel, c2, sel2, mass_tgt=m[sel], mass_ref=m[sel2], fit=fit)
        assert maxdiff(got, oracle.rms2d_full(c, sel, c2, sel2, massTgt=m[sel], massRef=m[sel2], fit=fit)) <= TOL


def test_shards_tile_the_triangle(b200, oracle):
    c, m, sel = synth_case(51, 260, 64)
    want = oracle.rms2d_tri(c, sel)
    for count in (2, 3, 8):
        out = np.full(want.size, -7.0, np.float32)
        covered = 0
        for r in range(count):
            _, first, n = b200.rms2d_tri_shard(c, sel, r, count, out=out)
            assert first == covered
            covered += n
        assert covered == want.size
        assert maxdiff(out, want) <= TOL


def test_device_api_matches_host_api(b200, oracle):
    import torch
    c, m, sel = synth_case(61, 190, 210, 230, 11)
    want = oracle.rms2d_tri(c, sel, mass=m[sel])
    dc = torch.from_numpy(c).cuda()
    ds = torch.from_numpy(sel).cuda()
    dm = torch.from_numpy(m[sel].copy()).cuda()
    out = torch.full((want.size,), -1.0, dtype=torch.float32, device="cuda")
    b200.dev_rms2d_tri(dc, c.shape[1], c.shape[0], ds, len(sel), out, d_mass=dm,
                       stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert maxdiff(out.cpu().numpy(), want) <= TOL
    # sharded device call fills only its rows
    out2 = torch.full((want.size,), -1.0, dtype=torch.float32, device="cuda")
    for r in range(2):
        b200.dev_rms2d_tri(dc, c.shape[1], c.shape[0], ds, len(sel), out2, d_mass=dm, rank=r, count=2,
                           stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert maxdiff(out2.cpu().numpy(), want) <= TOL


def test_two_atom_selection(b200, oracle):
    """2 atoms => rank-1 covariance (see test_degenerate_geometries): exact value, loose vs the reference."""
    c, m, sel = synth_case(107, 5, 2)
    X = c.reshape(5, 2, 3)
    for mass in (None, m[sel]):
        got = tri_to_square(b200.rms2d_tri(c, sel, mass=mass), 5)
        exact = np.array([[exact_fit_rmsd(X[i], X[j], mass) if i != j else 0.0 for j in range(5)] for i in range(5)])
        assert maxdiff(got, exact) <= TOL
        assert maxdiff(got, tri_to_square(oracle.rms2d_tri(c, sel, mass=mass), 5)) <= 5e-3
    assert maxdiff(b200.rms2d_tri(c, sel, fit=False), oracle.rms2d_tri(c, sel, fit=False)) <= TOL


def test_degenerate_geometries(b200, oracle):
    """Collinear / planar / mirrored selections: the quartic has (near-)double roots there."""
    rng = np.random.default_rng(9)
    nf, na = 40, 30
    t = rng.uniform(-20, 20, (na, 1))
    line = t * np.array([[0.3, -0.5, 0.81]])                       # collinear atoms (rank-1 covariance)
    plane = np.c_[rng.uniform(-15, 15, (na, 2)), np.zeros(na)]     # planar atoms (rank-2)
    blob = rng.uniform(-10, 10, (na, 3))
    for base in (line, plane, blob):
        frames = []
        for f in range(nf):
            q = rng.standard_normal(4); q /= np.linalg.norm(q)
            w, x, y, z = q
            R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                          [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                          [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
            xyz = base @ R.T + rng.uniform(-5, 5, (1, 3)) + (f % 4) * 0.05 * rng.standard_normal((na, 3))
            if f % 5 == 4:
                xyz = xyz * np.array([[1.0, 1.0, -1.0]])            # mirror image: det(S) < 0
            frames.append(xyz.reshape(-1))
        c = np.array(frames, np.float32)
        sel = np.arange(na, dtype=np.int32)
        m = rng.uniform(1, 32, na)
        if base is not blob:
            # rank-deficient covariance (rank 1 for the line, rank 2 for the plane): the reference takes the
            # sqrt of eigenvalues of the SQUARED matrix, so a zero singular value comes back as
            # sqrt(rounding noise) and its own error is ~1e-3 A for near-duplicate frames (measured: it
            # returns 7.0e-4 A for a rotated copy of a planar frame whose exact RMSD is 3e-7 A).  Check
            # against a float64 SVD evaluation and require the oracle to agree with us to ITS accuracy.
            X = c.reshape(nf, na, 3)
            for mass in (None, m):
                got = tri_to_square(b200.rms2d_tri(c, sel, mass=mass), nf)
                exact = np.array([[exact_fit_rmsd(X[i], X[j], mass) if i != j else 0.0 for j in range(nf)] for i in range(nf)])
                assert maxdiff(got, exact) <= TOL
                assert maxdiff(got, tri_to_square(oracle.rms2d_tri(c, sel, mass=mass), nf)) <= 5e-3
        else:
            assert maxdiff(b200.rms2d_tri(c, sel), oracle.rms2d_tri(c, sel)) <= TOL
            assert maxdiff(b200.rms2d_tri(c, sel, mass=m), oracle.rms2d_tri(c, sel, mass=m)) <= TOL


# ---------------------------------------------------------------- one-vs-many
def test_one_vs_many_streaming(b200, oracle):
    c, m, sel = synth_case(71, 500, 700, 800, 13)
    ref_raw = c[17].reshape(-1)[:2400].reshape(-1, 3)[sel].astype(np.float64)
    for mass in (None, m[sel]):
        w = np.ones(len(sel)) if mass is None else mass
        ref = ref_raw - (w[:, None] * ref_raw).sum(0) / w.sum()
        want, wrot, wtr, _ = oracle.rmsd_1vN(c, sel, ref_raw, mass=mass, want_rot=True)
        # float COORDS pushed in uneven chunks
        with b200.Rmsd1vN(ref, sel, mass, True, True) as h:
            for a, e in ((0, 1), (1, 130), (130, 131), (131, 500)):
                h.push(c[a:e])
            r, rot, tr, best = h.flush()
        assert maxdiff(r, want) <= 1e-5
        check_argmin(best, want)
        ok = want > 1e-3
        assert maxdiff(rot[ok], wrot[ok]) <= 1e-6 and maxdiff(tr, wtr) <= 1e-9
        # double Frames (cpptraj's Frame::xAddress()), pageable
        cd = np.ascontiguousarray(c[:, :2400], np.float64)
        r2, _, _, best2 = b200.rmsd_1vN(cd, sel, ref, mass=mass, chunk=64)
        assert maxdiff(r2, want) <= 1e-5 and best2 == best
        # no-fit: reference passed raw
        wantnf = oracle.rmsd_1vN(c, sel, ref_raw, mass=mass, fit=False)
        rnf, _, _, _ = b200.rmsd_1vN(c, sel, ref_raw, mass=mass, fit=False)
        assert maxdiff(rnf, wantnf) <= 1e-9


def test_one_vs_many_pinned_and_device(b200, oracle):
    import torch
    c, m, sel = synth_case(81, 300, 5000)
    ref_raw = c[0].reshape(-1, 3)[sel].astype(np.float64)
    ref = ref_raw - ref_raw.mean(0)
    want = oracle.rmsd_1vN(c, sel, ref_raw)
    pinned = torch.from_numpy(c).pin_memory()
    r, _, _, best = b200.rmsd_1vN(pinned.numpy(), sel, ref)
    assert maxdiff(r, want) <= 2e-5
    check_argmin(best, want)
    dc = pinned.cuda()
    out = torch.empty(300, dtype=torch.float64, device="cuda")
    b200.dev_rmsd_1vN(dc, c.shape[1], 300, torch.from_numpy(sel).cuda(), len(sel), torch.from_numpy(ref).cuda(), out,
                      stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert maxdiff(out.cpu().numpy(), want) <= 2e-5


# ---------------------------------------------------------------- errors
def test_argument_errors(b200):
    c, m, sel = synth_case(3, 4, 10)
    with pytest.raises(b200.B200Error):
        b200.rms2d_tri(c, np.array([0, 1, 99], np.int32))          # atom outside the frame
    with pytest.raises(b200.B200Error):
        b200.rms2d_tri(c, sel, frame_idx=np.array([0, 9], np.int32))  # frame outside the set


# ---------------------------------------------------------------- full-size properties (BASELINE configs[1])
def test_config2_full_size_properties(b200, oracle):
    """10k frames x 1k atoms: too big for the oracle, so check (1) a random 96-frame sub-triangle
    against the oracle, (2) duplicates are 0, (3) permutation invariance of the pair values."""
    from cpptraj_b200.synth import make_trajectory
    nf, na = 10000, 1000
    crd, mass = make_trajectory(20261017, nf, na)
    sel = np.arange(na, dtype=np.int32)
    out = b200.rms2d_tri(crd, sel)
    assert out.size == nf * (nf - 1) // 2 and np.isfinite(out).all() and out.min() >= 0.0
    rng = np.random.default_rng(5)
    pick = np.sort(rng.choice(nf, 96, replace=False)).astype(np.int32)
    sub = oracle.rms2d_tri(crd, sel, frame_idx=pick)
    ii, jj = np.triu_indices(96, 1)
    I, J = pick[ii].astype(np.int64), pick[jj].astype(np.int64)
    got = out[nf * I - (I + 1) * I // 2 + J - I - 1]
    assert maxdiff(got, sub) <= TOL
    dup = np.arange(63, nf, 64, dtype=np.int64)
    d = out[nf * (dup - 1) - dup * (dup - 1) // 2 + 0]       # element (dup-1, dup)
    assert d.max() < 2e-5
    perm = rng.permutation(nf).astype(np.int32)[:2048]
    a = b200.rms2d_tri(crd, sel, frame_idx=perm)
    pi, pj = np.triu_indices(2048, 1)
    P, Q = perm[pi].astype(np.int64), perm[pj].astype(np.int64)
    lo, hi = np.minimum(P, Q), np.maximum(P, Q)
    assert maxdiff(a, out[nf * lo - (lo + 1) * lo // 2 + hi - lo - 1]) <= 2e-6


@pytest.mark.parametrize("fit,use_mass", [(True, False), (True, True), (False, False)])
def test_frames_to_centroids(b200, oracle, fit, use_mass):
    """Metric_RMS::FrameCentroidDist for many frames (src/Cluster/Metric_RMS.cpp:75-81) and the nearest
    centroid per frame as List::AddFramesByCentroid picks it (first minimum wins, src/Cluster/List.cpp:183-189)."""
    c, m, sel = synth_case(4242, 301, 150, 170, 3)
    mass = m[sel] if use_mass else None
    K = 5
    # centroids: selected atoms of a few frames, averaged with a neighbour and centred (cpptraj keeps centroids at the origin)
    X = c[:, : 3 * 170].reshape(301, 170, 3)[:, sel].astype(np.float64)
    w = np.ones(len(sel)) if mass is None else mass
    cen = []
    for k in range(K):
        a = 0.5 * (X[40 * k] + X[40 * k + 1])
        if fit:
            a = a - (w[:, None] * a).sum(0) / w.sum()
        cen.append(a)
    cen = np.array(cen)
    fidx = np.arange(2, 301, 2, dtype=np.int32)[::-1].copy()          # a sieved, unordered frame list
    for frame_idx in (None, fidx):
        dist, closest, cdist = b200.frames_to_centroids(c, sel, cen, mass=mass, fit=fit, frame_idx=frame_idx)
        frames = c if frame_idx is None else c[frame_idx]
        want = np.stack([oracle.rmsd_1vN(frames, sel, cen[k], mass=mass, fit=fit) for k in range(K)], axis=1)
        assert dist.shape == want.shape
        assert np.abs(dist - want).max() <= TOL
        assert np.abs(cdist - want.min(1)).max() <= TOL
        # nearest centroid: exact where the two best are separated by more than the tolerance
        srt = np.sort(want, axis=1)
        clear = (srt[:, 1] - srt[:, 0]) > 2 * TOL
        assert np.array_equal(closest[clear], want.argmin(1)[clear])
        assert np.array_equal(closest, dist.argmin(1))               # first minimum of the table it returned
