"""Hierarchical agglomerative clustering on the pairwise cache.

CPU part: the C restatement (oracle/rmsd_oracle.c: orc_hieragglo) is pinned against the reference's own
Cluster::DynamicMatrix compiled where it lies (oracle/_ref: ref_hieragglo) -- merge sequence and FindMin values
identical, ties included.  GPU part (-m gpu): b200_hieragglo through the C ABI against the restatement: which pair
merges at every step must be identical (integers: bit-exact), FindMin values bit-equal for single / complete linkage
and for average linkage on these sizes (the sums are exactly representable)."""
import os

import numpy as np
import pytest

from oracle.pyoracle import tri_size


def cache_from_points(rng, n, dim=3, quant=None, dup=0):
    """A float32 distance triangle of n random points (row-major upper triangle = calcTriIndex order)."""
    x = rng.standard_normal((n, dim)) * 3.0
    for k in range(dup):                       # exact duplicates: zero distances and equal rows
        x[rng.integers(n)] = x[rng.integers(n)]
    d = np.sqrt(((x[:, None, :] - x[None, :, :]) ** 2).sum(-1))
    if quant:
        d = np.round(d * quant) / quant
    return d[np.triu_indices(n, 1)].astype(np.float32)


CASES = [(2, None, 0), (3, None, 0), (17, None, 2), (64, 4, 0), (150, None, 5), (150, 2, 0), (333, 16, 8)]


@pytest.mark.parametrize("linkage", [0, 1, 2])
def test_restatement_matches_reference_dynamicmatrix(oracle, reference, linkage):
    rng = np.random.default_rng(100 + linkage)
    for n, quant, dup in CASES:
        tri = cache_from_points(rng, n, quant=quant, dup=dup)
        for target, eps in ((None, None), (5, None), (None, 2.5), (3, 1.0)):
            a = oracle.hieragglo(tri, n, linkage, target, eps)
            b = reference.hieragglo(tri, n, linkage, target, eps)
            for x, y in zip(a, b):
                assert np.array_equal(x, y), (n, quant, dup, target, eps)


def test_restatement_stop_rules(oracle):
    rng = np.random.default_rng(7)
    n = 40
    tri = cache_from_points(rng, n)
    into, frm, fmin = oracle.hieragglo(tri, n, 1, 10, None)
    assert len(into) == n - 10 and len(fmin) == n - 10             # target reached: every call merged
    into, frm, fmin = oracle.hieragglo(tri, n, 1, None, 1.5)
    assert len(fmin) == len(into) + 1 and fmin[-1] > 1.5            # the last call only reported the minimum
    assert np.all(fmin[:-1] <= 1.5) and np.all(into < frm)
    into, frm, fmin = oracle.hieragglo(tri, n, 1, 100, None)        # target above N: one merge is still attempted
    assert len(into) == 1                                           # (DoClustering :153-160)
    assert oracle.hieragglo(np.zeros(0, np.float32), 1, 1)[0].size == 0


# ------------------------------------------------------------------------------------------------ GPU
def assert_same(got, want, what):
    for name, x, y in zip(("mergeInto", "mergeFrom", "findMin"), got, want):
        assert x.shape == y.shape, (what, name, x.shape, y.shape)
        if not np.array_equal(x, y):
            k = int(np.nonzero(x != y)[0][0])
            raise AssertionError("%s: %s differs first at merge %d: got %r want %r" % (what, name, k, x[k], y[k]))


@pytest.mark.gpu
@pytest.mark.parametrize("linkage", [0, 1, 2])
def test_gpu_matches_restatement(b200, oracle, linkage):
    rng = np.random.default_rng(200 + linkage)
    for n, quant, dup in CASES + [(700, None, 3), (700, 8, 0)]:
        tri = cache_from_points(rng, n, quant=quant, dup=dup)
        for target, eps in ((None, None), (5, None), (None, 2.5), (3, 1.0), (10 * n, None)):
            want = oracle.hieragglo(tri, n, linkage, target, eps)
            got = b200.hieragglo(tri, n, linkage, target, eps)
            assert_same(got, want, (n, quant, dup, linkage, target, eps))


@pytest.mark.gpu
def test_gpu_cluster_sizes_agree(b200, oracle, monkeypatch):
    """The merge loop runs in ONE thread-block cluster; every cluster size must give the same merges."""
    rng = np.random.default_rng(300)
    n = 1500
    tri = cache_from_points(rng, n, quant=32, dup=4)
    for linkage in (0, 1, 2):
        want = oracle.hieragglo(tri, n, linkage, 4, None)
        for team in (1, 2, 8, 16):
            monkeypatch.setenv("B200_HA_TEAM", str(team))
            got = b200.hieragglo(tri, n, linkage, 4, None)
            assert_same(got, want, (linkage, team))


@pytest.mark.gpu
def test_gpu_rmsd_cache_end_to_end(b200, oracle):
    """cluster hieragglo on an RMSD cache filled by the device path: same merges as the restatement on the same cache,
    and -- the distances being within 1e-4 A of the reference's -- the same final partition as on the reference's cache
    for well separated conformers."""
    from cpptraj_b200.synth import _rotations
    rng = np.random.default_rng(400)
    nconf, per, na = 6, 40, 60
    confs = rng.standard_normal((nconf, na, 3)) * 4.0
    nf = nconf * per
    R = _rotations(rng, nf)
    xyz = np.einsum("fij,faj->fai", R, confs[np.arange(nf) % nconf]) + rng.uniform(-10, 10, (nf, 1, 3))
    xyz += 0.05 * rng.standard_normal(xyz.shape)
    crd = xyz.reshape(nf, -1).astype(np.float32)
    sel = np.arange(na, dtype=np.int32)
    tri = b200.rms2d_tri(crd, sel)
    tri_ref = oracle.cluster_tri(crd, sel)
    assert np.abs(tri.astype(np.float64) - tri_ref).max() <= 1e-4
    for linkage in (0, 1, 2):
        got = b200.hieragglo(tri, nf, linkage, nconf, None)
        assert_same(got, oracle.hieragglo(tri, nf, linkage, nconf, None), linkage)
        lab = partition(got, nf)
        lab_ref = partition(oracle.hieragglo(tri_ref, nf, linkage, nconf, None), nf)
        assert np.array_equal(lab, lab_ref)
        assert all(len(set(lab[np.arange(nf) % nconf == c])) == 1 for c in range(nconf))


@pytest.mark.gpu
def test_gpu_tight_cluster_long_rescan_lists(b200, oracle):
    """A quarter of the frames are rigid copies of one conformation (RMSDs ~1e-6 A apart), every 64th frame is a bit-exact
    duplicate: the tight cluster is the closest of almost every other cluster, so merging it away queues hundreds of
    rows for a re-scan -- the long-list path (one warp per row), the lower-bound shortcut that skips re-scans whose
    outcome is certain, zero distances and exact ties, all against the restatement."""
    from cpptraj_b200.synth import make_trajectory
    nf, na = 900, 40
    crd, _ = make_trajectory(31, nf, na)
    tri = b200.rms2d_tri(crd, np.arange(na, dtype=np.int32))
    assert (tri < 1e-5).sum() >= nf // 64                    # the duplicates and the rigid copies
    for linkage in (0, 1, 2):
        for target, eps in ((1, None), (7, None), (None, 1.0)):
            want = oracle.hieragglo(tri, nf, linkage, target, eps)
            for team in (None, 4):
                if team is None:
                    os.environ.pop("B200_HA_TEAM", None)
                else:
                    os.environ["B200_HA_TEAM"] = str(team)
                try:
                    got = b200.hieragglo(tri, nf, linkage, target, eps)
                finally:
                    os.environ.pop("B200_HA_TEAM", None)
                assert_same(got, want, (linkage, target, eps, team))


@pytest.mark.gpu
def test_gpu_cache_consumers(b200):
    """Sums over cluster members (BestReps cumulative distance: sequential, in list order, bit-equal to the reference's
    inner loop; the within-cluster numerators of Summary) and the linkage tables between clusters from one pass over the
    triangle; with and without the cache announced as resident."""
    rng = np.random.default_rng(500)
    n = 700
    tri = cache_from_points(rng, n, dup=6)
    sq = np.zeros((n, n), np.float32)
    sq[np.triu_indices(n, 1)] = tri
    sq = sq + sq.T
    perm = rng.permutation(n)
    lists = [perm[:300], perm[300:301], perm[301:303], perm[303:650]]          # 50 frames in no cluster
    def want_sums():
        cum, up, up2 = [], [], []
        for m in lists:
            for a, i in enumerate(m):
                s = u = u2 = 0.0
                for b_, j in enumerate(m):
                    if b_ == a:
                        continue
                    d = float(sq[i, j])
                    s += d
                    if b_ > a:
                        u += d; u2 += d * d
                cum.append(s); up.append(u); up2.append(u2)
        return np.array(cum), np.array(up), np.array(up2)
    wc, wu, wu2 = want_sums()
    label = np.full(n, -1, np.int32)
    for c, m in enumerate(lists):
        label[m] = c
    for resident in (False, True):
        t = b200.cache_resident_begin(tri, n) if resident else tri
        cum, up, up2 = b200.cache_cluster_sums(t, n, lists)
        assert np.array_equal(cum, wc) and np.array_equal(up, wu) and np.array_equal(up2, wu2)
        mn, mx, sm, cnt = b200.cache_cluster_links(t, n, label, len(lists))
        for c1 in range(len(lists)):
            for c2 in range(c1 + 1, len(lists)):
                blk = sq[np.ix_(lists[c1], lists[c2])].astype(np.float64)
                assert cnt[c1, c2] == blk.size and mn[c1, c2] == blk.min() and mx[c1, c2] == blk.max()
                assert abs(sm[c1, c2] - blk.sum()) <= 1e-9 * blk.sum()
        got = b200.hieragglo(t, n, 1, 5, None)                                  # the resident copy is only read
        if resident:
            b200.cache_resident_end(t)
            assert_same(got, first, "resident cache")
        else:
            first = got


def partition(merges, n):
    """Cluster label (lowest member) of every frame after replaying the merges."""
    parent = np.arange(n)
    for a, b in zip(merges[0], merges[1]):
        parent[parent == b] = a
    return parent
