"""GPU parity tests (through the C ABI) of what round 2 added: the reference's total-mass guard, the limits of the
tcgen05 engine's eligibility (extent, atom count), pageable host buffers, the pinned fixed-point scale, reference
changes on a live one-vs-many handle, frames x centroids as one contraction, several devices in one process.
Tolerance: 1e-4 A absolute (BASELINE.json north_star); argmin exact."""
import numpy as np
import pytest

from helpers import TOL, check_argmin, synth_case

pytestmark = pytest.mark.gpu


def maxdiff(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())


def rigid_copies(rng, base, nf, sigmas, shift=20.0):
    """Frames = random rigid motions of `base` (+ per-frame gaussian noise); float32 COORDS rows."""
    from cpptraj_b200.synth import _rotations
    R = _rotations(rng, nf)
    T = rng.uniform(-shift, shift, (nf, 1, 3))
    xyz = np.einsum("fij,aj->fai", R, base) + T
    xyz += np.asarray(sigmas)[:, None, None] * rng.standard_normal(xyz.shape)
    return xyz.reshape(nf, -1).astype(np.float32)


# ---------------------------------------------------------------- src/Frame.cpp:1160-1163, :1300-1303
def test_total_mass_below_small_is_minus_one(b200, oracle):
    c, m, sel = synth_case(5, 40, 30)
    zero = np.zeros(len(sel))
    for fit in (True, False):
        want = oracle.rms2d_tri(c, sel, mass=zero, fit=fit)
        assert np.all(want == -1.0)                       # what the reference reports
        got = b200.rms2d_tri(c, sel, mass=zero, fit=fit)
        assert np.array_equal(got, want)
    ref = c[0].reshape(-1, 3)[sel].astype(np.float64)
    for fit in (True, False):
        r, _, _, _ = b200.rmsd_1vN(c, sel, ref, mass=zero, fit=fit)
        assert np.all(r == -1.0)
    tiny = np.full(len(sel), 1e-17)                       # sum 3e-16 < Constants::SMALL
    assert np.all(b200.rms2d_tri(c, sel, mass=tiny) == -1.0)
    full = b200.rms2d_full(c[:5], sel, c, sel, mass_tgt=zero, mass_ref=zero)
    assert np.all(full == -1.0)


# ---------------------------------------------------------------- eligibility of the tcgen05 engine
def test_i8_eligibility_edge_of_extent(b200, oracle):
    """Unit masses: qs >= 15 fractional bits are needed (sqrt(3) 2^-qs <= 8.5e-5), i.e. centred coordinates up to 255 A.
    An extended 300-atom object whose largest centred radius is 250 A sits just inside: it must take the tcgen05
    engine with the minimum number of bits and still meet the contract -- near-duplicates and exact duplicates
    included; the measured error is compared with the bound the eligibility rule promises."""
    rng = np.random.default_rng(20261021)
    base = rng.standard_normal((300, 3)) * np.array([90.0, 25.0, 8.0])     # elongated, far from spherical
    base -= base.mean(0)
    base *= 250.0 / np.linalg.norm(base, axis=1).max()
    nf = 180
    sig = np.array([(0.0, 1e-3, 0.05, 0.5, 2.0)[f % 5] for f in range(nf)])
    c = rigid_copies(rng, base, nf, sig, shift=5.0)
    c[7] = c[6]; c[100] = c[3]                                              # exact duplicates
    sel = np.arange(300, dtype=np.int32)
    b200.set_pair_engine("auto")
    got = b200.rms2d_tri(c, sel)
    eng, qs = b200.last_pair_engine()
    assert eng == 2 and qs == 15, (eng, qs)
    want = oracle.rms2d_tri(c, sel)
    err = maxdiff(got, want)
    assert err <= 8.5e-5 + 4e-6, err            # rounding bound + float32 store
    assert err <= TOL
    n = nf      # exact duplicates, frames (6, 7) and (3, 100): zero up to the FP64 noise of E0 - lambda (1e-16 E0, i.e. ~1e-8 Rg)
    assert got[n * 6 - 7 * 6 // 2 + 7 - 6 - 1] <= 1e-5 and got[n * 3 - 4 * 3 // 2 + 100 - 3 - 1] <= 1e-5
    # the same object 8 % larger: one coordinate beyond 255 A -> too few bits -> FP64 engine, same contract
    big = base * 1.08
    cb = rigid_copies(rng, big, 60, sig[:60], shift=5.0)
    cb[0] = (big + np.array([1.0, 2.0, 3.0])).reshape(-1).astype(np.float32)   # an unrotated frame: radius 270 A along x
    gotb = b200.rms2d_tri(cb, sel)
    engb, _ = b200.last_pair_engine()
    if np.abs(cb.reshape(60, -1, 3) - cb.reshape(60, -1, 3).mean(1, keepdims=True)).max() > 255.0:
        assert engb == 1
        with pytest.raises(b200.B200Error):
            b200.set_pair_engine("i8")
            try:
                b200.rms2d_tri(cb, sel)
            finally:
                b200.set_pair_engine("auto")
    assert maxdiff(gotb, oracle.rms2d_tri(cb, sel)) <= TOL


def test_i8_atom_count_limit(b200, oracle):
    """int32 TMEM accumulators are exact below 2^17 atoms: larger selections must go to the FP64 engine."""
    rng = np.random.default_rng(3)
    na = 130000
    base = rng.standard_normal((na, 3)) * 30.0
    c = rigid_copies(rng, base, 6, np.array([0.0, 0.1, 0.5, 1.0, 2.0, 0.0]))
    sel = np.arange(na, dtype=np.int32)
    got = b200.rms2d_tri(c, sel)
    assert b200.last_pair_engine()[0] == 1
    assert maxdiff(got, oracle.rms2d_tri(c, sel)) <= TOL
    b200.set_pair_engine("i8")
    try:
        with pytest.raises(b200.B200Error):
            b200.rms2d_tri(c, sel)
    finally:
        b200.set_pair_engine("auto")
    # just below the limit the tcgen05 engine runs and is exact enough
    sel2 = np.arange(na - 1, dtype=np.int32)
    got2 = b200.rms2d_tri(c, sel2)
    assert b200.last_pair_engine()[0] == 2
    assert maxdiff(got2, oracle.rms2d_tri(c, sel2)) <= TOL


# ---------------------------------------------------------------- host buffers as cpptraj passes them
def test_pageable_and_pinned_host_buffers_agree(b200, oracle):
    """cpptraj's COORDS vector and Matrix<float> are pageable; bench.py's are pinned.  Both go through the pipelined
    path (>= 1024 frames) and must give the same bits; the pageable result also against the oracle (sampled rows)."""
    import torch
    c, m, sel = synth_case(91, 2300, 120, 130, 2)
    n = 2300
    got_pageable = b200.rms2d_tri(c, sel, mass=m[sel])
    eng = b200.last_pair_engine()
    pc = torch.from_numpy(c).pin_memory()
    po = torch.empty(n * (n - 1) // 2, dtype=torch.float32).pin_memory()
    got_pinned = b200.rms2d_tri(pc.numpy(), sel, mass=m[sel], out=po.numpy())
    assert b200.last_pair_engine() == eng and eng[0] == 2
    assert np.array_equal(got_pageable, got_pinned)
    rows = [0, 1, 517, 1023, 1024, 2298]
    for i in rows:
        idx = n * i - (i + 1) * i // 2 + np.arange(i + 1, n) - i - 1
        pick = np.concatenate([[i], np.arange(i + 1, n)]).astype(np.int32)[:200]
        w = oracle.rms2d_tri(c, sel, mass=m[sel], frame_idx=pick)[: len(pick) - 1]      # row 0 of the sub-triangle
        assert maxdiff(got_pageable[idx[: len(pick) - 1]], w) <= TOL
    # a frame list that is the identity (cluster without sieve) takes the same path and gives the same bits
    ident = np.arange(n, dtype=np.int32)
    assert np.array_equal(b200.rms2d_tri(c, sel, mass=m[sel], frame_idx=ident), got_pageable)
    # full matrix into a pageable array
    full = b200.rms2d_full(c[:300], sel, c[1000:1400], sel)
    assert maxdiff(full, oracle.rms2d_full(c[:300], sel, c[1000:1400], sel)) <= TOL


def test_whole_frame_selection_takes_the_linear_upload(b200, oracle):
    """Pinned COORDS whose selected span is the whole frame (span == frame stride) are uploaded by one linear DMA per chunk,
    any other layout by a pitched copy: same bits either way, and the oracle's values; one-vs-many likewise."""
    import torch
    nf, na = 1600, 96
    c, m, sel = synth_case(77, nf, na)                       # stride == 3 * na: whole frames
    pc = torch.from_numpy(c).pin_memory()
    got = b200.rms2d_tri(pc.numpy(), sel, mass=m[sel])
    assert b200.last_pair_engine()[0] == 2
    wide = np.zeros((nf, 3 * na + 5), np.float32); wide[:, : 3 * na] = c     # same frames, 5 floats of padding per frame
    pw = torch.from_numpy(wide).pin_memory()
    assert np.array_equal(b200.rms2d_tri(pw.numpy(), sel, mass=m[sel]), got)
    assert np.array_equal(b200.rms2d_tri(c, sel, mass=m[sel]), got)         # pageable
    sub = np.r_[0:25, 790:815, nf - 25:nf]
    want = oracle.rms2d_tri(c[sub], sel, mass=m[sel])
    sq = np.zeros((nf, nf), np.float32); sq[np.triu_indices(nf, 1)] = got
    assert np.abs(sq[np.ix_(sub, sub)][np.triu_indices(len(sub), 1)].astype(np.float64) - want).max() <= TOL
    ref_raw = c[3].reshape(-1, 3)[sel].astype(np.float64)
    w = m[sel]
    ref = ref_raw - (w[:, None] * ref_raw).sum(0) / w.sum()      # the action hands over a centred reference
    r_lin = b200.rmsd_1vN(pc.numpy(), sel, ref, mass=w)[0]
    r_pitch = b200.rmsd_1vN(pw.numpy(), sel, ref, mass=w)[0]
    assert np.array_equal(r_lin, r_pitch)
    assert np.abs(r_lin - oracle.rmsd_1vN(c, sel, ref_raw, mass=w)).max() <= TOL


def test_shards_agree_on_the_fixed_point_grid(b200, oracle):
    """Every shard of a matrix takes its scale from the same (top) frames: the shards of a 4-way split, computed one
    after the other, reproduce the single-shard result bit for bit; pinning fewer bits changes the result within the
    bound; pinning more bits than the extent allows is refused."""
    c, m, sel = synth_case(17, 4200, 64)
    n = 4200
    whole = b200.rms2d_tri(c, sel)
    eng, qs = b200.last_pair_engine()
    assert eng == 2
    out = np.zeros_like(whole)
    for r in range(4):
        _, first, cnt = b200.rms2d_tri_shard(c, sel, r, 4, out=out)
        assert b200.last_pair_engine() == (2, qs)
    assert np.array_equal(out, whole)
    b200.set_fixed_point_bits(qs - 1)
    try:
        coarse = b200.rms2d_tri(c, sel)
        assert b200.last_pair_engine() == (2, qs - 1)
        assert 0 < maxdiff(coarse, whole) <= 2 * np.sqrt(3.0) * 2.0 ** -(qs - 1)
        b200.set_fixed_point_bits(30)
        with pytest.raises(b200.B200Error):
            b200.rms2d_tri(c, sel)
    finally:
        b200.set_fixed_point_bits(0)


# ---------------------------------------------------------------- one-vs-many
def test_one_vs_many_reference_changes_and_buffer_reuse(b200, oracle):
    """reftraj / previous (src/ReferenceAction.h:73-88): the reference changes between pushes; a pinned frame buffer is
    overwritten right after push() returns (the DMA must have read it by then)."""
    import torch
    c, m, sel = synth_case(33, 64, 900, 1000, 4)
    X = c[:, :3000].reshape(64, 1000, 3)[:, sel].astype(np.float64)
    mass = m[sel]

    def centred(x):
        return x - (mass[:, None] * x).sum(0) / mass.sum()

    # previous: frame f against frame f-1 (frame 0 against itself)
    want = np.array([oracle.rmsd_1vN(c[f:f + 1], sel, X[max(f - 1, 0)], mass=mass)[0] for f in range(64)])
    buf = torch.empty((1, c.shape[1]), dtype=torch.float32).pin_memory()
    with b200.Rmsd1vN(centred(X[0]), sel, mass, True, True) as h:
        for f in range(64):
            if f > 0:
                h.set_ref(centred(X[f - 1]))
            buf.numpy()[0] = c[f]
            h.push(buf.numpy())
            buf.numpy()[0] = 1e9              # the caller reuses its buffer at once
        r, rot, tr, best = h.flush()
    assert maxdiff(r, want) <= 1e-5
    check_argmin(best, want)
    assert np.allclose(np.einsum("fij,fkj->fik", rot.reshape(-1, 3, 3), rot.reshape(-1, 3, 3)), np.eye(3), atol=1e-9)
    # one frame at a time, synchronously (Action_Rmsd::DoAction): push, flush, push, flush
    with b200.Rmsd1vN(centred(X[5]), sel, mass, True, True) as h:
        for f in (0, 9, 33):
            h.push(np.ascontiguousarray(c[f:f + 1, :3000], np.float64))
            r1, rot1, tr1, _ = h.flush()
            w, wrot, wtr, _ = oracle.rmsd_1vN(c[f:f + 1], sel, X[5], mass=mass, want_rot=True)
            assert maxdiff(r1, w) <= 1e-5 and maxdiff(rot1, wrot) <= 1e-6 and maxdiff(tr1, wtr) <= 1e-9


# ---------------------------------------------------------------- frames x centroids
@pytest.mark.parametrize("K,use_mass", [(7, False), (33, True), (2, False)])
def test_frames_to_centroids_one_contraction(b200, oracle, K, use_mass):
    """K >= 3 fitted: one frames x centroids contraction on the tcgen05 engine (frames quantised once);
    K = 2: streaming passes.  Against K one-vs-many evaluations of the oracle."""
    c, m, sel = synth_case(777, 1500, 200, 230, 1)
    mass = m[sel] if use_mass else None
    w = np.ones(len(sel)) if mass is None else mass
    X = c[:, : 3 * 230].reshape(1500, 230, 3)[:, sel].astype(np.float64)
    cen = []
    for k in range(K):
        a = (X[37 * k] + X[37 * k + 1] + X[37 * k + 2]) / 3.0
        cen.append(a - (w[:, None] * a).sum(0) / w.sum())
    cen = np.array(cen)
    fidx = np.arange(1, 1500, 3, dtype=np.int32)[::-1].copy()
    for frame_idx in (None, fidx):
        dist, closest, cdist = b200.frames_to_centroids(c, sel, cen, mass=mass, frame_idx=frame_idx)
        assert b200.last_pair_engine()[0] == (2 if K >= 3 else 1)
        frames = c if frame_idx is None else c[frame_idx]
        want = np.stack([oracle.rmsd_1vN(frames, sel, cen[k], mass=mass) for k in range(K)], axis=1)
        assert maxdiff(dist, want) <= TOL and maxdiff(cdist, want.min(1)) <= TOL
        srt = np.sort(want, axis=1)
        clear = (srt[:, 1] - srt[:, 0]) > 2 * TOL
        assert np.array_equal(closest[clear], want.argmin(1)[clear])
        assert np.array_equal(closest, dist.argmin(1))


# ---------------------------------------------------------------- centroid building with fit
@pytest.mark.parametrize("fit,use_mass", [(True, False), (True, True), (False, False)])
def test_build_centroids(b200, oracle, fit, use_mass):
    """Metric_RMS::CalculateCentroid (src/Cluster/Metric_RMS.cpp:86-113): frames fitted in order to the running sum.
    The oracle's restatement is bit-identical to the reference's own Frame arithmetic (tests/test_oracle.py)."""
    c, m, sel = synth_case(4711, 400, 260, 300, 5)
    mass = m[sel] if use_mass else None
    rng = np.random.default_rng(1)
    lists = [np.array([7], np.int32), np.arange(0, 400, 4, dtype=np.int32), rng.permutation(400)[:150].astype(np.int32),
             np.zeros(0, np.int32), np.array([64, 63, 399, 5, 5], np.int32)]      # bit-identical neighbours, a repeated frame, an empty cluster
    got = b200.build_centroids(c, sel, lists, mass=mass, fit=fit)
    assert got.shape == (len(lists), len(sel), 3)
    for k, fr in enumerate(lists):
        if len(fr) == 0:
            assert np.all(got[k] == 0.0)
            continue
        want = oracle.build_centroid(c, sel, fr, mass=mass, fit=fit)
        assert maxdiff(got[k], want) <= 1e-8, (k, maxdiff(got[k], want))
    # COORDS resident on the device (a clustering run): same bits, no frame upload
    b200.coords_resident_begin(c, sel)
    try:
        b200.reset_stats()
        assert np.array_equal(b200.build_centroids(c, sel, lists, mass=mass, fit=fit), got)
        assert b200.get_stats()["h2d_bytes"] < 0.05 * c.nbytes
        if fit:
            d1 = b200.frames_to_centroids(c, sel, got[[1, 2, 4]].reshape(3, -1), mass=mass)[0]
    finally:
        b200.coords_resident_end()
    if fit:
        assert np.array_equal(b200.frames_to_centroids(c, sel, got[[1, 2, 4]].reshape(3, -1), mass=mass)[0], d1)
    # the centroids feed the frame-to-centroid distances: same nearest centroid as with the oracle's centroids
    if fit:
        cen = got[[0, 1, 2, 4]]
        _, closest, _ = b200.frames_to_centroids(c, sel, cen.reshape(4, -1), mass=mass)
        wantc = np.stack([oracle.rmsd_1vN(c, sel, oracle.build_centroid(c, sel, lists[k], mass=mass), mass=mass) for k in (0, 1, 2, 4)], axis=1)
        srt = np.sort(wantc, axis=1)
        clear = (srt[:, 1] - srt[:, 0]) > 2 * TOL
        assert np.array_equal(closest[clear], wantc.argmin(1)[clear])


# ---------------------------------------------------------------- several devices, one process (what CPPTRAJ_B200_NGPU drives)
def test_single_process_multi_device(oracle):
    import torch
    import cpptraj_b200 as b
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    nd = min(4, torch.cuda.device_count())
    c, m, sel = synth_case(55, 3000, 96, 100, 0)
    mass = m[sel]
    try:
        b.init(1)
        one = b.rms2d_tri(c, sel, mass=mass)
        one_eng = b.last_pair_engine()
        one_sieve = b.rms2d_tri(c, sel, frame_idx=np.arange(0, 3000, 3, dtype=np.int32))
        one_nofit = b.rms2d_tri(c[:700], sel, fit=False)
        one_full = b.rms2d_full(c[:900], sel, c[2000:2500], sel, mass_tgt=mass, mass_ref=mass)
        assert b.init(nd) == nd
        many = b.rms2d_tri(c, sel, mass=mass)
        assert b.last_pair_engine() == one_eng
        assert np.array_equal(many, one)                      # one grid for all shards: bit-identical
        assert np.array_equal(b.rms2d_tri(c, sel, frame_idx=np.arange(0, 3000, 3, dtype=np.int32)), one_sieve)
        assert np.array_equal(b.rms2d_tri(c[:700], sel, fit=False), one_nofit)
        assert np.array_equal(b.rms2d_full(c[:900], sel, c[2000:2500], sel, mass_tgt=mass, mass_ref=mass), one_full)
        # one-vs-many: chunks go round-robin to the devices, results come back in push order
        ref_raw = c[11].reshape(-1, 3)[sel].astype(np.float64)
        ref = ref_raw - (mass[:, None] * ref_raw).sum(0) / mass.sum()
        want = oracle.rmsd_1vN(c, sel, ref_raw, mass=mass)
        with b.Rmsd1vN(ref, sel, mass, True, True) as h:
            for a, e in ((0, 1), (1, 700), (700, 701), (701, 3000)):
                h.push(c[a:e])
            r, rot, tr, best = h.flush()
        assert maxdiff(r, want) <= 1e-5
        check_argmin(best, want)
        # frames x centroids over the devices
        X = c[:, :300].reshape(3000, 100, 3)[:, sel].astype(np.float64)
        cen = np.array([X[k * 100] - (mass[:, None] * X[k * 100]).sum(0) / mass.sum() for k in range(6)])
        big = np.concatenate([c, c[::-1]])                    # 6000 frames: enough to be split
        dist, closest, cdist = b.frames_to_centroids(big, sel, cen, mass=mass)
        wantd = np.stack([oracle.rmsd_1vN(big, sel, cen[k], mass=mass) for k in range(6)], axis=1)
        assert maxdiff(dist, wantd) <= TOL and np.array_equal(closest, dist.argmin(1))
        # rmsavgcorr: the window sizes are dealt to the devices; every device builds its own prefix sums
        win = np.arange(1, 601, dtype=np.int32)
        many_ac = b.rmsavgcorr(c[:600], sel, win, mass=mass)
        b.init(1)
        one_ac = b.rmsavgcorr(c[:600], sel, win, mass=mass)
        assert np.array_equal(many_ac[0], one_ac[0]) and np.array_equal(many_ac[1], one_ac[1])
    finally:
        b.shutdown()
