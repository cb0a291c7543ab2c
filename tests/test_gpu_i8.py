"""GPU tests of the tcgen05 int8 pair engine (cpptraj_b200/csrc/pair_i8.cuh), through the C ABI.

Layers checked separately so a failure points at one stage:
  1. the packed operand image (centre, sqrt(mass) scale, fixed-point rounding, balanced base-256 digits,
     UMMA core-matrix layout) against a numpy restatement;
  2. the raw integer covariances out of TMEM against an exact int64 computation from the decoded image
     (bit-exact: this is integer work);
  3. the RMSDs against the oracle (cpptraj's Frame::RMSD_CenteredRef restated, src/Frame.cpp:1137-1273),
     tolerance 1e-4 A absolute (BASELINE.json north_star).
"""
import numpy as np
import pytest

from helpers import synth_case, tri_to_square, TOL

pytestmark = pytest.mark.gpu


def maxdiff(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max()) if len(a) else 0.0


def decode_image(img, nframes, natoms):
    """Inverse of i8_quant_kernel's layout: returns int64 q[frame, atom, plane]."""
    nC = (natoms + 63) // 64
    q = np.zeros((nframes, natoms, 3), np.int64)
    b = img.view(np.int8)
    k = np.arange(natoms)
    c, kb = k // 64, k % 64
    koff = c * 8192 + (kb // 16) * 128 + kb % 16
    for f in range(nframes):
        g, il = divmod(f, 14)
        base = g * nC * 8192
        for p in range(3):
            for s in range(3):
                r = 9 * il + 3 * p + s
                d = b[base + koff + (r // 8) * 512 + (r % 8) * 16].astype(np.int64)
                q[f, :, p] += d << (8 * s)
    return q


def numpy_quant(crd, sel, mass, qs):
    X = crd[:, :].reshape(crd.shape[0], -1)[:, : 3 * (sel.max() + 1)].reshape(crd.shape[0], -1, 3)[:, sel].astype(np.float64)
    w = np.ones(len(sel)) if mass is None else np.asarray(mass, np.float64)
    c = (X * w[None, :, None]).sum(1) / w.sum()
    v = (X - c[:, None, :]) * np.sqrt(w)[None, :, None]
    return v, np.rint(v * 2.0 ** qs).astype(np.int64)


MODES = [2, 1]
MODE_IDS = ["cta_pair", "single_cta"]


@pytest.fixture(params=MODES, ids=MODE_IDS)
def i8(b200, request):
    """tcgen05 engine forced, in both MMA CTA-group modes (tcgen05.mma.cta_group::2 pairs / ::1)."""
    b200.set_pair_engine("i8")
    b200.set_i8_cta_group(request.param)
    yield b200
    b200.set_i8_cta_group(2)
    b200.set_pair_engine("auto")


@pytest.mark.parametrize("nf,na,use_mass", [(30, 70, False), (45, 130, True), (17, 64, True)])
def test_packed_image_and_G(b200, nf, na, use_mass):
    c, m, sel = synth_case(3 + nf, nf, na, na + 7, 5)
    mass = m[sel] if use_mass else None
    r = b200.debug_i8(c, sel, mass=mass)
    v, want_q = numpy_quant(c, sel, mass, r["qs"])
    assert np.abs(want_q).max() <= 8355711 and np.abs(want_q).max() * 2 > 8355711, "scale not tight"
    got_q = decode_image(r["image"][: r["image_bytes"]], nf, na)
    # rint of a double product: allow one unit where the FP64 centre differs in the last bit
    assert np.abs(got_q - want_q).max() <= 1
    assert (got_q != want_q).mean() < 1e-3
    G = (got_q.astype(object) ** 2).sum((1, 2)).astype(np.float64) * 2.0 ** (-2 * r["qs"])
    assert np.allclose(r["G"], G, rtol=1e-15, atol=0)
    # padding rows / atoms of the image are zero
    used = np.zeros(r["image_bytes"], bool)
    nC = (na + 63) // 64
    k = np.arange(na); koff = (k // 64) * 8192 + ((k % 64) // 16) * 128 + k % 16
    for f in range(nf):
        g, il = divmod(f, 14)
        for x in range(9):
            rr = 9 * il + x
            used[g * nC * 8192 + koff + (rr // 8) * 512 + (rr % 8) * 16] = True
    assert not r["image"][: r["image_bytes"]][~used].any()


@pytest.mark.parametrize("mode", MODES, ids=MODE_IDS)
@pytest.mark.parametrize("nf,na", [(20, 64), (30, 70), (61, 200), (90, 130), (130, 1000)])
def test_integer_covariance_is_exact(b200, nf, na, mode):
    c, m, sel = synth_case(11 + nf, nf, na)
    b200.set_i8_cta_group(mode)
    try:
        r = b200.debug_i8(c, sel, mass=m[sel])
    finally:
        b200.set_i8_cta_group(2)
    q = decode_image(r["image"][: r["image_bytes"]], nf, na)
    S = np.einsum("iap,jaq->ijpq", q, q).reshape(nf, nf, 9)    # int64, exact (|q| < 2^23, na small)
    iu = np.triu_indices(nf, 1)
    got = r["S"][iu]
    if na <= 256:    # every partial sum is an integer below 2^53: bit-exact
        assert np.array_equal(got, S[iu].astype(np.float64)), "tcgen05 int8 covariance differs from the exact integers"
    else:            # the digit recombination rounds to FP64 (2^-53 relative per add)
        want = S[iu].astype(np.float64)
        assert np.all(np.abs(got - want) <= 4.5e-16 * np.abs(S[iu]).max() + 4e-16 * np.abs(want))


@pytest.mark.parametrize("nf,na,ntot,extra", [
    (2, 3, 3, 0), (15, 12, 223, 0), (29, 64, 64, 0), (100, 65, 80, 3), (333, 1000, 1000, 0), (57, 1023, 1100, 6),
    (700, 200, 200, 0), (90, 1100, 1100, 0)])
def test_i8_fit_parity_shapes(i8, oracle, nf, na, ntot, extra):
    c, m, sel = synth_case(500 + nf, nf, na, ntot, extra)
    for mass in (None, m[sel]):
        got = i8.rms2d_tri(c, sel, mass=mass)
        assert i8.last_pair_engine()[0] == 2
        assert maxdiff(got, oracle.rms2d_tri(c, sel, mass=mass)) <= TOL


def test_i8_tz2_goldens(i8, tz2, saves, live):
    crd = tz2["crd"]
    sel = tz2["res"](3, 7)
    got = tri_to_square(i8.rms2d_tri(crd[:10], sel), 10)
    assert i8.last_pair_engine()[0] == 2
    assert np.abs(got - saves["rmsd"]).max() <= 5.1e-4          # 3-decimal golden of test/Test_2DRMS
    assert maxdiff(got[np.triu_indices(10, 1)], live["tz2_3to7_fit"]) <= TOL                      # reference binary, prec 14.8


def test_i8_duplicates_sieve_full_and_shards(i8, oracle):
    c, m, sel = synth_case(77, 200, 300, 320, 4)
    c[50] = c[10]
    tri = i8.rms2d_tri(c, sel)
    sq = tri_to_square(tri, 200)
    assert sq[10, 50] <= 1e-5 and sq[62, 63] <= 1e-5               # bit-exact duplicates
    # sieve (cluster pairwise cache: frameIdx = framesToCache)
    fidx = np.arange(3, 200, 3, dtype=np.int32)[::-1].copy()
    assert maxdiff(i8.rms2d_tri(c, sel, mass=m[sel], frame_idx=fidx), oracle.rms2d_tri(c, sel, mass=m[sel], frame_idx=fidx)) <= TOL
    # full matrix, different masks and masses (orientation pin: out[itgt*nRef + iref])
    sel2 = sel + 5
    got = i8.rms2d_full(c[:90], sel, c[90:], sel2, mass_tgt=m[sel], mass_ref=m[sel2])
    assert i8.last_pair_engine()[0] == 2
    want = oracle.rms2d_full(c[:90], sel, c[90:], sel2, massTgt=m[sel], massRef=m[sel2])
    assert maxdiff(got.reshape(-1), want.reshape(-1)) <= TOL
    # shards tile the triangle
    whole = np.zeros_like(tri)
    for r in range(3):
        part, first, n = i8.rms2d_tri_shard(c, sel, r, 3)
        whole[first:first + n] = part[first:first + n]
    assert np.array_equal(whole, tri)


def test_auto_engine_falls_back_for_huge_extents(b200, oracle):
    c, m, sel = synth_case(5, 40, 50)
    big = c.copy()
    big.reshape(40, -1, 3)[:, ::2, :] *= 40.0                       # atoms up to ~10^4 A from the centre
    b200.set_pair_engine("auto")
    got = b200.rms2d_tri(big, sel)
    assert b200.last_pair_engine()[0] == 1
    want = oracle.rms2d_tri(big, sel)
    assert maxdiff(got, want) <= TOL * max(1.0, float(want.max()) / 100.0)
    b200.set_pair_engine("i8")
    with pytest.raises(b200.B200Error):
        b200.rms2d_tri(big, sel)
    with pytest.raises(b200.B200Error):
        b200.rms2d_tri(c, sel, fit=False)                           # nofit is FP64-only
    b200.set_pair_engine("auto")
    assert maxdiff(b200.rms2d_tri(c, sel, fit=False), oracle.rms2d_tri(c, sel, fit=False)) <= TOL


def test_i8_matches_fp64_engine_on_config2_prefix(b200, oracle):
    """Both engines on a 2,000-frame prefix of BASELINE configs[1] (1,000 atoms): they must agree with
    each other to well inside the contract, and with the oracle on sampled rows."""
    from cpptraj_b200.synth import make_trajectory
    crd, _ = make_trajectory(20261017, 2000, 1000)
    sel = np.arange(1000, dtype=np.int32)
    b200.set_pair_engine("fp64")
    a = b200.rms2d_tri(crd, sel)
    b200.set_pair_engine("i8")
    b = b200.rms2d_tri(crd, sel)
    assert b200.last_pair_engine()[0] == 2
    b200.set_pair_engine("auto")
    assert maxdiff(a, b) <= 5e-5
    sub = np.r_[0:40, 980:1020, 1960:2000]
    want = oracle.rms2d_tri(crd[sub], sel)
    got = tri_to_square(b, 2000)[np.ix_(sub, sub)][np.triu_indices(len(sub), 1)]
    assert maxdiff(got, want) <= TOL


def _pinned(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


@pytest.mark.parametrize("nf,na", [(1500, 64), (3001, 130)])
def test_pipelined_host_path_matches_two_pass(b200, oracle, nf, na, monkeypatch):
    """Pinned COORDS take the pipelined host path (reverse-order chunk upload overlapped with compute and download, scale
    fixed from the first chunk): it must give what the two-pass path gives (same rounding grid or one bit coarser), and
    the oracle's values; shards included."""
    c, m, sel = synth_case(900 + nf, nf, na, na + 7, 2)
    keep, cp = _pinned(c)
    b200.set_pair_engine("auto")
    monkeypatch.setenv("B200_HOST_PIPELINE", "0")
    two_pass = b200.rms2d_tri(cp, sel, mass=m[sel])
    assert b200.last_pair_engine()[0] == 2
    monkeypatch.setenv("B200_HOST_PIPELINE", "1")
    piped = b200.rms2d_tri(cp, sel, mass=m[sel])
    assert b200.last_pair_engine()[0] == 2
    assert maxdiff(piped, two_pass) <= 5e-5
    sub = np.r_[0:30, nf // 2:nf // 2 + 30, nf - 30:nf]
    want = oracle.rms2d_tri(c[sub], sel, mass=m[sel])
    got = tri_to_square(piped, nf)[np.ix_(sub, sub)][np.triu_indices(len(sub), 1)]
    assert maxdiff(got, want) <= TOL
    whole = np.zeros_like(piped)
    for r in range(3):
        part, first, n = b200.rms2d_tri_shard(cp, sel, r, 3, mass=m[sel])
        whole[first:first + n] = part[first:first + n]
    assert np.array_equal(whole, piped)


def test_pipelined_host_path_headroom_fallback(b200, oracle, monkeypatch):
    """The last frames (uploaded first) are compact, earlier ones 3x more extended: the scale guessed from the first chunk
    does not hold, the call must notice and recompute on the two-pass path."""
    c, m, sel = synth_case(31, 2200, 40)
    X = c.reshape(2200, -1, 3)
    X[:600] = (X[:600] - X[:600].mean(1, keepdims=True)) * 3.0 + X[:600].mean(1, keepdims=True)
    keep, cp = _pinned(X.reshape(2200, -1))
    monkeypatch.setenv("B200_HOST_PIPELINE", "1")
    got = b200.rms2d_tri(cp, sel)
    sub = np.r_[0:25, 590:615, 2175:2200]
    want = oracle.rms2d_tri(cp[sub], sel)
    assert maxdiff(tri_to_square(got, 2200)[np.ix_(sub, sub)][np.triu_indices(len(sub), 1)], want) <= TOL


def test_i8_config4_like_mass_weighted(b200, oracle):
    """BASELINE configs[3] in small: Metric_RMS, mass-weighted, 2,000 atoms (32 K-chunks per tile), sieved frame list;
    the tcgen05 engine must be the one that runs and must agree with the oracle."""
    from cpptraj_b200.synth import make_trajectory, masses
    crd, _ = make_trajectory(20261019, 400, 2000)
    m = masses(2000)
    sel = np.arange(2000, dtype=np.int32)
    b200.set_pair_engine("auto")
    fidx = np.arange(0, 400, 3, dtype=np.int32)
    got = b200.rms2d_tri(crd, sel, mass=m, frame_idx=fidx)
    assert b200.last_pair_engine()[0] == 2, "mass-weighted 2,000-atom selection should be eligible for the tcgen05 engine"
    want = oracle.rms2d_tri(crd, sel, mass=m, frame_idx=fidx)
    assert maxdiff(got, want) <= TOL
