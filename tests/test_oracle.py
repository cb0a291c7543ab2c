"""CPU-only: pin the oracle against the reference's golden vectors and its real code."""
import numpy as np
import pytest

from helpers import tri_to_square, synth_case


def test_tri_index_matches_row_major_upper():
    from oracle.pyoracle import tri_index, tri_size
    n = 7
    k = 0
    for i in range(n):
        for j in range(i + 1, n):
            assert tri_index(n, i, j) == k
            k += 1
    assert k == tri_size(n)


def test_golden_2drms_fit(oracle, tz2, saves):
    # test/Test_2DRMS/RunTest.sh:15-22  "2drms crd1 :3-7" on frames 1-10
    out = oracle.rms2d_tri(tz2["crd"][:10], tz2["res"](3, 7))
    assert np.array_equal(np.round(tri_to_square(out, 10), 3), saves["rmsd"])


def test_golden_2drms_mass(oracle, tz2, saves):
    sel = tz2["res"](3, 7)
    out = oracle.rms2d_tri(tz2["crd"][:10], sel, mass=tz2["mass"][sel])
    assert np.array_equal(np.round(tri_to_square(out, 10), 3), saves["rmsd_mass"])


def test_golden_2drms_refmask_full(oracle, tz2, saves):
    # test 5: ":2 :11" => non-symmetric full matrix; pins the [ntgt][nref] orientation
    out = oracle.rms2d_full(tz2["crd"][:10], tz2["res"](2, 2), tz2["crd"][:10], tz2["res"](11, 11))
    assert np.abs(np.round(out.astype(np.float64), 3) - saves["trp"]).max() < 1e-6
    assert np.abs(np.round(out.T.astype(np.float64), 3) - saves["trp"]).max() > 0.01


def test_golden_rms_nomod(oracle, tz2, saves):
    # test/Test_RMSD/RunTest.sh:64-76  "rms First_CA :2-12@CA nomod"
    ca = np.nonzero((tz2["names"] == "CA") & (tz2["resnum"] >= 2) & (tz2["resnum"] <= 12))[0].astype(np.int32)
    ref = tz2["crd"][0].reshape(-1, 3)[ca].astype(np.float64)
    r = oracle.rmsd_1vN(tz2["crd"], ca, ref)
    assert np.array_equal(np.round(r, 4), saves["nomod"])


def test_golden_rms_previous(oracle, tz2, saves):
    # test/Test_RMSD/RunTest.sh:78-91 "rms ToPrevious :2-12@CA previous": frame f fitted to frame f-1 as it is
    # AFTER its own fit+move; RMSD is invariant to that rigid move, so frame f vs raw frame f-1 is the same number.
    ca = np.nonzero((tz2["names"] == "CA") & (tz2["resnum"] >= 2) & (tz2["resnum"] <= 12))[0].astype(np.int32)
    crd = tz2["crd"]
    vals = [0.0]
    for f in range(1, crd.shape[0]):
        ref = crd[f - 1].reshape(-1, 3)[ca].astype(np.float64)
        vals.append(oracle.rmsd_1vN(crd[f:f + 1], ca, ref)[0])
    assert np.abs(np.round(np.array(vals), 4) - saves["previous"]).max() <= 1.0001e-4


def test_oracle_vs_live_reference_fixtures(oracle, tz2, live):
    """ref_live.npz was produced by the reference's own classes (tools/make_golden.py)."""
    crd, mass = tz2["crd"], tz2["mass"]
    s37 = tz2["res"](3, 7)
    assert np.array_equal(oracle.rms2d_tri(crd[:10], s37), live["tz2_3to7_fit"])
    assert np.array_equal(oracle.rms2d_tri(crd[:10], s37, mass=mass[s37]), live["tz2_3to7_mass"])
    assert np.array_equal(oracle.rms2d_full(crd[:10], tz2["res"](2, 2), crd[:10], tz2["res"](11, 11)), live["tz2_trp_full"])
    allca = np.nonzero(tz2["names"] == "CA")[0].astype(np.int32)
    assert np.array_equal(oracle.rms2d_tri(crd, allca), live["tz2_allca_fit"])
    assert np.array_equal(oracle.cluster_tri(crd, allca, mass=mass[allca]), live["tz2_allca_cluster_mass"])
    for tag, seed, nf, na, ntot in (("s1", 11, 48, 100, 100), ("s2", 12, 40, 257, 300), ("s3", 13, 33, 7, 20)):
        c, m, sel = synth_case(seed, nf, na, ntot)
        assert np.array_equal(sel, live[tag + "_sel"])
        assert np.array_equal(oracle.rms2d_tri(c, sel), live[tag + "_fit"])
        assert np.array_equal(oracle.rms2d_tri(c, sel, mass=m[sel]), live[tag + "_mass"])
        assert np.array_equal(oracle.rms2d_tri(c, sel, fit=False), live[tag + "_nofit"])
        r, rot, tr, rt = oracle.rmsd_1vN(c, sel, c[0].reshape(-1, 3)[sel].astype(np.float64), mass=m[sel], want_rot=True)
        assert np.allclose(r, live[tag + "_1vN_rms"], atol=1e-12, rtol=0)
        assert np.allclose(rot, live[tag + "_1vN_rot"], atol=1e-10, rtol=0)
        assert np.allclose(tr, live[tag + "_1vN_tr"], atol=1e-12, rtol=0)


def test_oracle_vs_compiled_reference(oracle, reference):
    """Where oracle/_ref exists, the restatement must equal the reference's own code."""
    for seed, nf, na, ntot, extra in ((1, 40, 64, 64, 0), (2, 25, 30, 90, 0), (3, 20, 50, 50, 3 * 50 + 9)):
        c, m, sel = synth_case(seed, nf, na, ntot, extra)
        assert np.array_equal(oracle.rms2d_tri(c, sel), reference.rms2d_tri(c, sel, natom_total=ntot))
        assert np.array_equal(oracle.rms2d_tri(c, sel, mass=m[sel]), reference.rms2d_tri(c, sel, mass=m[sel], natom_total=ntot))
        assert np.array_equal(oracle.rms2d_tri(c, sel, fit=False), reference.rms2d_tri(c, sel, fit=False, natom_total=ntot))
        assert np.array_equal(oracle.cluster_tri(c, sel, mass=m[sel]), reference.cluster_tri(c, sel, mass=m[sel], natom_total=ntot))
        fidx = np.arange(nf - 1, -1, -3, dtype=np.int32)
        assert np.array_equal(oracle.rms2d_tri(c, sel, frame_idx=fidx), reference.rms2d_tri(c, sel, frame_idx=fidx, natom_total=ntot))
        sel2 = np.roll(sel, 1)
        a = oracle.rms2d_full(c[:7], sel, c, sel2, massTgt=m[sel], massRef=m[sel2])
        b = reference.rms2d_full(c[:7], sel, c, sel2, massTgt=m[sel], massRef=m[sel2], natomT=ntot, natomR=ntot)
        assert np.array_equal(a, b)


def test_cluster_equals_rms2d_within_rounding(oracle):
    """Metric_RMS::FrameDist (ref re-centred per pair, roles swapped) vs rms2d: same number to ~1e-6."""
    c, m, sel = synth_case(5, 30, 80)
    a = oracle.rms2d_tri(c, sel, mass=m[sel]).astype(np.float64)
    b = oracle.cluster_tri(c, sel, mass=m[sel]).astype(np.float64)
    assert np.abs(a - b).max() < 2e-6


def test_edge_cases(oracle):
    c, m, sel = synth_case(7, 5, 4)
    assert oracle.rms2d_tri(c[:1], sel).size == 0           # one frame: empty triangle
    out = oracle.rms2d_tri(np.vstack([c[:1], c[:1]]), sel)  # duplicate: ~0
    assert out.shape == (1,) and out[0] < 1e-5
    z = oracle.rms2d_tri(c, sel, mass=np.zeros(4))          # total mass < SMALL: reference returns -1
    assert np.all(z == -1.0)


def test_centroid_builder_vs_compiled_reference(oracle, reference):
    """Metric_RMS::CalculateCentroid restated (oracle/rmsd_oracle.c: orc_build_centroid) against the
    reference's own Frame::RMSD_CenteredRef / Rotate / += / Divide (oracle/ref_driver.cpp: ref_build_centroid)."""
    from helpers import synth_case
    c, m, sel = synth_case(9, 60, 40, 50, 2)
    for frames in (np.array([3, 17, 4, 55, 20, 21, 0], np.int32), np.array([8], np.int32), np.arange(60, dtype=np.int32)):
        for mass in (None, m[sel]):
            for fit in (True, False):
                a = oracle.build_centroid(c, sel, frames, mass, fit)
                b = reference.build_centroid(c, sel, frames, mass, fit, natom_total=50)
                assert np.array_equal(a, b)
