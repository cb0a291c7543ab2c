"""rmsavgcorr (Analysis_RmsAvgCorr::Analyze, src/Analysis_RmsAvgCorr.cpp:119-316).

CPU part: the restatement (oracle/rmsd_oracle.c: orc_rmsavgcorr) against the reference's golden vectors
(test/Test_RmsAvgCorr/*.save, numbers in tests/golden/ref_saves.npz) and against the reference's own Frame arithmetic
(oracle/_ref: ref_rmsavgcorr, bit-equal).  GPU part (-m gpu): b200_rmsavgcorr through the C ABI against the
restatement, tolerance 1e-4 A on the mean and the standard deviation of every window size (north_star's tolerance;
measured differences are ~1e-9: FP64 prefix sums instead of a running sum)."""
import numpy as np
import pytest

from helpers import synth_case

TOL = 1e-4


def tz2_ca(tz2):
    return np.nonzero((tz2["names"] == "CA") & (tz2["resnum"] >= 2) & (tz2["resnum"] <= 12))[0].astype(np.int32)


def test_golden_first_mode(oracle, tz2, saves):
    # test/Test_RmsAvgCorr/RunTest.sh:27-34: strip !(:2-12@CA); rmsavgcorr ... first
    avg, sd = oracle.rmsavgcorr(tz2["crd"], tz2_ca(tz2), np.arange(1, 101))
    assert np.abs(np.round(avg, 4) - saves["rmsavgcorr_first"][:, 0]).max() < 1.01e-4
    assert np.abs(np.round(sd, 4) - saves["rmsavgcorr_first"][:, 1]).max() < 1.01e-4


def test_golden_fixed_reference(oracle, tz2, saves):
    # RunTest.sh:11-24: reference avg.CA.rst7, centred without mass (Analysis_RmsAvgCorr.cpp:86-90); offset 10
    ref = saves["avg_ca_rst7"] - saves["avg_ca_rst7"].mean(0)
    avg, sd = oracle.rmsavgcorr(tz2["crd"], tz2_ca(tz2), np.arange(1, 101), ref_sel_xyz=ref)
    assert np.abs(np.round(avg, 4) - saves["rmsavgcorr_ref"][:, 0]).max() < 1.01e-4
    assert np.abs(np.round(sd, 4) - saves["rmsavgcorr_ref"][:, 1]).max() < 1.01e-4
    avg, sd = oracle.rmsavgcorr(tz2["crd"], tz2_ca(tz2), np.arange(1, 101, 10), ref_sel_xyz=ref)
    assert np.abs(np.round(avg, 4) - saves["rmsavgcorr_ref10"][:, 0]).max() < 1.01e-4
    assert np.abs(np.round(sd, 4) - saves["rmsavgcorr_ref10"][:, 1]).max() < 1.01e-4


def test_restatement_matches_reference_frame_arithmetic(oracle, reference):
    c, m, sel = synth_case(3, 60, 25, ntot=40, extra=3)
    win = np.array([1, 2, 3, 7, 30, 59, 60], np.int32)
    fixed = c[5].reshape(-1)[:120].reshape(-1, 3)[sel].astype(np.float64)
    fixed -= fixed.mean(0)
    for mass in (None, m[sel]):
        for ref in (None, fixed):
            a = oracle.rmsavgcorr(c, sel, win, mass, ref)
            b = reference.rmsavgcorr(c, sel, win, mass, ref, natom_total=40)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_goldens(b200, tz2, saves):
    sel = tz2_ca(tz2)
    avg, sd = b200.rmsavgcorr(tz2["crd"], sel, np.arange(1, 101))
    assert np.abs(np.round(avg, 4) - saves["rmsavgcorr_first"][:, 0]).max() < 1.01e-4
    assert np.abs(np.round(sd, 4) - saves["rmsavgcorr_first"][:, 1]).max() < 1.01e-4
    ref = saves["avg_ca_rst7"] - saves["avg_ca_rst7"].mean(0)
    avg, sd = b200.rmsavgcorr(tz2["crd"], sel, np.arange(1, 101), ref_selected=ref)
    assert np.abs(np.round(avg, 4) - saves["rmsavgcorr_ref"][:, 0]).max() < 1.01e-4
    assert np.abs(np.round(sd, 4) - saves["rmsavgcorr_ref"][:, 1]).max() < 1.01e-4


@pytest.mark.gpu
@pytest.mark.parametrize("nf,na,ntot,extra", [(2, 5, 5, 0), (61, 25, 40, 3), (300, 97, 120, 5), (129, 256, 256, 0)])
def test_gpu_matches_restatement(b200, oracle, nf, na, ntot, extra):
    c, m, sel = synth_case(11 + nf, nf, na, ntot=ntot, extra=extra)
    c = c + np.float32(37.5)                                   # away from the origin: the centre is removed algebraically
    win = np.arange(1, nf + 1, dtype=np.int32)
    fixed = c[nf // 2].reshape(-1)[:3 * ntot].reshape(-1, 3)[sel].astype(np.float64)
    fixed -= np.average(fixed, axis=0, weights=m[sel])        # (reference centred with ITS weights, target unweighted below)
    for mass in (None, m[sel]):
        for ref in (None, fixed):
            want = oracle.rmsavgcorr(c, sel, win, mass, ref)
            got = b200.rmsavgcorr(c, sel, win, mass, ref)
            assert np.abs(got[0] - want[0]).max() <= TOL, np.abs(got[0] - want[0]).max()
            assert np.abs(got[1] - want[1]).max() <= TOL, np.abs(got[1] - want[1]).max()
    # a subset of window sizes in arbitrary order, and zero total mass (src/Frame.cpp:1160-1163: RMSD -1)
    if nf > 10:
        sub = np.array([7, 1, nf, 2, nf - 1], np.int32)
        want = oracle.rmsavgcorr(c, sel, sub, m[sel], None)
        got = b200.rmsavgcorr(c, sel, sub, m[sel], None)
        assert np.abs(got[0] - want[0]).max() <= TOL and np.abs(got[1] - want[1]).max() <= TOL
        got = b200.rmsavgcorr(c, sel, sub, np.zeros(len(sel)), None)
        want = oracle.rmsavgcorr(c, sel, sub, np.zeros(len(sel)), None)
        assert np.array_equal(got[0], want[0]) and np.all(got[0] == -1.0)


@pytest.mark.gpu
def test_gpu_argument_errors(b200):
    c, m, sel = synth_case(1, 20, 10)
    for bad in ([0], [21], [-3]):
        with pytest.raises(b200.B200Error):
            b200.rmsavgcorr(c, sel, np.array(bad, np.int32))
