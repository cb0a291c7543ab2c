"""End to end through cpptraj itself: the reference's own Test_2DRMS decks run by a cpptraj binary whose rms2d pair
loops were replaced by the B200 path (cpptraj_host/ glue + reference.patch, built by tools/build_cpptraj_b200.sh --build
into oracle/_ref/cpptraj_b200/, which travels to the GPU box).  Outputs are compared with the reference's golden
`*.save` files (test/Test_2DRMS/RunTest.sh tests 1, 2, 5, 6; tz2.crd instead of tz2.nc: this build has no NetCDF).
"""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE = os.path.join(ROOT, "oracle", "_ref", "cpptraj_b200")
BIN = os.path.join(STAGE, "cpptraj.B200")


def table(path):
    rows = [l.split() for l in open(path) if l.strip() and not l.lstrip().startswith("#")]
    return np.array([[float(x) for x in r] for r in rows])


DECKS = [
    ("fit", "2drms crd1 :3-7 rmsout OUT", "rmsd.dat.save"),
    ("mass", "2drms crd1 :3-7 rmsout OUT mass", "rmsd.mass.dat.save"),
    ("refmask_full_matrix", "2drms :2 :11 out OUT", "trp.dat.save"),
    ("nofit", "rms first :2-12@CA\n2drms crd1 :2 nofit out OUT", "nofit.dat.save"),
    ("qrmsd_same_quantity", "2drms crd1 :3-7 rmsout OUT qrmsd", "rmsd.dat.save"),
]


@pytest.mark.parametrize("name,cmd,save", DECKS, ids=[d[0] for d in DECKS])
def test_cpptraj_2drms_decks_on_b200(tmp_path, name, cmd, save):
    if not os.path.exists(BIN):
        pytest.skip("cpptraj.B200 not staged (tools/build_cpptraj_b200.sh --build needs the reference tree)")
    out = tmp_path / "out.dat"
    deck = "noprogress\nparm %s\ntrajin %s 1 10\n%s\n" % (os.path.join(STAGE, "tz2.parm7"), os.path.join(STAGE, "tz2.crd"),
                                                          cmd.replace("OUT", str(out)))
    (tmp_path / "rms.in").write_text(deck)
    r = subprocess.run([BIN, "-i", str(tmp_path / "rms.in")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-3000:]
    assert "B200 RMSD path" in r.stdout, "the B200 branch did not run:\n" + r.stdout[-2000:]
    got, want = table(out), table(os.path.join(STAGE, save))
    assert got.shape == want.shape
    # goldens are printed with 3 decimals from NetCDF (float) coordinates; tz2.crd carries 3-decimal ASCII coordinates
    assert np.abs(got - want).max() <= 1.1e-3, np.abs(got - want).max()


def test_cpptraj_cluster_pairwise_cache_on_b200(tmp_path):
    """test/Test_Cluster/RunTest.sh, first deck: hierarchical agglomerative clustering whose in-memory pairwise cache
    (MetricArray::calcFrameDistances) is filled by the B200 path; cluster number vs time must match the golden exactly."""
    if not os.path.exists(BIN):
        pytest.skip("cpptraj.B200 not staged")
    deck = ("noprogress\nparm %s\ntrajin %s\n"
            "cluster C1 :2-10 clusters 3 epsilon 4.0 out cnumvtime.dat summary avg.summary.dat nofit\n"
            "cluster crd1 :2-10 clusters 3 epsilon 4.0 summary summary.dat complete nofit\n"
            % (os.path.join(STAGE, "tz2.parm7"), os.path.join(STAGE, "tz2.crd")))
    (tmp_path / "cluster.in").write_text(deck)
    r = subprocess.run([BIN, "-i", str(tmp_path / "cluster.in")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-3000:]
    assert "B200 RMSD path" in r.stdout, "the B200 branch did not run:\n" + r.stdout[-2000:]
    got, want = table(tmp_path / "cnumvtime.dat"), table(os.path.join(STAGE, "cnumvtime.dat.save"))
    assert np.array_equal(got, want)
    # complete-linkage summary: numeric columns (cluster, frames, fraction, avg dist, stdev, centroid frame, avg c-dist)
    gs = [l.split() for l in open(tmp_path / "summary.dat") if not l.startswith("#")]
    ws = [l.split() for l in open(os.path.join(STAGE, "summary.dat.save")) if not l.startswith("#")]
    assert len(gs) == len(ws)
    for g, w in zip(gs, ws):
        assert g[0] == w[0] and g[1] == w[1] and g[5] == w[5]
        assert abs(float(g[3]) - float(w[3])) <= 2e-3 and abs(float(g[6]) - float(w[6])) <= 2e-3


def test_cpptraj_rms_nomod_on_b200(tmp_path):
    """test/Test_RMSD/RunTest.sh "RMS nomod" against the reference's own NoMod.dat.save: the rmsd action, every frame
    fitted on the device (Action_Rmsd::B200_Fit) and added to the data set by frame number as on the CPU."""
    if not os.path.exists(BIN) or not os.path.exists(os.path.join(STAGE, "NoMod.dat.save")):
        pytest.skip("cpptraj.B200 (with the Action_Rmsd branch) not staged")
    deck = "parm %s\ntrajin %s\nrms First_CA :2-12@CA out NoMod.dat nomod\n" % (
        os.path.join(STAGE, "tz2.parm7"), os.path.join(STAGE, "tz2.crd"))
    (tmp_path / "rms.in").write_text(deck)
    r = subprocess.run([BIN, "-i", str(tmp_path / "rms.in")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-3000:]
    assert "B200 RMSD path" in r.stdout, "the B200 branch did not run:\n" + r.stdout[-2000:]
    got, want = table(tmp_path / "NoMod.dat"), table(os.path.join(STAGE, "NoMod.dat.save"))
    assert got.shape == want.shape == (101, 2)
    # golden printed with 4 decimals from NetCDF coordinates; tz2.crd carries 3-decimal ASCII coordinates
    assert np.abs(got - want).max() <= 2.1e-4, np.abs(got - want).max()


def test_cpptraj_cluster_sieve_restore_on_b200(tmp_path):
    """SURVEY 8(f) rank 1 through cpptraj itself: `cluster ... sieve 5` clusters every 5th frame (pairwise cache on the B200
    path) and restores the sieved-out frames by their nearest centroid (List::AddFramesByCentroid, src/Cluster/List.cpp:160-207),
    here computed by b200_rmsd_frames_to_centroids.  Golden: the UNMODIFIED reference on the CPU, same deck
    (tools/make_golden_cluster_sieve.sh -> tests/golden/cluster_sieve5.*)."""
    gold = os.path.join(ROOT, "tests", "golden", "cluster_sieve5.out")
    if not os.path.exists(BIN) or not os.path.exists(gold):
        pytest.skip("cpptraj.B200 not staged or golden missing")
    deck = ("noprogress\nparm %s\ntrajin %s\n"
            "cluster crd1 @CA clusters 5 rms out sieve5.out summary sieve5.summary.dat sieve 5 bestrep cumulative includesieveincalc\n"
            % (os.path.join(STAGE, "tz2.parm7"), os.path.join(STAGE, "tz2.crd")))
    (tmp_path / "cluster.in").write_text(deck)
    r = subprocess.run([BIN, "-i", str(tmp_path / "cluster.in")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=300, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-3000:]
    assert "B200 RMSD path" in r.stdout, "the B200 branch did not run:\n" + r.stdout[-2000:]
    got, want = table(tmp_path / "sieve5.out"), table(gold)
    assert np.array_equal(got, want), "cluster number vs time differs from the CPU reference"
    gs = [l.split() for l in open(tmp_path / "sieve5.summary.dat") if not l.startswith("#")]
    ws = [l.split() for l in open(os.path.join(ROOT, "tests", "golden", "cluster_sieve5.summary.dat")) if not l.startswith("#")]
    assert len(gs) == len(ws)
    for g, w in zip(gs, ws):
        assert g[0] == w[0] and g[1] == w[1] and g[5] == w[5]
        assert abs(float(g[3]) - float(w[3])) <= 2e-3 and abs(float(g[6]) - float(w[6])) <= 2e-3


# ---------------------------------------------------------------- decks against the unmodified reference
import sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from cpptraj_decks import DECKS as REF_DECKS   # noqa: E402


def _numbers(path):
    """Every token of a text file: floats where it parses, the string otherwise (comment lines included)."""
    out = []
    for line in open(path):
        for tok in line.split():
            try:
                out.append(float(tok))
            except ValueError:
                out.append(tok)
    return out


def _amber_crd(path):
    """Amber ASCII trajectory: title line, then 8.3f fields."""
    vals = []
    for line in open(path).readlines()[1:]:
        vals.extend(float(line[i:i + 8]) for i in range(0, len(line.rstrip("\n")), 8))
    return np.array(vals)


@pytest.mark.parametrize("name", sorted(REF_DECKS))
def test_cpptraj_deck_matches_unmodified_reference(tmp_path, name):
    """The reference's Test_RMSD decks 1-4 (first / mass / savevectors / reftraj; norotate / rotate / savematrices /
    outtraj; nomod; previous -- test/Test_RMSD/RunTest.sh:14-91), `crdaction ... rms` with coordinate modification, an
    action that reads the RMSD set while it is being filled, `rms2d ... corr`, k-means and hierarchical clustering with the
    epsilon sieve restore: cpptraj.B200 (every RMSD, rotation, translation, centroid and frame-to-centroid distance from the
    B200 path) against the UNMODIFIED reference run on the CPU (tools/make_golden_cpptraj.py -> tests/golden/cpptraj/)."""
    gold = os.path.join(ROOT, "tests", "golden", "cpptraj", name)
    if not os.path.exists(BIN) or not os.path.isdir(gold):
        pytest.skip("cpptraj.B200 not staged or goldens missing")
    text, outs = REF_DECKS[name]
    if "tz2.truncoct" in text and not os.path.exists(os.path.join(STAGE, "tz2.truncoct.crd")):
        pytest.skip("tz2.truncoct.* not staged")
    (tmp_path / "in").write_text(text.replace("{D}", STAGE))
    r = subprocess.run([BIN, "-i", "in"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0 and "Error" not in r.stdout, r.stdout[-3000:]
    assert "B200 RMSD path" in r.stdout, "the B200 branch did not run:\n" + r.stdout[-2000:]
    if name == "rmsavgcorr":
        assert r.stdout.count("window sizes over 101 frames on the device") == 4, "not on the device:\n" + r.stdout[-3000:]
    if name == "cluster_cmatrix_roundtrip":
        assert r.stdout.count("initial clusters on the device") == 2, "the device merge loop did not run:\n" + r.stdout[-3000:]
    if name == "cluster_hier_linkages":
        assert r.stdout.count("initial clusters on the device") == 4, "the device merge loop did not run:\n" + r.stdout[-3000:]
    for fname, kind in outs:
        got_p, want_p = str(tmp_path / fname), os.path.join(gold, fname)
        if kind == "table":
            got, want = table(got_p), table(want_p)
            assert got.shape == want.shape, (fname, got.shape, want.shape)
            # values are printed with 4 decimals: the contract (1e-4 A) plus one unit of the last printed digit
            assert np.abs(got - want).max() <= 2.01e-4, (fname, np.abs(got - want).max())
        elif kind == "cmatrix":
            # Cmatrix_Binary version 2 (src/Cluster/Cmatrix_Binary.cpp:12-21,150-188): magic, 3 x 8-byte header, float triangle
            g, w = open(got_p, "rb").read(), open(want_p, "rb").read()
            assert len(g) == len(w) and g[:28] == w[:28], fname
            gf, wf = np.frombuffer(g[28:], np.float32), np.frombuffer(w[28:], np.float32)
            assert np.abs(gf.astype(np.float64) - wf).max() <= 1e-4, (fname, np.abs(gf.astype(np.float64) - wf).max())
        elif kind == "crd":
            got, want = _amber_crd(got_p), _amber_crd(want_p)
            assert got.shape == want.shape, (fname, got.shape, want.shape)
            assert np.abs(got - want).max() <= 1.01e-3, (fname, np.abs(got - want).max())   # 8.3f fields: one unit of the last digit
        else:
            got, want = _numbers(got_p), _numbers(want_p)
            assert len(got) == len(want), fname
            for g, w in zip(got, want):
                if isinstance(w, float):
                    assert isinstance(g, float) and abs(g - w) <= 2e-3 + 1e-6 * abs(w), (fname, g, w)
                else:
                    assert g == w, (fname, g, w)
