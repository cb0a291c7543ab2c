#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's own test data.

Run in the build container only (needs /root/reference; never at test time):
    python tools/make_golden.py

Writes
  tests/golden/tz2.npz        -- tz2.nc coordinates (float32, 101 x 223 x 3), atom
                                 names, residue number per atom, masses (tz2.parm7)
  tests/golden/ref_saves.npz  -- the numbers inside the reference's golden files
                                 test/Test_2DRMS/{rmsd,rmsd.mass,trp,nofit}.dat.save and
                                 test/Test_RMSD/{NoMod.dat,rmatrices.dat}.save
  tests/golden/ref_live.npz   -- full-precision outputs of the reference's own
                                 Frame/Matrix_3x3 code (oracle/_ref, build_ref.sh) on
                                 the tz2 decks and on seeded synthetic inputs, so the
                                 1e-4 A contract can be checked where /root/reference
                                 and oracle/_ref are both absent.
Only numbers are extracted; no reference source is copied.
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("CPPTRAJ_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def read_parm7(path):
    flags = {}
    cur, fmt = None, None
    with open(path) as fh:
        for line in fh:
            if line.startswith("%FLAG"):
                cur = line.split()[1]
                flags[cur] = []
            elif line.startswith("%FORMAT"):
                fmt = line.strip()[8:-1]
                flags[cur] = [fmt]
            elif line.startswith("%"):
                continue
            elif cur is not None:
                flags[cur].append(line.rstrip("\n"))

    def parse(name):
        fmt = flags[name][0]
        body = flags[name][1:]
        import re
        m = re.match(r"(\d+)([aIE])(\d+)", fmt)
        width = int(m.group(3))
        kind = m.group(2)
        vals = []
        for ln in body:
            for i in range(0, len(ln), width):
                tok = ln[i:i + width]
                if tok.strip() == "" and kind != "a":
                    continue
                if kind == "a":
                    if tok == "":
                        continue
                    vals.append(tok.strip())
                elif kind == "I":
                    vals.append(int(tok))
                else:
                    vals.append(float(tok))
        return vals

    ptr = parse("POINTERS")
    natom, nres = ptr[0], ptr[11]
    names = parse("ATOM_NAME")[:natom]
    mass = np.array(parse("MASS")[:natom], np.float64)
    rptr = parse("RESIDUE_POINTER")[:nres]
    resnum = np.zeros(natom, np.int32)
    for r, start in enumerate(rptr):
        stop = rptr[r + 1] - 1 if r + 1 < nres else natom
        resnum[start - 1:stop] = r + 1
    return names, resnum, mass


def read_save_matrix(path):
    rows = []
    with open(path) as fh:
        for ln in fh:
            if ln.startswith("#"):
                continue
            rows.append([float(x) for x in ln.split()[1:]])
    return np.array(rows, np.float64)


def read_rst7(path):
    """Amber ASCII restart: title, atom count, 6F12.7 coordinates."""
    lines = open(path).read().splitlines()
    n = int(lines[1].split()[0])
    vals = []
    for ln in lines[2:]:
        vals.extend(float(ln[i:i + 12]) for i in range(0, len(ln.rstrip()), 12))
    return np.array(vals[:3 * n], np.float64).reshape(n, 3)


def synth(seed, nframes, natoms, natom_total=None, stride_extra=0):
    """Small deterministic trajectory in the BASELINE cfg-2 style (see cpptraj_b200.synth)."""
    from cpptraj_b200.synth import make_trajectory
    return make_trajectory(seed, nframes, natoms, natom_total=natom_total, stride_extra=stride_extra)


def main():
    from scipy.io import netcdf_file
    from oracle.pyoracle import Reference
    os.makedirs(OUT, exist_ok=True)
    names, resnum, mass = read_parm7(os.path.join(REF, "test", "tz2.parm7"))
    nc = netcdf_file(os.path.join(REF, "test", "tz2.nc"), "r", mmap=False)
    crd = np.array(nc.variables["coordinates"][:], np.float32)  # (101,223,3)
    nc.close()
    assert crd.shape == (101, 223, 3), crd.shape
    np.savez_compressed(os.path.join(OUT, "tz2.npz"), crd=crd,
                        names=np.array(names), resnum=resnum, mass=mass)

    t2d = os.path.join(REF, "test", "Test_2DRMS")
    trm = os.path.join(REF, "test", "Test_RMSD")
    saves = dict(
        rmsd=read_save_matrix(os.path.join(t2d, "rmsd.dat.save")),
        rmsd_mass=read_save_matrix(os.path.join(t2d, "rmsd.mass.dat.save")),
        trp=read_save_matrix(os.path.join(t2d, "trp.dat.save")),
        nofit=read_save_matrix(os.path.join(t2d, "nofit.dat.save")),
        nomod=read_save_matrix(os.path.join(trm, "NoMod.dat.save"))[:, 0],
        previous=read_save_matrix(os.path.join(trm, "Previous.dat.save"))[:, 0],
        # test/Test_RmsAvgCorr/RunTest.sh:11-34: fixed reference (avg.CA.rst7), the same with offset 10, "first" mode
        rmsavgcorr_ref=read_save_matrix(os.path.join(REF, "test", "Test_RmsAvgCorr", "rmscorr.dat.save")),
        rmsavgcorr_ref10=read_save_matrix(os.path.join(REF, "test", "Test_RmsAvgCorr", "rmscorr.10.dat.save")),
        rmsavgcorr_first=read_save_matrix(os.path.join(REF, "test", "Test_RmsAvgCorr", "rmscorr.first.dat.save")),
        avg_ca_rst7=read_rst7(os.path.join(REF, "test", "Test_RmsAvgCorr", "avg.CA.rst7")),
    )
    np.savez_compressed(os.path.join(OUT, "ref_saves.npz"), **saves)

    # ---- live reference outputs (full float/double precision) ----
    ref = Reference()
    flat = crd.reshape(101, -1)
    res = lambda lo, hi: np.nonzero((resnum >= lo) & (resnum <= hi))[0].astype(np.int32)
    names_a = np.array(names)
    live = {}
    s37 = res(3, 7)
    live["tz2_3to7_fit"] = ref.rms2d_tri(flat[:10], s37)
    live["tz2_3to7_mass"] = ref.rms2d_tri(flat[:10], s37, mass=mass[s37])
    live["tz2_trp_full"] = ref.rms2d_full(flat[:10], res(2, 2), flat[:10], res(11, 11))
    ca = np.nonzero((names_a == "CA") & (resnum >= 2) & (resnum <= 12))[0].astype(np.int32)
    live["tz2_ca_1vN"] = ref.rmsd_1vN(flat, ca, flat[0].reshape(-1, 3)[ca].astype(np.float64))
    allca = np.nonzero(names_a == "CA")[0].astype(np.int32)
    live["tz2_allca_fit"] = ref.rms2d_tri(flat, allca)
    live["tz2_allca_cluster_mass"] = ref.cluster_tri(flat, allca, mass=mass[allca])
    # synthetic, seeded (same generator the tests call)
    for tag, seed, nf, na, ntot in (("s1", 11, 48, 100, 100), ("s2", 12, 40, 257, 300), ("s3", 13, 33, 7, 20)):
        c, m = synth(seed, nf, na, natom_total=ntot)
        sel = np.arange(0, ntot, max(1, ntot // na), dtype=np.int32)[:na] if ntot != na else np.arange(na, dtype=np.int32)
        live[tag + "_sel"] = sel
        live[tag + "_fit"] = ref.rms2d_tri(c, sel)
        live[tag + "_mass"] = ref.rms2d_tri(c, sel, mass=m[sel])
        live[tag + "_nofit"] = ref.rms2d_tri(c, sel, fit=False)
        r, rot, tr, rt = ref.rmsd_1vN(c, sel, c[0].reshape(-1, 3)[sel].astype(np.float64), mass=m[sel], want_rot=True)
        live[tag + "_1vN_rms"], live[tag + "_1vN_rot"], live[tag + "_1vN_tr"], live[tag + "_1vN_rt"] = r, rot, tr, rt
    np.savez_compressed(os.path.join(OUT, "ref_live.npz"), **live)
    print("wrote", OUT, {k: v.shape for k, v in live.items()})


if __name__ == "__main__":
    main()
