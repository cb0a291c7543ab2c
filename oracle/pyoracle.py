"""ctypes loaders for the CPU checkers (TEST INFRASTRUCTURE ONLY).

`oracle`  = oracle/liboracle_rmsd.so   (our C restatement, rmsd_oracle.c)
`ref`     = oracle/_ref/libcpptraj_ref_rmsd.so (the reference's own Frame /
            Matrix_3x3 / CompactFrameArray / Matrix<float> code, build_ref.sh)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "liboracle_rmsd.so")
REF_SO = os.path.join(_HERE, "_ref", "libcpptraj_ref_rmsd.so")

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def _opt(a, dtype):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def tri_size(n):
    return n * (n - 1) // 2


def tri_index(n, i, j):
    """src/Matrix.h:110-122 calcTriIndex (i<j)."""
    return n * i - (i + 1) * i // 2 + j - i - 1


def _hieragglo(fn, tri, n, linkage, target_clusters, epsilon):
    """Shared wrapper: (mergeInto, mergeFrom, findMin[nCalls]) of Algorithm_HierAgglo::DoClustering on a cache triangle.
    linkage 0 single / 1 average / 2 complete; target_clusters, epsilon None = the reference's defaults (1, DBL_MAX)."""
    tri = np.ascontiguousarray(tri, np.float32)
    assert tri.size == tri_size(n)
    into = np.zeros(max(n, 1), np.int32)
    frm = np.zeros(max(n, 1), np.int32)
    fmin = np.zeros(max(n, 1), np.float32)
    nc = C.c_int(0)
    nm = C.c_int(0)
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                   C.POINTER(C.c_int), C.POINTER(C.c_int)]
    rc = fn(_ptr(tri), n, linkage, 1 if target_clusters is None else target_clusters,
            np.finfo(np.float64).max if epsilon is None else epsilon, _ptr(into), _ptr(frm), _ptr(fmin),
            C.byref(nc), C.byref(nm))
    if rc:
        raise RuntimeError("hieragglo checker failed rc=%d" % rc)
    return into[:nm.value].copy(), frm[:nm.value].copy(), fmin[:nc.value].copy()


def _rmsavgcorr(fn, extra, crd, sel, windows, mass, ref_sel_xyz):
    crd = np.ascontiguousarray(crd, np.float32)
    sel = np.ascontiguousarray(sel, np.int32)
    windows = np.ascontiguousarray(windows, np.int32)
    mass = _opt(mass, np.float64)
    ref = None if ref_sel_xyz is None else np.ascontiguousarray(ref_sel_xyz, np.float64).reshape(-1)
    avg = np.zeros(len(windows), np.float64)
    sd = np.zeros(len(windows), np.float64)
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_size_t, C.c_int] + [C.c_int] * len(extra) + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                                                C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    rc = fn(_ptr(crd), crd.shape[1], crd.shape[0], *extra, _ptr(sel), len(sel), _ptr(mass), _ptr(ref), _ptr(windows),
            len(windows), _ptr(avg), _ptr(sd))
    if rc:
        raise RuntimeError("rmsavgcorr checker failed rc=%d" % rc)
    return avg, sd


class _Base:
    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.path = path


class Oracle(_Base):
    """Our C restatement."""

    def __init__(self):
        super().__init__(ORACLE_SO)
        L = self.lib
        L.orc_num_threads.restype = C.c_int
        for name in ("orc_rms2d_tri", "orc_cluster_tri"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                          C.c_void_p, C.c_int, C.c_void_p]
        L.orc_rms2d_full.restype = C.c_int
        L.orc_rms2d_full.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_rmsd_1vN.restype = C.c_int
        L.orc_rmsd_1vN.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]
        L.orc_build_centroid.restype = C.c_int
        L.orc_build_centroid.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_int, C.c_void_p]

    def build_centroid(self, crd, sel, frames, mass=None, fit=True):
        """Metric_RMS::CalculateCentroid (src/Cluster/Metric_RMS.cpp:86-113): [len(sel), 3] float64."""
        crd = np.ascontiguousarray(crd, np.float32)
        sel = np.ascontiguousarray(sel, np.int32)
        frames = np.ascontiguousarray(frames, np.int32)
        mass = _opt(mass, np.float64)
        out = np.zeros((len(sel), 3), np.float64)
        rc = self.lib.orc_build_centroid(_ptr(crd), crd.shape[1], _ptr(frames), len(frames), _ptr(sel), len(sel),
                                         _ptr(mass), int(bool(fit)), _ptr(out))
        if rc:
            raise RuntimeError("oracle failed rc=%d" % rc)
        return out

    def rmsavgcorr(self, crd, sel, windows, mass=None, ref_sel_xyz=None):
        """Analysis_RmsAvgCorr::Analyze: (avg, sd) per window size; ref_sel_xyz None = 'first' mode."""
        return _rmsavgcorr(self.lib.orc_rmsavgcorr, (), crd, sel, windows, mass, ref_sel_xyz)

    def hieragglo(self, tri, n, linkage=1, target_clusters=None, epsilon=None):
        return _hieragglo(self.lib.orc_hieragglo, tri, n, linkage, target_clusters, epsilon)

    def threads(self):
        return self.lib.orc_num_threads()

    def set_threads(self, n):
        self.lib.orc_set_threads(C.c_int(n))

    def _tri(self, fn, crd, sel, mass, fit, frame_idx):
        crd = np.ascontiguousarray(crd, np.float32)
        assert crd.ndim == 2
        sel = np.ascontiguousarray(sel, np.int32)
        mass = _opt(mass, np.float64)
        fidx = _opt(frame_idx, np.int32)
        nF = crd.shape[0] if fidx is None else len(fidx)
        out = np.zeros(tri_size(nF), np.float32)
        rc = fn(_ptr(crd), crd.shape[1], _ptr(fidx), nF, _ptr(sel), len(sel), _ptr(mass),
                int(bool(fit)), _ptr(out))
        if rc:
            raise RuntimeError("oracle failed rc=%d" % rc)
        return out

    def rms2d_tri(self, crd, sel, mass=None, fit=True, frame_idx=None):
        return self._tri(self.lib.orc_rms2d_tri, crd, sel, mass, fit, frame_idx)

    def cluster_tri(self, crd, sel, mass=None, fit=True, frame_idx=None):
        return self._tri(self.lib.orc_cluster_tri, crd, sel, mass, fit, frame_idx)

    def rms2d_full(self, crdT, selT, crdR, selR, massTgt=None, massRef=None, fit=True):
        crdT = np.ascontiguousarray(crdT, np.float32)
        crdR = np.ascontiguousarray(crdR, np.float32)
        selT = np.ascontiguousarray(selT, np.int32)
        selR = np.ascontiguousarray(selR, np.int32)
        assert len(selT) == len(selR)
        massTgt = _opt(massTgt, np.float64)
        massRef = _opt(massRef, np.float64)
        out = np.zeros((crdT.shape[0], crdR.shape[0]), np.float32)
        rc = self.lib.orc_rms2d_full(_ptr(crdT), crdT.shape[1], crdT.shape[0], _ptr(selT),
                                     _ptr(crdR), crdR.shape[1], crdR.shape[0], _ptr(selR),
                                     len(selT), _ptr(massTgt), _ptr(massRef), int(bool(fit)), _ptr(out))
        if rc:
            raise RuntimeError("oracle failed rc=%d" % rc)
        return out

    def rmsd_1vN(self, crd, sel, ref_sel_xyz, mass=None, fit=True, want_rot=False):
        crd = np.ascontiguousarray(crd, np.float32)
        sel = np.ascontiguousarray(sel, np.int32)
        ref = np.ascontiguousarray(ref_sel_xyz, np.float64).reshape(-1)
        assert ref.size == 3 * len(sel)
        mass = _opt(mass, np.float64)
        nF = crd.shape[0]
        rms = np.zeros(nF, np.float64)
        rot = np.zeros((nF, 9), np.float64) if want_rot else None
        tr = np.zeros((nF, 3), np.float64) if want_rot else None
        rt = np.zeros(3, np.float64)
        rc = self.lib.orc_rmsd_1vN(_ptr(crd), crd.shape[1], nF, _ptr(sel), len(sel), _ptr(ref),
                                   _ptr(mass), int(bool(fit)), _ptr(rms), _ptr(rot), _ptr(tr), _ptr(rt))
        if rc:
            raise RuntimeError("oracle failed rc=%d" % rc)
        return (rms, rot, tr, rt) if want_rot else rms


class Reference(_Base):
    """The reference's own classes behind oracle/ref_driver.cpp."""

    def __init__(self):
        super().__init__(REF_SO)
        L = self.lib
        L.ref_num_threads.restype = C.c_int
        L.ref_last_loop_seconds.restype = C.c_double
        for name in ("ref_rms2d_tri", "ref_cluster_tri"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int,
                          C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_rms2d_full.restype = C.c_int
        L.ref_rms2d_full.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.ref_rmsd_1vN.restype = C.c_int
        L.ref_rmsd_1vN.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]

    def rmsavgcorr(self, crd, sel, windows, mass=None, ref_sel_xyz=None, natom_total=None):
        crd = np.ascontiguousarray(crd, np.float32)
        sel = np.ascontiguousarray(sel, np.int32)
        return _rmsavgcorr(self.lib.ref_rmsavgcorr, (natom_total or self._natom_total(crd, sel),), crd, sel, windows, mass,
                           ref_sel_xyz)

    def hieragglo(self, tri, n, linkage=1, target_clusters=None, epsilon=None):
        """The reference's own Cluster::DynamicMatrix driven by ref_driver.cpp's restated merge loop."""
        return _hieragglo(self.lib.ref_hieragglo, tri, n, linkage, target_clusters, epsilon)

    def threads(self):
        return self.lib.ref_num_threads()

    def set_threads(self, n):
        self.lib.ref_set_threads(C.c_int(n))

    def last_loop_seconds(self):
        return self.lib.ref_last_loop_seconds()

    @staticmethod
    def _natom_total(crd, sel):
        # frames are [pos(3*natom) | optional extras]; natomTotal only has to
        # cover the selection and fit inside the stride.
        return int(crd.shape[1] // 3) if crd.shape[1] % 3 == 0 and crd.shape[1] // 3 > int(np.max(sel)) \
            else int(np.max(sel)) + 1

    def _tri(self, fn, crd, sel, mass, fit, frame_idx, natom_total):
        crd = np.ascontiguousarray(crd, np.float32)
        sel = np.ascontiguousarray(sel, np.int32)
        mass = _opt(mass, np.float64)
        fidx = _opt(frame_idx, np.int32)
        nF = crd.shape[0] if fidx is None else len(fidx)
        nat = natom_total or self._natom_total(crd, sel)
        out = np.zeros(tri_size(nF), np.float32)
        rc = fn(_ptr(crd), crd.shape[1], crd.shape[0], nat, _ptr(fidx), nF, _ptr(sel), len(sel),
                _ptr(mass), int(bool(fit)), _ptr(out))
        if rc:
            raise RuntimeError("reference driver failed rc=%d" % rc)
        return out

    def rms2d_tri(self, crd, sel, mass=None, fit=True, frame_idx=None, natom_total=None):
        return self._tri(self.lib.ref_rms2d_tri, crd, sel, mass, fit, frame_idx, natom_total)

    def cluster_tri(self, crd, sel, mass=None, fit=True, frame_idx=None, natom_total=None):
        return self._tri(self.lib.ref_cluster_tri, crd, sel, mass, fit, frame_idx, natom_total)

    def rms2d_full(self, crdT, selT, crdR, selR, massTgt=None, massRef=None, fit=True,
                   natomT=None, natomR=None):
        crdT = np.ascontiguousarray(crdT, np.float32)
        crdR = np.ascontiguousarray(crdR, np.float32)
        selT = np.ascontiguousarray(selT, np.int32)
        selR = np.ascontiguousarray(selR, np.int32)
        massTgt = _opt(massTgt, np.float64)
        massRef = _opt(massRef, np.float64)
        out = np.zeros((crdT.shape[0], crdR.shape[0]), np.float32)
        rc = self.lib.ref_rms2d_full(_ptr(crdT), crdT.shape[1], crdT.shape[0],
                                     natomT or self._natom_total(crdT, selT), _ptr(selT),
                                     _ptr(crdR), crdR.shape[1], crdR.shape[0],
                                     natomR or self._natom_total(crdR, selR), _ptr(selR),
                                     len(selT), _ptr(massTgt), _ptr(massRef), int(bool(fit)), _ptr(out))
        if rc:
            raise RuntimeError("reference driver failed rc=%d" % rc)
        return out

    def build_centroid(self, crd, sel, frames, mass=None, fit=True, natom_total=None):
        crd = np.ascontiguousarray(crd, np.float32)
        sel = np.ascontiguousarray(sel, np.int32)
        frames = np.ascontiguousarray(frames, np.int32)
        mass = _opt(mass, np.float64)
        out = np.zeros((len(sel), 3), np.float64)
        f = self.lib.ref_build_centroid
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                      C.c_void_p, C.c_int, C.c_void_p]
        rc = f(_ptr(crd), crd.shape[1], crd.shape[0], natom_total or self._natom_total(crd, sel), _ptr(frames), len(frames),
               _ptr(sel), len(sel), _ptr(mass), int(bool(fit)), _ptr(out))
        if rc:
            raise RuntimeError("reference driver failed rc=%d" % rc)
        return out

    def rmsd_1vN(self, crd, sel, ref_sel_xyz, mass=None, fit=True, want_rot=False, natom_total=None):
        crd = np.ascontiguousarray(crd, np.float32)
        sel = np.ascontiguousarray(sel, np.int32)
        ref = np.ascontiguousarray(ref_sel_xyz, np.float64).reshape(-1)
        mass = _opt(mass, np.float64)
        nF = crd.shape[0]
        rms = np.zeros(nF, np.float64)
        rot = np.zeros((nF, 9), np.float64) if want_rot else None
        tr = np.zeros((nF, 3), np.float64) if want_rot else None
        rt = np.zeros(3, np.float64)
        rc = self.lib.ref_rmsd_1vN(_ptr(crd), crd.shape[1], nF, natom_total or self._natom_total(crd, sel),
                                   _ptr(sel), len(sel), _ptr(ref), _ptr(mass), int(bool(fit)),
                                   _ptr(rms), _ptr(rot), _ptr(tr), _ptr(rt))
        if rc:
            raise RuntimeError("reference driver failed rc=%d" % rc)
        return (rms, rot, tr, rt) if want_rot else rms


def have_reference():
    return os.path.exists(REF_SO)
