"""ctypes binding of include/b200_rmsd.h (host + device-pointer entry points)."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_RMSD_LIB") or os.path.join(_HERE, "libb200rmsd.so")   # (override: kernel A/B experiments)

ERRORS = {1: "B200_ERR_NO_DEVICE", 2: "B200_ERR_CUDA", 3: "B200_ERR_ARG", 4: "B200_ERR_NOMEM", 5: "B200_ERR_STATE"}


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERRORS.get(code, "B200_ERR"), code, msg))
        self.code = code


class Stats(C.Structure):
    _fields_ = [("pack_ms", C.c_double), ("pair_ms", C.c_double), ("onevn_ms", C.c_double),
                ("pack_launches", C.c_long), ("pair_launches", C.c_long), ("onevn_launches", C.c_long),
                ("pairs", C.c_double), ("frames_1vN", C.c_double), ("h2d_bytes", C.c_double),
                ("d2h_bytes", C.c_double), ("kernel_launches", C.c_long),
                ("onevn_stream_ms", C.c_double), ("onevn_stream_launches", C.c_long)]


_lib = None


def lib():
    """Load libb200rmsd.so; there is deliberately no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("cpptraj_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, sz, i, dbl = C.c_void_p, C.c_size_t, C.c_int, C.c_double
    L.b200_init.argtypes = [i, C.POINTER(i)]
    L.b200_init_devices.argtypes = [C.POINTER(i), i]
    L.b200_last_error.restype = C.c_char_p
    L.b200_shard_rows.argtypes = [i, i, i, C.POINTER(i), C.POINTER(i)]
    L.b200_rms2d_tri.argtypes = [vp, sz, i, vp, i, vp, i, vp, i, vp]
    L.b200_rms2d_tri_shard.argtypes = [vp, sz, i, vp, i, vp, i, vp, i, i, i, vp, C.POINTER(sz), C.POINTER(sz)]
    L.b200_rms2d_full.argtypes = [vp, sz, i, vp, vp, sz, i, vp, i, vp, vp, i, vp]
    L.b200_rmsd_1vN_begin.argtypes = [vp, vp, i, vp, i, i, C.POINTER(vp)]
    L.b200_rmsd_1vN_push_f64.argtypes = [vp, vp, sz, i]
    L.b200_rmsd_1vN_push_f32.argtypes = [vp, vp, sz, i]
    L.b200_rmsd_1vN_set_ref.argtypes = [vp, vp]
    L.b200_rmsd_1vN_set_ref.restype = i
    L.b200_rmsd_build_centroids.argtypes = [vp, sz, i, vp, vp, i, vp, i, vp, i, vp]
    L.b200_rmsd_build_centroids.restype = i
    L.b200_rmsavgcorr.argtypes = [vp, sz, i, vp, i, vp, vp, vp, i, vp, vp]
    L.b200_rmsavgcorr.restype = i
    L.b200_cache_resident_begin.argtypes = [vp, i]
    L.b200_cache_resident_begin.restype = i
    L.b200_cache_resident_end.argtypes = [vp]
    L.b200_cache_resident_end.restype = i
    L.b200_cache_cluster_sums.argtypes = [vp, i, vp, vp, i, vp, vp, vp]
    L.b200_cache_cluster_sums.restype = i
    L.b200_cache_cluster_links.argtypes = [vp, i, vp, i, vp, vp, vp, vp]
    L.b200_cache_cluster_links.restype = i
    L.b200_hieragglo.argtypes = [vp, i, i, i, dbl, vp, vp, vp, C.POINTER(i), C.POINTER(i)]
    L.b200_hieragglo.restype = i
    L.b200_coords_resident_begin.argtypes = [vp, sz, i, vp, i]
    L.b200_coords_resident_begin.restype = i
    L.b200_coords_resident_end.argtypes = [vp]
    L.b200_coords_resident_end.restype = i
    L.b200_set_fixed_point_bits.argtypes = [i]
    L.b200_set_fixed_point_bits.restype = i
    L.b200_rmsd_1vN_pending.argtypes = [vp]
    L.b200_rmsd_1vN_pending.restype = C.c_long
    L.b200_rmsd_1vN_flush.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_long)]
    L.b200_rmsd_1vN_end.argtypes = [vp]
    L.b200_dev_rms2d_tri.argtypes = [vp, sz, vp, i, vp, i, vp, i, i, i, vp, vp]
    L.b200_rmsd_frames_to_centroids.argtypes = [vp, sz, i, vp, i, vp, i, vp, i, vp, i, vp, vp, vp]
    L.b200_rmsd_frames_to_centroids.restype = i
    L.b200_dev_rmsd_1vN.argtypes = [vp, sz, i, vp, i, vp, vp, i, vp, vp, vp, vp]
    L.b200_set_pair_engine.argtypes = [i]
    L.b200_set_i8_cta_group.argtypes = [i]
    L.b200_last_pair_engine.argtypes = [C.POINTER(i)]
    L.b200_debug_i8.argtypes = [vp, sz, i, vp, i, vp, vp, sz, C.POINTER(sz), vp, vp, vp, C.POINTER(i)]
    L.b200_get_stats.argtypes = [C.POINTER(Stats)]
    L.b200_set_profiling.argtypes = [i]
    L.b200_measure_fp64_mma_peak.argtypes = [i]
    L.b200_measure_fp64_mma_peak.restype = dbl
    L.b200_measure_i8_mma_peak.restype = dbl
    L.b200_measure_i8_mma_peak_variant.argtypes = [i]
    L.b200_measure_i8_mma_peak_variant.restype = dbl
    for name in ("b200_init", "b200_init_devices", "b200_shard_rows", "b200_rms2d_tri", "b200_rms2d_tri_shard",
                 "b200_rms2d_full", "b200_rmsd_1vN_begin", "b200_rmsd_1vN_push_f64", "b200_rmsd_1vN_push_f32",
                 "b200_rmsd_1vN_flush", "b200_rmsd_1vN_end", "b200_dev_rms2d_tri", "b200_dev_rmsd_1vN",
                 "b200_version", "b200_num_devices", "b200_set_pair_engine", "b200_set_i8_cta_group", "b200_get_i8_cta_group", "b200_last_pair_engine", "b200_debug_i8",
                 "b200_set_mma_variant"):
        getattr(L, name).restype = i
    _lib = L
    return L


def _check(rc):
    if rc:
        raise B200Error(rc, lib().b200_last_error().decode("utf-8", "replace"))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _arr(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


def tri_size(n):
    return n * (n - 1) // 2 if n > 1 else 0


def tri_index(n, i, j):
    return n * i - (i + 1) * i // 2 + j - i - 1


def init(ngpu=1, devices=None):
    L = lib()
    if devices is not None:
        ids = (C.c_int * len(devices))(*devices)
        _check(L.b200_init_devices(ids, len(devices)))
        return len(devices)
    used = C.c_int(0)
    _check(L.b200_init(int(ngpu), C.byref(used)))
    return used.value


def shutdown():
    lib().b200_shutdown()


def num_devices():
    return lib().b200_num_devices()


def shard_rows(nframes, rank, count):
    a, b = C.c_int(0), C.c_int(0)
    _check(lib().b200_shard_rows(int(nframes), int(rank), int(count), C.byref(a), C.byref(b)))
    return a.value, b.value


def _crd2d(crd):
    crd = np.asarray(crd)
    if crd.dtype != np.float32 or crd.ndim != 2 or not crd.flags.c_contiguous:
        crd = np.ascontiguousarray(crd, np.float32)
        assert crd.ndim == 2
    return crd


def rms2d_tri(crd, atom_idx, mass=None, fit=True, frame_idx=None, out=None):
    """rms2d / pairwise-cache triangle over all initialised devices."""
    crd = _crd2d(crd)
    sel = _arr(atom_idx, np.int32)
    mass = _arr(mass, np.float64)
    fidx = _arr(frame_idx, np.int32)
    nF = crd.shape[0] if fidx is None else len(fidx)
    if out is None:
        out = np.empty(tri_size(nF), np.float32)
    assert out.dtype == np.float32 and out.size >= tri_size(nF)
    _check(lib().b200_rms2d_tri(_p(crd), crd.shape[1], crd.shape[0], _p(fidx), nF, _p(sel), len(sel), _p(mass),
                                int(bool(fit)), _p(out)))
    return out


def rms2d_tri_shard(crd, atom_idx, rank, count, mass=None, fit=True, frame_idx=None, out=None, query_only=False):
    crd = _crd2d(crd)
    sel = _arr(atom_idx, np.int32)
    mass = _arr(mass, np.float64)
    fidx = _arr(frame_idx, np.int32)
    nF = crd.shape[0] if fidx is None else len(fidx)
    first, n = C.c_size_t(0), C.c_size_t(0)
    if out is None and not query_only:
        out = np.zeros(tri_size(nF), np.float32)
    _check(lib().b200_rms2d_tri_shard(_p(crd), crd.shape[1], crd.shape[0], _p(fidx), nF, _p(sel), len(sel), _p(mass),
                                      int(bool(fit)), int(rank), int(count), None if query_only else _p(out),
                                      C.byref(first), C.byref(n)))
    return out, first.value, n.value


def rms2d_full(crd_tgt, idx_tgt, crd_ref, idx_ref, mass_tgt=None, mass_ref=None, fit=True, out=None):
    ct, cr = _crd2d(crd_tgt), _crd2d(crd_ref)
    st, sr = _arr(idx_tgt, np.int32), _arr(idx_ref, np.int32)
    assert len(st) == len(sr)
    mt, mr = _arr(mass_tgt, np.float64), _arr(mass_ref, np.float64)
    if out is None:
        out = np.empty((ct.shape[0], cr.shape[0]), np.float32)
    _check(lib().b200_rms2d_full(_p(ct), ct.shape[1], ct.shape[0], _p(st), _p(cr), cr.shape[1], cr.shape[0], _p(sr),
                                 len(st), _p(mt), _p(mr), int(bool(fit)), _p(out)))
    return out


class Rmsd1vN:
    """Streaming one-vs-many handle (Action_Rmsd body)."""

    def __init__(self, ref_selected, atom_idx, mass=None, fit=True, want_rot=False):
        self.ref = np.ascontiguousarray(ref_selected, np.float64).reshape(-1)
        self.sel = _arr(atom_idx, np.int32)
        assert self.ref.size == 3 * len(self.sel)
        self.mass = _arr(mass, np.float64)
        self.want_rot = bool(want_rot) and bool(fit)
        self.h = C.c_void_p(None)
        _check(lib().b200_rmsd_1vN_begin(_p(self.ref), _p(self.sel), len(self.sel), _p(self.mass), int(bool(fit)),
                                         int(self.want_rot), C.byref(self.h)))

    def push(self, frames):
        frames = np.asarray(frames)
        assert frames.ndim == 2 and frames.flags.c_contiguous
        if frames.dtype == np.float32:
            _check(lib().b200_rmsd_1vN_push_f32(self.h, _p(frames), frames.shape[1], frames.shape[0]))
        elif frames.dtype == np.float64:
            _check(lib().b200_rmsd_1vN_push_f64(self.h, _p(frames), frames.shape[1], frames.shape[0]))
        else:
            raise TypeError(frames.dtype)

    def set_ref(self, ref_selected):
        """Replace the reference (reftraj / previous); frames already pushed keep theirs."""
        ref = np.ascontiguousarray(ref_selected, np.float64).reshape(-1)
        assert ref.size == 3 * len(self.sel)
        _check(lib().b200_rmsd_1vN_set_ref(self.h, _p(ref)))

    def pending(self):
        return lib().b200_rmsd_1vN_pending(self.h)

    def flush(self):
        n = self.pending()
        rms = np.empty(n, np.float64)
        rot = np.empty((n, 9), np.float64) if self.want_rot else None
        tr = np.empty((n, 3), np.float64) if self.want_rot else None
        best = C.c_long(-1)
        _check(lib().b200_rmsd_1vN_flush(self.h, _p(rms), _p(rot), _p(tr), C.byref(best)))
        return rms, rot, tr, best.value

    def close(self):
        if self.h:
            lib().b200_rmsd_1vN_end(self.h)
            self.h = C.c_void_p(None)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def rmsd_1vN(crd, atom_idx, ref_selected, mass=None, fit=True, want_rot=False, chunk=None):
    crd = np.asarray(crd)
    with Rmsd1vN(ref_selected, atom_idx, mass, fit, want_rot) as h:
        if chunk:
            for f0 in range(0, crd.shape[0], chunk):
                h.push(crd[f0:f0 + chunk])
        else:
            h.push(crd)
        return h.flush()


def frames_to_centroids(crd, atom_idx, centroids, mass=None, fit=True, frame_idx=None, want_dist=True):
    """Frame-to-centroid RMSDs (Metric_RMS::FrameCentroidDist for many frames): returns (dist [nF, K] or None,
    closest [nF] int32, closest_dist [nF])."""
    crd = _crd2d(crd)
    sel = _arr(atom_idx, np.int32)
    cen = np.ascontiguousarray(centroids, np.float64).reshape(-1, 3 * len(sel))
    mass = _arr(mass, np.float64)
    fidx = _arr(frame_idx, np.int32)
    nF = len(fidx) if fidx is not None else crd.shape[0]
    dist = np.empty((nF, cen.shape[0]), np.float64) if want_dist else None
    closest = np.empty(nF, np.int32)
    cdist = np.empty(nF, np.float64)
    _check(lib().b200_rmsd_frames_to_centroids(_p(crd), crd.shape[1], crd.shape[0], _p(fidx), nF, _p(sel), len(sel), _p(mass),
                                               int(bool(fit)), _p(cen), cen.shape[0], _p(dist), _p(closest), _p(cdist)))
    return dist, closest, cdist


def build_centroids(crd, atom_idx, frame_lists, mass=None, fit=True):
    """Metric_RMS::CalculateCentroid for several clusters: frame_lists is a sequence of frame-number arrays;
    returns [K, nAtoms, 3] float64."""
    crd = _crd2d(crd)
    sel = _arr(atom_idx, np.int32)
    mass = _arr(mass, np.float64)
    frames = np.ascontiguousarray(np.concatenate([np.asarray(f, np.int32) for f in frame_lists]) if len(frame_lists) else np.zeros(0, np.int32), np.int32)
    offsets = np.zeros(len(frame_lists) + 1, np.int32)
    offsets[1:] = np.cumsum([len(f) for f in frame_lists])
    out = np.zeros((len(frame_lists), len(sel), 3), np.float64)
    _check(lib().b200_rmsd_build_centroids(_p(crd), crd.shape[1], crd.shape[0], _p(frames), _p(offsets), len(frame_lists),
                                           _p(sel), len(sel), _p(mass), int(bool(fit)), _p(out)))
    return out


def rmsavgcorr(crd, atom_idx, windows, mass=None, ref_selected=None):
    """Analysis_RmsAvgCorr::Analyze: (avg, sd) per window size; ref_selected None = 'first' mode."""
    crd = _crd2d(crd)
    idx = _arr(atom_idx, np.int32)
    win = _arr(windows, np.int32)
    m = _arr(mass, np.float64)
    ref = None if ref_selected is None else _arr(ref_selected, np.float64).reshape(-1)
    avg = np.zeros(len(win), np.float64)
    sd = np.zeros(len(win), np.float64)
    _check(lib().b200_rmsavgcorr(_p(crd), crd.shape[1], crd.shape[0], _p(idx), len(idx), _p(m), _p(ref), _p(win), len(win),
                                 _p(avg), _p(sd)))
    return avg, sd


def cache_resident_begin(tri, nframes):
    tri = _arr(tri, np.float32)
    _check(lib().b200_cache_resident_begin(_p(tri), nframes))
    return tri            # (keep this array alive and pass IT to the calls that should find the resident copy)


def cache_resident_end(tri=None):
    _check(lib().b200_cache_resident_end(_p(tri) if tri is not None else None))


def cache_cluster_sums(tri, nframes, member_lists):
    """(cum, up, up2) per listed member, concatenated in list order (BestReps cumulative distance; Summary's within-cluster sums)."""
    tri = _arr(tri, np.float32)
    members = np.ascontiguousarray(np.concatenate([np.asarray(m, np.int32) for m in member_lists]) if member_lists else np.zeros(0, np.int32), np.int32)
    offsets = np.zeros(len(member_lists) + 1, np.int32)
    offsets[1:] = np.cumsum([len(m) for m in member_lists])
    cum = np.zeros(len(members), np.float64); up = np.zeros_like(cum); up2 = np.zeros_like(cum)
    _check(lib().b200_cache_cluster_sums(_p(tri), nframes, _p(members), _p(offsets), len(member_lists), _p(cum), _p(up), _p(up2)))
    return cum, up, up2


def cache_cluster_links(tri, nframes, label, nclusters):
    """(min, max, sum, count) tables [K, K] (entries c1 < c2) of the cached distances between clusters."""
    tri = _arr(tri, np.float32)
    label = _arr(label, np.int32)
    K = nclusters
    mn = np.zeros((K, K), np.float64); mx = np.zeros((K, K), np.float64); sm = np.zeros((K, K), np.float64)
    cnt = np.zeros((K, K), np.int64)
    _check(lib().b200_cache_cluster_links(_p(tri), nframes, _p(label), K, _p(mn), _p(mx), _p(sm), _p(cnt)))
    return mn, mx, sm, cnt


def hieragglo(tri, nframes, linkage=1, target_clusters=None, epsilon=None):
    """Algorithm_HierAgglo::DoClustering on a cache triangle: (mergeInto, mergeFrom, findMin[nCalls]).
    linkage 0 single / 1 average / 2 complete; None = the reference's defaults (1 cluster, no epsilon)."""
    tri = _arr(tri, np.float32)
    assert tri.size == tri_size(nframes)
    into = np.zeros(max(nframes, 1), np.int32)
    frm = np.zeros(max(nframes, 1), np.int32)
    fmin = np.zeros(max(nframes, 1), np.float32)
    nc, nm = C.c_int(0), C.c_int(0)
    _check(lib().b200_hieragglo(_p(tri), nframes, linkage, 1 if target_clusters is None else target_clusters,
                                np.finfo(np.float64).max if epsilon is None else epsilon, _p(into), _p(frm), _p(fmin),
                                C.byref(nc), C.byref(nm)))
    return into[:nm.value].copy(), frm[:nm.value].copy(), fmin[:nc.value].copy()


def coords_resident_begin(crd, atom_idx):
    crd = _crd2d(crd)
    sel = _arr(atom_idx, np.int32)
    _check(lib().b200_coords_resident_begin(_p(crd), crd.shape[1], crd.shape[0], _p(sel), len(sel)))


def coords_resident_end(crd=None):
    _check(lib().b200_coords_resident_end(None if crd is None else _p(_crd2d(crd))))


def set_profiling(on):
    lib().b200_set_profiling(int(bool(on)))


def reset_stats():
    lib().b200_reset_stats()


def get_stats():
    s = Stats()
    lib().b200_get_stats(C.byref(s))
    return {k: getattr(s, k) for k, _ in Stats._fields_}


ENGINES = {"auto": 0, "fp64": 1, "i8": 2}


def set_pair_engine(engine):
    """'auto' | 'fp64' | 'i8' (or 0/1/2): which kernel computes the pair covariances."""
    _check(lib().b200_set_pair_engine(ENGINES.get(engine, engine)))


def set_fixed_point_bits(bits):
    """Pin the fractional bits of the tcgen05 engine's fixed-point grid (0 = automatic)."""
    _check(lib().b200_set_fixed_point_bits(int(bits)))


def set_i8_cta_group(cta_group):
    """MMA CTA group of the tcgen05 int8 kernel: 2 = CTA pairs (default), 1 = single CTA."""
    _check(lib().b200_set_i8_cta_group(int(cta_group)))


def get_i8_cta_group():
    return int(lib().b200_get_i8_cta_group())


def last_pair_engine():
    """(engine, fractional_bits) of the last rms2d call: engine 1 = FP64 DMMA, 2 = tcgen05 int8."""
    q = C.c_int(0)
    e = lib().b200_last_pair_engine(C.byref(q))
    return e, q.value


def debug_i8(crd, atom_idx, mass=None, want_image=True, want_S=True):
    """Test hook: packed int8 operand image, G, raw integer covariances, triangle, fractional bits."""
    crd = _crd2d(crd)
    sel = _arr(atom_idx, np.int32)
    mass = _arr(mass, np.float64)
    nF = crd.shape[0]
    nbytes = C.c_size_t(0)
    qs = C.c_int(0)
    nrg = (nF + 13) // 14
    nrg += nrg & 1
    cap = nrg * ((len(sel) + 63) // 64) * 8192
    img = np.zeros(cap, np.uint8) if want_image else None
    G = np.zeros(nF, np.float64)
    S = np.zeros((nF, nF, 9), np.float64) if want_S else None
    tri = np.zeros(tri_size(nF), np.float32)
    _check(lib().b200_debug_i8(_p(crd), crd.shape[1], nF, _p(sel), len(sel), _p(mass), _p(img), cap, C.byref(nbytes),
                               _p(G), _p(S), _p(tri), C.byref(qs)))
    return dict(image=img, G=G, S=S, tri=tri, qs=qs.value, image_bytes=nbytes.value)


def set_mma_variant(v):
    _check(lib().b200_set_mma_variant(int(v)))


def measure_i8_mma_peak(variant=0):
    return lib().b200_measure_i8_mma_peak_variant(int(variant))


def measure_fp64_mma_peak(variant=0):
    return lib().b200_measure_fp64_mma_peak(int(variant))


# ---- device-pointer entry points (torch tensors or raw ints) -----------------
def _dp(t):
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(int(t))


def dev_rms2d_tri(d_crd, stride, nframes, d_atom_idx, natoms, d_out, d_mass=None, fit=True, d_frame_idx=None,
                  rank=0, count=1, stream=None):
    _check(lib().b200_dev_rms2d_tri(_dp(d_crd), int(stride), _dp(d_frame_idx), int(nframes), _dp(d_atom_idx),
                                    int(natoms), _dp(d_mass), int(bool(fit)), int(rank), int(count), _dp(d_out),
                                    C.c_void_p(stream or 0)))


def dev_rmsd_1vN(d_crd, stride, nframes, d_atom_idx, natoms, d_ref, d_rmsd, d_mass=None, fit=True, d_rot=None,
                 d_trans=None, stream=None):
    _check(lib().b200_dev_rmsd_1vN(_dp(d_crd), int(stride), int(nframes), _dp(d_atom_idx), int(natoms), _dp(d_ref),
                                   _dp(d_mass), int(bool(fit)), _dp(d_rmsd), _dp(d_rot), _dp(d_trans),
                                   C.c_void_p(stream or 0)))
