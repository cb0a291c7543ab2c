// b200_rmsd.cu -- host side of the C ABI declared in include/b200_rmsd.h.
//
// No torch, no CPU fallback: every compute entry point needs an sm_100 device.
// Device memory, streams and pinned staging are owned here; cpptraj owns the
// host buffers it passes in.
#include "../../include/b200_rmsd.h"
#include "rmsd_kernels.cuh"
#include "pair_i8.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace b200;

namespace {

// ------------------------------------------------------------------ errors
std::mutex g_errMu;
std::string g_err;
int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  std::lock_guard<std::mutex> lk(g_errMu);
  g_err = buf;
  return code;
}
#define CU(call)                                                                               \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return fail(B200_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
  } while (0)

// ------------------------------------------------------------------ buffers
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return B200_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    // grow a little beyond the request to avoid re-allocation churn
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(B200_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); }
    cap = want;
    return B200_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return B200_OK;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(B200_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
    cap = bytes;
    return B200_OK;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

constexpr int NSLOT = 3;  // compute/copy streams per device

struct Device {
  int id = -1;
  cudaStream_t stream[NSLOT] = {nullptr, nullptr, nullptr};
  cudaEvent_t done[NSLOT] = {nullptr, nullptr, nullptr};
  // workspaces (grow-only)
  DevBuf crd, crdB, idxA, idxB, frameIdx, massA, massB, planesA, planesB, GA, GB, scal, onevnWs;
  DevBuf imgA, imgB, cenA, cenB, dbgS;   // tcgen05 int8 path: operand images, frame centres
  PinBuf hostScal;                        // pinned slot for the few scalars read back per call
  int numSMs = 0;
  DevBuf outChunk[NSLOT];
  PinBuf outStage[NSLOT];
  bool attrSet = false;
  void destroy() {
    if (id < 0) return;
    cudaSetDevice(id);
    for (int s = 0; s < NSLOT; ++s) {
      if (stream[s]) cudaStreamDestroy(stream[s]);
      if (done[s]) cudaEventDestroy(done[s]);
      stream[s] = nullptr; done[s] = nullptr;
      outChunk[s].release(); outStage[s].release();
    }
    DevBuf* all[] = {&crd, &crdB, &idxA, &idxB, &frameIdx, &massA, &massB, &planesA, &planesB, &GA, &GB, &scal, &onevnWs,
                     &imgA, &imgB, &cenA, &cenB, &dbgS};
    for (DevBuf* b : all) b->release();
    hostScal.release();
    id = -1;
  }
};

std::mutex g_mu;               // serialises public entry points (re-entrant across sequential calls)
std::vector<Device> g_devs;
bool g_inited = false;

// ------------------------------------------------------------------ stats
std::mutex g_statMu;
b200_stats g_stats = {};
bool g_profiling = false;
std::atomic<long> g_launches{0};
#define COUNT_LAUNCH() (g_launches.fetch_add(1, std::memory_order_relaxed))

struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
};
struct Timer {  // collects (start,stop) events; resolved after a sync
  std::vector<EventPair> ev;
  cudaEvent_t begin(cudaStream_t s) {
    if (!g_profiling) return nullptr;
    EventPair p;
    cudaEventCreate(&p.a); cudaEventCreate(&p.b);
    cudaEventRecord(p.a, s);
    ev.push_back(p);
    return p.a;
  }
  void end(cudaStream_t s) {
    if (!g_profiling || ev.empty()) return;
    cudaEventRecord(ev.back().b, s);
  }
  double resolve() {  // ms; call after the streams are synchronised
    double ms = 0.0;
    for (auto& p : ev) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) ms += t;
      cudaEventDestroy(p.a); cudaEventDestroy(p.b);
    }
    ev.clear();
    return ms;
  }
};

bool host_ptr_is_pinned(const void* p) {
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost;
}

int ensure_init_locked() {
  if (g_inited && !g_devs.empty()) return B200_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fail(B200_ERR_NO_DEVICE, "no CUDA device available (%s); this build has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  return fail(B200_ERR_NO_DEVICE, "b200_init() has not been called");
}

int init_device(Device& d, int id) {
  d.id = id;
  CU(cudaSetDevice(id));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, id));
  if (prop.major < 10)
    return fail(B200_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", id, prop.major, prop.minor);
  d.numSMs = prop.multiProcessorCount;
  for (int s = 0; s < NSLOT; ++s) {
    CU(cudaStreamCreateWithFlags(&d.stream[s], cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&d.done[s], cudaEventDisableTiming));
  }
  return B200_OK;
}

// ------------------------------------------------------------------ launch helpers
template <int VAR, bool FIT, bool TRI>
int launch_pair_t(const PairArgs& a, dim3 grid, cudaStream_t st) {
  static bool attr[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr[dev & 63]) {
    CU(cudaFuncSetAttribute(pair_kernel<VAR, FIT, TRI>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_BYTES));
    attr[dev & 63] = true;
  }
  COUNT_LAUNCH();
  pair_kernel<VAR, FIT, TRI><<<grid, PAIR_THREADS, PAIR_SMEM_BYTES, st>>>(a);
  CU(cudaGetLastError());
  return B200_OK;
}

int g_variant = -1;  // MMA shape variant; env B200_MMA_VARIANT overrides (0..3)
int mma_variant() {
  if (g_variant < 0) {
    const char* e = getenv("B200_MMA_VARIANT");
    g_variant = e ? atoi(e) : 3;
    if (g_variant < 0 || g_variant > 3) g_variant = 3;
  }
  return g_variant;
}

int launch_pair(const PairArgs& a, dim3 grid, bool fit, bool tri, cudaStream_t st) {
#define DISPATCH(V)                                                         \
  case V:                                                                   \
    if (fit) return tri ? launch_pair_t<V, true, true>(a, grid, st) : launch_pair_t<V, true, false>(a, grid, st); \
    else     return tri ? launch_pair_t<V, false, true>(a, grid, st) : launch_pair_t<V, false, false>(a, grid, st);
  switch (mma_variant()) {
    DISPATCH(0)
    DISPATCH(1)
    DISPATCH(2)
    DISPATCH(3)
  }
#undef DISPATCH
  return fail(B200_ERR_ARG, "bad MMA variant");
}

struct PackSet {  // one packed frame set resident on the device
  double* planes = nullptr;
  double* G = nullptr;
  int nFrames = 0, Fpad = 0, Kpad = 0;
};

/// Pack output frames [f0, nFrames) (f0 multiple of 32) of a device-resident COORDS array.
int run_pack(const float* d_crd, size_t stride, const int* d_frameIdx, long srcBase, int nFrames, int f0,
             const int* d_atomIdx, int nAtoms, const double* d_centerMass, const double* d_covMass,
             const double* d_shift, int fit, PackSet& ps, cudaStream_t st) {
  PackArgs a;
  a.crd = d_crd; a.stride = stride; a.frameIdx = d_frameIdx; a.srcBase = srcBase; a.nFrames = nFrames;
  a.f0 = f0; a.atomIdx = d_atomIdx; a.nAtoms = nAtoms; a.Kpad = ps.Kpad; a.centerMass = d_centerMass;
  a.covMass = d_covMass; a.shift = d_shift; a.fit = fit; a.planes = ps.planes; a.G = ps.G;
  const int nrg = (ps.Fpad - f0) / ROWG;
  if (nrg <= 0) return B200_OK;
  COUNT_LAUNCH();
  pack_kernel<<<nrg, 256, 0, st>>>(a);
  CU(cudaGetLastError());
  return B200_OK;
}

/// Rows of a band launch: grid covers row groups [rg0, rg0+nRgI) x column groups [cg0, nCg).
int run_pair_band(const PackSet& A, const PackSet& B, int rg0, int nRgI, bool tri, bool fit,
                  const double* d_totalMass, float* out, size_t outBase, size_t ldo, cudaStream_t st) {
  PairArgs a;
  a.PA = A.planes; a.PB = B.planes; a.GA = A.G; a.GB = B.G;
  a.nKb = A.Kpad / KBLK; a.nRows = A.nFrames; a.nCols = B.nFrames;
  a.rg0 = rg0; a.nRgI = nRgI; a.cg0 = tri ? rg0 : 0;
  a.totalMass = d_totalMass; a.out = out; a.outBase = outBase; a.ldo = ldo;
  const int nCg = B.Fpad / ROWG;
  dim3 grid((unsigned)(nCg - a.cg0), (unsigned)nRgI);
  if (grid.x == 0 || grid.y == 0) return B200_OK;
  return launch_pair(a, grid, fit, tri, st);
}


// ------------------------------------------------------------------ pair engine selection
// 0 = auto (tcgen05 int8 path when the selection's extent allows >= the required fractional bits,
//     else FP64 DMMA), 1 = FP64 DMMA always, 2 = tcgen05 int8 or fail.  Env B200_PAIR_ENGINE.
int g_engine = -1;
int pair_engine() {
  if (g_engine < 0) {
    const char* e = getenv("B200_PAIR_ENGINE");
    g_engine = e ? atoi(e) : 0;
    if (g_engine < 0 || g_engine > 2) g_engine = 0;
  }
  return g_engine;
}
std::atomic<int> g_lastEngine{0}, g_lastQs{0};
long long* g_dbgClk = nullptr;   // device buffer for the kernel's cycle counters (timing experiments)

struct I8Set {  // one quantised frame set resident on the device
  uint8_t* image = nullptr;
  double* G = nullptr;
  double* centers = nullptr;
  int nFrames = 0, nRg = 0, nC = 0;
};

/// Worst-case RMSD change caused by rounding to a grid of spacing 2^-qs (both frames):
/// every coordinate off by half a grid step => sqrt(3)/2 * 2^-qs * sqrt(N/M) per frame.
inline double i8_worst_error(int qs, int nAtoms, double totalMass) {
  return std::ldexp(1.0, -qs) * std::sqrt(3.0 * (double)nAtoms / totalMass);
}
// Worst-case bound on the RMSD change from rounding (every one of the 6N coordinates off by half a grid step in the
// worst direction); the rest of the 1e-4 A contract (north_star) is left to the float32 store (half an ulp: 3.8e-6 A
// below 128 A) and the per-pair solve (~1e-6 A).  Measured errors are ~10x below the bound.
constexpr double I8_MAX_WORST_ERROR = 8.5e-5;

int i8_reserve(I8Set& S, DevBuf& img, DevBuf& G, DevBuf& cen, int nFrames, int nAtoms) {
  S.nFrames = nFrames;
  S.nRg = (nFrames + I8_FR_PER_RG - 1) / I8_FR_PER_RG;
  S.nRg += S.nRg & 1;   // B tiles take row groups in pairs
  S.nC = (nAtoms + I8_KC - 1) / I8_KC;
  int rc;
  if ((rc = img.reserve(i8_image_bytes(S.nRg, S.nC)))) return rc;
  if ((rc = G.reserve((size_t)S.nRg * I8_FR_PER_RG * sizeof(double)))) return rc;
  if ((rc = cen.reserve((size_t)S.nRg * I8_FR_PER_RG * 3 * sizeof(double)))) return rc;
  S.image = (uint8_t*)img.p; S.G = (double*)G.p; S.centers = (double*)cen.p;
  return B200_OK;
}

int i8_stats(const I8Set& S, const float* d_crd, size_t stride, const int* d_frameIdx, long srcBase, int f0,
             const int* d_atomIdx, int nAtoms, const double* d_centerMass, const double* d_covMass,
             unsigned int* d_maxBits, cudaStream_t st, int fEnd = -1) {
  if (fEnd < 0) fEnd = S.nFrames;
  I8StatsArgs a;
  a.crd = d_crd; a.stride = stride; a.frameIdx = d_frameIdx; a.srcBase = srcBase; a.nFrames = fEnd; a.f0 = f0;
  a.atomIdx = d_atomIdx; a.nAtoms = nAtoms; a.centerMass = d_centerMass; a.covMass = d_covMass;
  a.centers = S.centers; a.maxAbsBits = d_maxBits;
  const int nb = (fEnd - f0 + 7) / 8;
  if (nb <= 0) return B200_OK;
  COUNT_LAUNCH();
  i8_stats_kernel<<<nb, 256, 0, st>>>(a);
  CU(cudaGetLastError());
  return B200_OK;
}

/// Zero image and G (padding atoms, rows and frames must read as zeros) before the first i8_quant of a set.
int i8_clear(const I8Set& S, cudaStream_t st) {
  CU(cudaMemsetAsync(S.image, 0, i8_image_bytes(S.nRg, S.nC), st));
  CU(cudaMemsetAsync(S.G, 0, (size_t)S.nRg * I8_FR_PER_RG * sizeof(double), st));
  return B200_OK;
}

int i8_quant(const I8Set& S, const float* d_crd, size_t stride, const int* d_frameIdx, long srcBase, int f0,
             const int* d_atomIdx, int nAtoms, const double* d_covMass, int qs, cudaStream_t st, int fEnd = -1) {
  if (fEnd < 0) fEnd = S.nFrames;
  I8QuantArgs a;
  a.crd = d_crd; a.stride = stride; a.frameIdx = d_frameIdx; a.srcBase = srcBase; a.nFrames = fEnd; a.f0 = f0;
  a.atomIdx = d_atomIdx; a.nAtoms = nAtoms; a.nC = S.nC; a.covMass = d_covMass; a.centers = S.centers;
  a.scale = std::ldexp(1.0, qs); a.invScale2 = std::ldexp(1.0, -2 * qs);
  a.image = S.image; a.G = S.G;
  const int nb = (fEnd - f0 + 7) / 8;
  if (nb <= 0) return B200_OK;
  COUNT_LAUNCH();
  i8_quant_kernel<<<nb, 256, 0, st>>>(a);
  CU(cudaGetLastError());
  return B200_OK;
}

/// Reads back (max |coordinate|, total mass) after the stats kernel and picks the number of
/// fractional bits.  *eligible = false when the grid would be too coarse for the contract.
int i8_choose_scale(Device& d, const unsigned int* d_maxBits, const double* d_total, int nAtoms, cudaStream_t st,
                    int* qs, bool* eligible) {
  int rc;
  if ((rc = d.hostScal.reserve(64))) return rc;
  unsigned int* hBits = (unsigned int*)d.hostScal.p;
  double* hTotal = (double*)((char*)d.hostScal.p + 8);
  CU(cudaMemcpyAsync(hBits, d_maxBits, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(hTotal, d_total, sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  float mx;
  std::memcpy(&mx, hBits, 4);
  const double total = *hTotal;
  *eligible = false; *qs = 0;
  if (!(total > 0.0) || !std::isfinite(mx)) return B200_OK;
  int q = 30;
  if (mx > 0.f) q = (int)std::floor(std::log2((double)I8_QMAX / (double)mx));
  if (q > 30) q = 30;
  if (q < 0) return B200_OK;
  *qs = q;
  *eligible = i8_worst_error(q, nAtoms, total) <= I8_MAX_WORST_ERROR;
  return B200_OK;
}

// MMA CTA group of the tcgen05 kernel: 2 = CTA pairs (tcgen05.mma.cta_group::2, 28 x 28 tiles), 1 = single CTA
// (14 x 28 tiles).  Env B200_I8_CTA_GROUP or b200_set_i8_cta_group().
int g_i8Cg = -1;
int i8_cta_group() {
  if (g_i8Cg < 0) {
    const char* e = getenv("B200_I8_CTA_GROUP");
    g_i8Cg = e ? atoi(e) : 2;
    if (g_i8Cg != 1 && g_i8Cg != 2) g_i8Cg = 2;
  }
  return g_i8Cg;
}

template <bool TRI, int CG, bool DBG>
int launch_pair_i8_t(const PairI8Args& a, int grid, cudaStream_t st) {
  static bool attr[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  constexpr int smem = i8_smem_bytes<CG>();
  if (!attr[dev & 63]) {
    CU(cudaFuncSetAttribute(pair_i8_kernel<TRI, CG, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr[dev & 63] = true;
  }
  COUNT_LAUNCH();
  if (CG == 1) {
    pair_i8_kernel<TRI, CG, DBG><<<grid, I8_THREADS, smem, st>>>(a);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(I8_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CU(cudaLaunchKernelEx(&cfg, pair_i8_kernel<TRI, CG, DBG>, a));
  }
  CU(cudaGetLastError());
  return B200_OK;
}

template <int CG>
int run_pair_i8_band_t(const Device& d, PairI8Args& a, int rowLo, int rowHi, bool tri, cudaStream_t st) {
  a.it0 = rowLo / (I8_TILE_I * CG); a.it1 = (rowHi + I8_TILE_I * CG - 1) / (I8_TILE_I * CG);
  if (a.it1 <= a.it0 || a.jt1 <= a.jt0) return B200_OK;
  const long nTiles = tri ? i8_count_tiles<true, CG>(a.it0, a.it1, a.jt0, a.jt1) : i8_count_tiles<false, CG>(a.it0, a.it1, a.jt0, a.jt1);
  if (nTiles <= 0) return B200_OK;
  if (nTiles > (1L << 30)) return fail(B200_ERR_ARG, "pair matrix too large for one launch (%ld tiles)", nTiles);
  const int groups = (d.numSMs > 0 ? d.numSMs : 148) / CG;   // persistent: one CTA per SM, CG CTAs per tile
  // chunks of consecutive tiles of the list: every group gets k chunks of equal length (env B200_I8_CHUNK = target
  // chunk length; 1 = tile t to group t % groups)
  long L = 1;
  {
    const char* e = getenv("B200_I8_CHUNK");
    const long target = e ? std::max(1, atoi(e)) : 1;
    if (target > 1) {
      const long k = std::max<long>(1, (nTiles + groups * target / 2) / (groups * target));
      L = (nTiles + groups * k - 1) / (groups * k);
    }
  }
  a.chunkLen = (int)L;
  const long nChunks = (nTiles + L - 1) / L;
  const int grid = (int)std::min<long>(nChunks, groups) * CG;
  const bool dbg = a.dbgMode != 0 || a.dbgClk != nullptr || a.dbgS != nullptr;
  if (dbg) return tri ? launch_pair_i8_t<true, CG, true>(a, grid, st) : launch_pair_i8_t<false, CG, true>(a, grid, st);
  return tri ? launch_pair_i8_t<true, CG, false>(a, grid, st) : launch_pair_i8_t<false, CG, false>(a, grid, st);
}

/// Rows [rowLo,rowHi) of the pair matrix through the tcgen05 int8 kernel (persistent, one CTA per SM).
int run_pair_i8_band(const Device& d, const I8Set& A, const I8Set& B, int rowLo, int rowHi, bool tri, int qs,
                     const double* d_totalMass, float* out, size_t outBase, size_t ldo, double* dbgS, cudaStream_t st) {
  if (rowHi <= rowLo) return B200_OK;
  PairI8Args a;
  a.PA = A.image; a.PB = B.image; a.GA = A.G; a.GB = B.G; a.nC = A.nC;
  a.nRows = A.nFrames; a.nCols = B.nFrames; a.rowLo = rowLo; a.rowHi = rowHi;
  a.jt0 = tri ? (rowLo + 1) / I8_TILE_J : 0;
  a.jt1 = (B.nFrames + I8_TILE_J - 1) / I8_TILE_J;
  a.totalMass = d_totalMass; a.invScale2 = std::ldexp(1.0, -2 * qs);
  a.out = out; a.outBase = outBase; a.ldo = ldo; a.dbgS = dbgS;
  { const char* e = getenv("B200_I8_DEBUG_MODE"); a.dbgMode = e ? atoi(e) : 0; }
  a.dbgClk = g_dbgClk;
  return i8_cta_group() == 2 ? run_pair_i8_band_t<2>(d, a, rowLo, rowHi, tri, st) : run_pair_i8_band_t<1>(d, a, rowLo, rowHi, tri, st);
}

int shard_rows(int nFrames, int rank, int count, int* row0, int* row1) {
  if (nFrames < 0 || count < 1 || rank < 0 || rank >= count) return fail(B200_ERR_ARG, "bad shard %d/%d", rank, count);
  // boundary b_s: smallest 32-aligned row with area(rows < b_s) >= s/count * total
  const double F = (double)nFrames;
  const double total = F * (F - 1.0) / 2.0;
  auto bound = [&](int s) -> int {
    if (s <= 0) return 0;
    if (s >= count) return nFrames;
    const double target = total * (double)s / (double)count;
    // area(r) = r*F - r*(r+1)/2  => r = ((2F-1) - sqrt((2F-1)^2 - 8 target))/2
    const double bq = 2.0 * F - 1.0;
    double r = (bq - std::sqrt(std::max(0.0, bq * bq - 8.0 * target))) / 2.0;
    long ri = (long)std::llround(r / ROWG) * ROWG;
    if (ri < 0) ri = 0;
    if (ri > nFrames) ri = nFrames;
    return (int)ri;
  };
  int a = bound(rank), b = bound(rank + 1);
  if (b < a) b = a;
  *row0 = a; *row1 = b;
  return B200_OK;
}

// ------------------------------------------------------------------ triangle: plan + bands
struct TriPlan {   // what prepare_tri() left on the device for the band launches
  bool i8 = false;
  int qs = 0;
  PackSet ps;      // FP64 planes   (engine FP64)
  I8Set q;         // int8 images   (engine tcgen05)
  double* d_total = nullptr;
};

/// Centre + pack output frames >= row0's tile for the engine that will run the pairs.
/// Blocks the host once (a few scalars read back) when the tcgen05 engine is a candidate.
int prepare_tri(Device& d, const float* d_crd, size_t stride, const int* d_frameIdx, long srcBase, int nFrames,
                const int* d_atomIdx, int nAtoms, const double* d_mass, int fit, int row0, cudaStream_t st,
                Timer* tpack, TriPlan& plan) {
  int rc;
  if ((rc = d.scal.reserve(64))) return rc;
  double* d_total = (double*)d.scal.p;
  double* d_shift = d_total + 1;
  unsigned int* d_maxBits = (unsigned int*)(d_total + 4);
  plan.d_total = d_total;
  if (tpack) tpack->begin(st);
  COUNT_LAUNCH();
  mass_sum_kernel<<<1, 32, 0, st>>>(d_mass, nAtoms, d_total);
  const int engine = pair_engine();
  if (fit && engine != 1) {
    if ((rc = i8_reserve(plan.q, d.imgA, d.GA, d.cenA, nFrames, nAtoms))) return rc;
    const int f0 = (row0 / I8_FR_PER_RG) * I8_FR_PER_RG;
    CU(cudaMemsetAsync(d_maxBits, 0, sizeof(unsigned int), st));
    if ((rc = i8_stats(plan.q, d_crd, stride, d_frameIdx, srcBase, f0, d_atomIdx, nAtoms, d_mass, d_mass, d_maxBits, st))) return rc;
    bool ok = false;
    if ((rc = i8_choose_scale(d, d_maxBits, d_total, nAtoms, st, &plan.qs, &ok))) return rc;
    if (ok) {
      if ((rc = i8_clear(plan.q, st))) return rc;
      if ((rc = i8_quant(plan.q, d_crd, stride, d_frameIdx, srcBase, f0, d_atomIdx, nAtoms, d_mass, plan.qs, st))) return rc;
      plan.i8 = true;
    } else if (engine == 2) {
      return fail(B200_ERR_ARG, "tcgen05 int8 engine forced but the selection's extent leaves only %d fractional bits", plan.qs);
    }
  } else if (engine == 2) {
    return fail(B200_ERR_ARG, "tcgen05 int8 engine forced but nofit RMSD runs on the FP64 engine only");
  }
  if (!plan.i8) {
    PackSet& ps = plan.ps;
    ps.nFrames = nFrames; ps.Fpad = round_up(nFrames, ROWG); ps.Kpad = round_up(nAtoms, KC);
    if ((rc = d.planesA.reserve(plane_doubles(ps.Fpad, ps.Kpad) * sizeof(double)))) return rc;
    if ((rc = d.GA.reserve((size_t)ps.Fpad * sizeof(double)))) return rc;
    ps.planes = (double*)d.planesA.p; ps.G = (double*)d.GA.p;
    // rows >= row0 only pair with columns > row0: frames below row0's row group are never read
    const int f0 = (row0 / ROWG) * ROWG;
    if (!fit) { COUNT_LAUNCH(); shift_kernel<<<1, 1, 0, st>>>(d_crd, stride, d_frameIdx, srcBase, f0, d_atomIdx, d_shift); }
    if ((rc = run_pack(d_crd, stride, d_frameIdx, srcBase, nFrames, f0, d_atomIdx, nAtoms, d_mass, d_mass, d_shift,
                       fit, ps, st))) return rc;
  }
  if (tpack) tpack->end(st);
  g_lastEngine.store(plan.i8 ? 2 : 1); g_lastQs.store(plan.i8 ? plan.qs : 0);
  return B200_OK;
}

/// Rows [i0,i1) of the triangle; `out` indexed as out[triIndex - outBase].
int run_tri_band(Device& d, const TriPlan& plan, int i0, int i1, bool fit, float* out, size_t outBase, cudaStream_t st) {
  if (plan.i8) return run_pair_i8_band(d, plan.q, plan.q, i0, i1, true, plan.qs, plan.d_total, out, outBase, 0, nullptr, st);
  // the FP64 kernel works on whole 32-row groups: neighbouring bands must be 32-aligned
  return run_pair_band(plan.ps, plan.ps, i0 / ROWG, (i1 + ROWG - 1) / ROWG - i0 / ROWG, true, fit, plan.d_total, out,
                       outBase, 0, st);
}

int dev_rms2d_tri(Device& d, const float* d_crd, size_t stride, const int* d_frameIdx, long srcBase, int nFrames,
                  const int* d_atomIdx, int nAtoms, const double* d_mass, int fit, int row0, int row1,
                  float* d_out, size_t outBase, cudaStream_t st, int bandRows, Timer* tpack, Timer* tpair) {
  if (nFrames < 2 || row1 <= row0) return B200_OK;
  TriPlan plan;
  int rc;
  if ((rc = prepare_tri(d, d_crd, stride, d_frameIdx, srcBase, nFrames, d_atomIdx, nAtoms, d_mass, fit, row0, st, tpack, plan)))
    return rc;
  if (bandRows <= 0) bandRows = 512;
  // tcgen05 engine: size the band so that its row operands (9 bytes per atom and frame, padded) stay in L2 while the
  // columns stream: ~38 MB (4096 rows at 1,000 atoms; measured against 2048 / 8192 / one launch per shard)
  if (plan.i8) {
    const double bytesPerFrame = (double)plan.q.nC * I8_BLK_BYTES / I8_FR_PER_RG;
    const char* e = getenv("B200_I8_BAND_MB");
    const double mb = e ? atof(e) : 38.4;
    bandRows = std::max(512, (int)(mb * 1e6 / bytesPerFrame) / 512 * 512);
  }
  // (one launch per shard was measured slower than 4096-row bands for the tcgen05 kernel: the row operands of a band,
  //  ~37 MB at 1000 atoms, stay in L2 while the columns stream)
  const int tileRows = I8_TILE_I * i8_cta_group();
  const int groups = std::max(1, (d.numSMs > 0 ? d.numSMs : 148) / i8_cta_group());
  for (long i0l = row0; i0l < row1; ) {
    const int i0 = (int)i0l;
    int i1 = (int)std::min<long>(row1, i0l + bandRows);
    // A band whose number of row tiles is a multiple of the CTA-group count deals every group the same row tiles in
    // every column: all groups then change columns in lockstep and hit the same L2 lines of the new column operand at
    // the same instant (measured: 25 % slower on two of eight cfg5 shards).  One tile row less breaks the lockstep.
    if (plan.i8 && i1 - i0 > 2 * tileRows && ((i1 + tileRows - 1) / tileRows - i0 / tileRows) % groups == 0) i1 -= tileRows;
    i0l = i1;
    if (tpair) tpair->begin(st);
    if ((rc = run_tri_band(d, plan, i0, i1, fit != 0, d_out, outBase, st))) return rc;
    if (tpair) tpair->end(st);
  }
  return B200_OK;
}

void add_stats(double packMs, long packN, double pairMs, long pairN, double pairs, double h2d, double d2h) {
  std::lock_guard<std::mutex> lk(g_statMu);
  g_stats.pack_ms += packMs; g_stats.pack_launches += packN;
  g_stats.pair_ms += pairMs; g_stats.pair_launches += pairN;
  g_stats.pairs += pairs; g_stats.h2d_bytes += h2d; g_stats.d2h_bytes += d2h;
}

// Upload the needed span of host COORDS (frames [fLo,fHi), floats [0,width)) to d.crd / d.crdB.
int upload_crd(DevBuf& buf, const float* crd, size_t stride, int fLo, int fHi, size_t widthFloats, cudaStream_t st,
               double* h2dBytes) {
  const size_t rows = (size_t)(fHi - fLo);
  int rc;
  if ((rc = buf.reserve(rows * widthFloats * sizeof(float)))) return rc;
  CU(cudaMemcpy2DAsync(buf.p, widthFloats * sizeof(float), crd + (size_t)fLo * stride, stride * sizeof(float),
                       widthFloats * sizeof(float), rows, cudaMemcpyHostToDevice, st));
  *h2dBytes += (double)(rows * widthFloats * sizeof(float));
  return B200_OK;
}

template <typename T>
int upload_vec(DevBuf& buf, const T* host, size_t n, cudaStream_t st) {
  int rc;
  if ((rc = buf.reserve(std::max<size_t>(n, 1) * sizeof(T)))) return rc;
  if (n) CU(cudaMemcpyAsync(buf.p, host, n * sizeof(T), cudaMemcpyHostToDevice, st));
  return B200_OK;
}

int validate_sel(const int* atomIdx, int nAtoms, size_t stride, int* maxAtom) {
  if (!atomIdx || nAtoms <= 0) return fail(B200_ERR_ARG, "no atoms selected");
  int mx = 0;
  for (int k = 0; k < nAtoms; ++k) {
    if (atomIdx[k] < 0) return fail(B200_ERR_ARG, "negative atom index");
    mx = std::max(mx, atomIdx[k]);
  }
  if ((size_t)3 * ((size_t)mx + 1) > stride) return fail(B200_ERR_ARG, "atom index %d outside frame stride %zu", mx, stride);
  *maxAtom = mx;
  return B200_OK;
}

// Pipelined host path of the tcgen05 engine (pinned COORDS, no frame list).  Frames are uploaded from the END of the
// trajectory towards the start, chunk by chunk.  Rows [a, b) of the triangle pair only with frames >= a, so as soon as a
// chunk has landed and is quantised its band of rows can be computed and its slice of the triangle copied back while the
// next chunk uploads: H2D, compute and D2H overlap (PCIe is full duplex) instead of upload -> compute/download.
// The fixed-point scale must be known before the first chunk is quantised: it is taken from the first chunk's extent
// with 25 % headroom, every later chunk keeps feeding the running maximum, and if at the end the headroom turned out
// too small (or the scale leaves too few fractional bits) *done stays false and the caller runs the two-pass path.
int host_tri_pipelined_i8(Device& d, const float* crd, size_t stride, int nFrames, const int* atomIdx, int nAtoms,
                          const double* mass, int row0, int row1, float* outTri, bool* done) {
  *done = false;
  int maxAtom = 0, rc;
  if ((rc = validate_sel(atomIdx, nAtoms, stride, &maxAtom))) return rc;
  const size_t F = (size_t)nFrames;
  const size_t width = (size_t)3 * ((size_t)maxAtom + 1);
  const int f0 = (row0 / I8_TILE_J) * I8_TILE_J;           // first frame needed (tile aligned)
  // chunk = band: <= 64 MB of output, a multiple of the 28-frame tile
  int C = (int)std::min<size_t>(2016, std::max<size_t>(8 * I8_TILE_J, ((size_t)40 << 20) / (4 * F) / I8_TILE_J * I8_TILE_J));
  { const char* e = getenv("B200_PIPE_ROWS"); if (e && atoi(e) >= I8_TILE_J) C = atoi(e) / I8_TILE_J * I8_TILE_J; }
  cudaStream_t sIn = d.stream[0];
  if ((rc = d.crd.reserve((size_t)(nFrames - f0) * width * sizeof(float)))) return rc;
  if ((rc = upload_vec(d.idxA, atomIdx, (size_t)nAtoms, sIn))) return rc;
  if (mass && (rc = upload_vec(d.massA, mass, (size_t)nAtoms, sIn))) return rc;
  const double* d_mass = mass ? (const double*)d.massA.p : nullptr;
  if ((rc = d.scal.reserve(64))) return rc;
  double* d_total = (double*)d.scal.p;
  unsigned int* d_maxBits = (unsigned int*)(d_total + 4);
  COUNT_LAUNCH();
  mass_sum_kernel<<<1, 32, 0, sIn>>>(d_mass, nAtoms, d_total);
  CU(cudaMemsetAsync(d_maxBits, 0, sizeof(unsigned int), sIn));
  I8Set q;
  if ((rc = i8_reserve(q, d.imgA, d.GA, d.cenA, nFrames, nAtoms))) return rc;
  if ((rc = i8_clear(q, sIn))) return rc;
  // chunk list, last chunk of the trajectory first.  The first chunk to be uploaded (top of the trajectory) and the
  // last bands to be computed (lowest rows = longest rows) are smaller: nothing overlaps the first upload and the
  // last download.
  struct Chunk { int fa, fb; };
  std::vector<Chunk> chunks;
  {
    std::vector<int> bnd;   // ascending boundaries from f0
    const int small = std::max(I8_TILE_J, C / 4 / I8_TILE_J * I8_TILE_J);
    int at = f0, step = small;
    while (at < nFrames) {
      bnd.push_back(at);
      at += step;
      if (step < C) step = std::min(C, step + small);
    }
    // split a small piece off the top
    if (nFrames - bnd.back() > 2 * small) bnd.push_back(std::max(bnd.back() + small, (nFrames - small) / I8_TILE_J * I8_TILE_J));
    bnd.push_back(nFrames);
    for (size_t i = bnd.size() - 1; i > 0; --i)
      if (bnd[i] > bnd[i - 1]) chunks.push_back({bnd[i - 1], bnd[i]});
  }
  size_t maxChunk = 0;
  for (const Chunk& c : chunks) {
    const int lo = std::max(c.fa, row0), hi = std::min(c.fb, row1);
    if (hi > lo) maxChunk = std::max(maxChunk, tri_row_start(F, hi) - tri_row_start(F, lo));
  }
  const bool pinnedOut = host_ptr_is_pinned(outTri + tri_row_start(F, (size_t)row0));   // (the shard's own range: the base may lie outside the caller's buffer)
  for (int s = 1; s < NSLOT; ++s) {
    if ((rc = d.outChunk[s].reserve(std::max<size_t>(maxChunk, 1) * sizeof(float)))) return rc;
    if (!pinnedOut && (rc = d.outStage[s].reserve(std::max<size_t>(maxChunk, 1) * sizeof(float)))) return rc;
  }
  struct Pending { size_t base = 0, n = 0; bool live = false; } pend[NSLOT];
  auto retire = [&](int s) -> int {
    if (!pend[s].live) return B200_OK;
    CU(cudaEventSynchronize(d.done[s]));
    if (!pinnedOut) std::memcpy(outTri + pend[s].base, d.outStage[s].p, pend[s].n * sizeof(float));
    pend[s].live = false;
    return B200_OK;
  };
  std::vector<cudaEvent_t> evs;
  auto cleanup = [&]() { for (cudaEvent_t e : evs) cudaEventDestroy(e); evs.clear(); };
  const float* d_crd = (const float*)d.crd.p;
  const long srcBase = (long)f0;      // row r of the device copy is frame f0 + r
  Timer tpair;
  double h2d = (double)nAtoms * 4 + (mass ? (double)nAtoms * 8 : 0), d2h = 0.0;
  int qs = 0, band = 0;
  long nLaunch = 0;
  for (size_t k = 0; k < chunks.size(); ++k) {
    const int fa = chunks[k].fa, fb = chunks[k].fb;
    CU(cudaMemcpy2DAsync((float*)d.crd.p + (size_t)(fa - f0) * width, width * sizeof(float), crd + (size_t)fa * stride,
                         stride * sizeof(float), width * sizeof(float), (size_t)(fb - fa), cudaMemcpyHostToDevice, sIn));
    h2d += (double)(fb - fa) * width * sizeof(float);
    if ((rc = i8_stats(q, d_crd, width, nullptr, srcBase, fa, (const int*)d.idxA.p, nAtoms, d_mass, d_mass, d_maxBits, sIn, fb))) { cleanup(); return rc; }
    if (k == 0) {
      // scale from the first chunk, with headroom for the frames not seen yet
      if ((rc = d.hostScal.reserve(64))) { cleanup(); return rc; }
      unsigned int* hBits = (unsigned int*)d.hostScal.p;
      double* hTotal = (double*)((char*)d.hostScal.p + 8);
      CU(cudaMemcpyAsync(hBits, d_maxBits, sizeof(unsigned int), cudaMemcpyDeviceToHost, sIn));
      CU(cudaMemcpyAsync(hTotal, d_total, sizeof(double), cudaMemcpyDeviceToHost, sIn));
      CU(cudaStreamSynchronize(sIn));
      float mx;
      std::memcpy(&mx, hBits, 4);
      const double total = *hTotal;
      if (!(total > 0.0) || !std::isfinite(mx) || !(mx > 0.f)) { cleanup(); return B200_OK; }
      qs = std::min(30, (int)std::floor(std::log2((double)I8_QMAX / (1.25 * (double)mx))));
      if (qs < 0 || i8_worst_error(qs, nAtoms, total) > I8_MAX_WORST_ERROR) { cleanup(); return B200_OK; }
    }
    if ((rc = i8_quant(q, d_crd, width, nullptr, srcBase, fa, (const int*)d.idxA.p, nAtoms, d_mass, qs, sIn, fb))) { cleanup(); return rc; }
    const int lo = std::max(fa, row0), hi = std::min(fb, row1);
    if (hi <= lo) continue;           // frames above this shard's rows: columns only
    cudaEvent_t ev;
    CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    evs.push_back(ev);
    CU(cudaEventRecord(ev, sIn));
    const int s = 1 + (band++ % (NSLOT - 1));
    if ((rc = retire(s))) { cleanup(); return rc; }
    cudaStream_t st = d.stream[s];
    CU(cudaStreamWaitEvent(st, ev, 0));
    const size_t base = tri_row_start(F, lo), n = tri_row_start(F, hi) - base;
    tpair.begin(st);
    if ((rc = run_pair_i8_band(d, q, q, lo, hi, true, qs, d_total, (float*)d.outChunk[s].p, base, 0, nullptr, st))) { cleanup(); return rc; }
    tpair.end(st);
    ++nLaunch;
    if (n) {
      float* dst = pinnedOut ? outTri + base : (float*)d.outStage[s].p;
      CU(cudaMemcpyAsync(dst, d.outChunk[s].p, n * sizeof(float), cudaMemcpyDeviceToHost, st));
      d2h += (double)n * sizeof(float);
    }
    CU(cudaEventRecord(d.done[s], st));
    pend[s].base = base; pend[s].n = n; pend[s].live = true;
  }
  // did the headroom hold?
  {
    unsigned int* hBits = (unsigned int*)d.hostScal.p;
    CU(cudaMemcpyAsync(hBits, d_maxBits, sizeof(unsigned int), cudaMemcpyDeviceToHost, sIn));
    CU(cudaStreamSynchronize(sIn));
    for (int s = 1; s < NSLOT; ++s) if ((rc = retire(s))) { cleanup(); return rc; }
    for (int s = 1; s < NSLOT; ++s) CU(cudaStreamSynchronize(d.stream[s]));
    cleanup();
    float mx;
    std::memcpy(&mx, hBits, 4);
    if (!((double)mx * std::ldexp(1.0, qs) <= (double)I8_QMAX)) return B200_OK;   // *done == false: two-pass path recomputes
  }
  const double pairs = (double)(tri_row_start(F, row1) - tri_row_start(F, row0));
  add_stats(0.0, 1, tpair.resolve(), nLaunch, pairs, h2d, d2h);
  g_lastEngine.store(2); g_lastQs.store(qs);
  *done = true;
  return B200_OK;
}

// One shard of the triangle on one device, host buffers.
int host_tri_on_device(Device& d, const float* crd, size_t stride, int nFramesTotal, const int* frameIdx, int nFrames,
                       const int* atomIdx, int nAtoms, const double* mass, int fit, int row0, int row1, float* outTri) {
  CU(cudaSetDevice(d.id));
  if (row1 <= row0 || nFrames < 2) return B200_OK;
  int maxAtom = 0, rc;
  if ((rc = validate_sel(atomIdx, nAtoms, stride, &maxAtom))) return rc;
  {
    const char* e = getenv("B200_HOST_PIPELINE");
    if (fit && !frameIdx && pair_engine() != 1 && nFrames >= 1024 && nFrames <= nFramesTotal && !(e && atoi(e) == 0) &&
        host_ptr_is_pinned(crd)) {
      bool done = false;
      if ((rc = host_tri_pipelined_i8(d, crd, stride, nFrames, atomIdx, nAtoms, mass, row0, row1, outTri, &done))) return rc;
      if (done) return B200_OK;
    }
  }
  // source frame range needed by output frames [f0, nFrames) (f0: first frame either engine packs)
  const int f0 = std::min((row0 / ROWG) * ROWG, (row0 / I8_FR_PER_RG) * I8_FR_PER_RG);
  int sLo = f0, sHi = nFrames;
  if (frameIdx) {
    sLo = nFramesTotal; sHi = 0;
    for (int f = f0; f < nFrames; ++f) {
      if (frameIdx[f] < 0 || frameIdx[f] >= nFramesTotal) return fail(B200_ERR_ARG, "frameIdx[%d]=%d out of range", f, frameIdx[f]);
      sLo = std::min(sLo, frameIdx[f]); sHi = std::max(sHi, frameIdx[f] + 1);
    }
  } else if (nFrames > nFramesTotal) {
    return fail(B200_ERR_ARG, "nFrames %d > nFramesTotal %d", nFrames, nFramesTotal);
  }
  cudaStream_t st0 = d.stream[0];
  double h2d = 0.0, d2h = 0.0;
  const size_t width = (size_t)3 * ((size_t)maxAtom + 1);
  if ((rc = upload_crd(d.crd, crd, stride, sLo, sHi, width, st0, &h2d))) return rc;
  if ((rc = upload_vec(d.idxA, atomIdx, (size_t)nAtoms, st0))) return rc;
  if (mass && (rc = upload_vec(d.massA, mass, (size_t)nAtoms, st0))) return rc;
  if (frameIdx && (rc = upload_vec(d.frameIdx, frameIdx, (size_t)nFrames, st0))) return rc;
  h2d += (double)nAtoms * 4 + (mass ? (double)nAtoms * 8 : 0) + (frameIdx ? (double)nFrames * 4 : 0);

  Timer tpack, tpair;
  TriPlan plan;
  const double* d_mass = mass ? (const double*)d.massA.p : nullptr;
  const int* d_fidx = frameIdx ? (const int*)d.frameIdx.p : nullptr;
  // pack on stream 0, then bands round-robin over the slots
  if ((rc = prepare_tri(d, (const float*)d.crd.p, width, d_fidx, (long)sLo, nFrames, (const int*)d.idxA.p, nAtoms, d_mass,
                        fit, row0, st0, &tpack, plan))) return rc;
  CU(cudaEventRecord(d.done[0], st0));
  for (int s = 1; s < NSLOT; ++s) CU(cudaStreamWaitEvent(d.stream[s], d.done[0], 0));

  // band size: <= ~64 MB of output per band, multiple of 32 rows
  const size_t F = (size_t)nFrames;
  int bandRows = (int)std::min<size_t>(2048, std::max<size_t>(ROWG, ((size_t)64 << 20) / (4 * F) / ROWG * ROWG));
  const bool pinnedOut = host_ptr_is_pinned(outTri + tri_row_start(F, (size_t)row0));   // (the shard's own range: the base may lie outside the caller's buffer)
  size_t maxChunk = 0;
  for (int i0 = row0; i0 < row1; i0 += bandRows) {
    const int i1 = std::min(row1, i0 + bandRows);
    maxChunk = std::max(maxChunk, tri_row_start(F, i1) - tri_row_start(F, i0));
  }
  for (int s = 0; s < NSLOT; ++s) {
    if ((rc = d.outChunk[s].reserve(std::max<size_t>(maxChunk, 1) * sizeof(float)))) return rc;
    if (!pinnedOut && (rc = d.outStage[s].reserve(std::max<size_t>(maxChunk, 1) * sizeof(float)))) return rc;
  }
  struct Pending { size_t base = 0, n = 0; bool live = false; } pend[NSLOT];
  auto retire = [&](int s) -> int {
    if (!pend[s].live) return B200_OK;
    CU(cudaEventSynchronize(d.done[s]));
    if (!pinnedOut) std::memcpy(outTri + pend[s].base, d.outStage[s].p, pend[s].n * sizeof(float));
    pend[s].live = false;
    return B200_OK;
  };
  int band = 0;
  long nLaunch = 0;
  for (int i0 = row0; i0 < row1; i0 += bandRows, ++band) {
    const int s = band % NSLOT;
    if ((rc = retire(s))) return rc;
    const int i1 = std::min(row1, i0 + bandRows);
    const size_t base = tri_row_start(F, i0), n = tri_row_start(F, i1) - base;
    cudaStream_t st = d.stream[s];
    tpair.begin(st);
    if ((rc = run_tri_band(d, plan, i0, i1, fit != 0, (float*)d.outChunk[s].p, base, st))) return rc;
    tpair.end(st);
    ++nLaunch;
    if (n) {
      float* dst = pinnedOut ? outTri + base : (float*)d.outStage[s].p;
      CU(cudaMemcpyAsync(dst, d.outChunk[s].p, n * sizeof(float), cudaMemcpyDeviceToHost, st));
      d2h += (double)n * sizeof(float);
    }
    CU(cudaEventRecord(d.done[s], st));
    pend[s].base = base; pend[s].n = n; pend[s].live = true;
  }
  for (int s = 0; s < NSLOT; ++s) if ((rc = retire(s))) return rc;
  for (int s = 0; s < NSLOT; ++s) CU(cudaStreamSynchronize(d.stream[s]));
  const double pairs = (double)(tri_row_start(F, row1) - tri_row_start(F, row0));
  add_stats(tpack.resolve(), 1, tpair.resolve(), nLaunch, pairs, h2d, d2h);
  return B200_OK;
}

}  // namespace

// =================================================================== C ABI
extern "C" {

int b200_version(void) { return 100; }

const char* b200_last_error(void) {
  std::lock_guard<std::mutex> lk(g_errMu);
  static thread_local std::string copy;
  copy = g_err;
  return copy.c_str();
}

static int init_ids_locked(const int* ids, int want) {
  bool same = g_inited && (int)g_devs.size() == want;
  for (int i = 0; same && i < want; ++i) same = (g_devs[i].id == ids[i]);
  if (same) return B200_OK;
  for (auto& d : g_devs) d.destroy();
  g_devs.clear();
  g_devs.resize(want);
  for (int i = 0; i < want; ++i) {
    int rc = init_device(g_devs[i], ids[i]);
    if (rc) { for (auto& d : g_devs) d.destroy(); g_devs.clear(); g_inited = false; return rc; }
  }
  cudaSetDevice(g_devs[0].id);
  g_inited = true;
  return B200_OK;
}

int b200_init(int ngpu_requested, int* ngpu_used) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (ngpu_used) *ngpu_used = 0;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fail(B200_ERR_NO_DEVICE, "no CUDA device available (%s); this build has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  const int want = (ngpu_requested <= 0) ? n : std::min(n, ngpu_requested);
  std::vector<int> ids(want);
  for (int i = 0; i < want; ++i) ids[i] = i;
  int rc = init_ids_locked(ids.data(), want);
  if (!rc && ngpu_used) *ngpu_used = want;
  return rc;
}

int b200_init_devices(const int* deviceIds, int n) {
  if (!deviceIds || n <= 0) return fail(B200_ERR_ARG, "empty device list");
  std::lock_guard<std::mutex> lk(g_mu);
  int cnt = 0;
  cudaError_t e = cudaGetDeviceCount(&cnt);
  if (e != cudaSuccess || cnt <= 0) {
    cudaGetLastError();
    return fail(B200_ERR_NO_DEVICE, "no CUDA device available (%s); this build has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  for (int i = 0; i < n; ++i)
    if (deviceIds[i] < 0 || deviceIds[i] >= cnt) return fail(B200_ERR_ARG, "device %d not present (count %d)", deviceIds[i], cnt);
  return init_ids_locked(deviceIds, n);
}

void b200_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& d : g_devs) d.destroy();
  g_devs.clear();
  g_inited = false;
}

int b200_num_devices(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  return (int)g_devs.size();
}

int b200_shard_rows(int nFrames, int shardRank, int shardCount, int* row0, int* row1) {
  if (!row0 || !row1) return fail(B200_ERR_ARG, "null output");
  return shard_rows(nFrames, shardRank, shardCount, row0, row1);
}

void b200_set_profiling(int on) { g_profiling = (on != 0); }
int b200_set_mma_variant(int v) { if (v < 0 || v > 3) return fail(B200_ERR_ARG, "variant must be 0..3"); g_variant = v; return B200_OK; }
void b200_reset_stats(void) { std::lock_guard<std::mutex> lk(g_statMu); g_stats = b200_stats(); g_launches.store(0); }
void b200_get_stats(b200_stats* out) { if (!out) return; std::lock_guard<std::mutex> lk(g_statMu); *out = g_stats; out->kernel_launches = g_launches.load(); }

int b200_rms2d_tri_shard(const float* crd, size_t frameStrideFloats, int nFramesTotal, const int* frameIdx, int nFrames,
                         const int* atomIdx, int nAtoms, const double* mass, int fit, int shardRank, int shardCount,
                         float* outTri, size_t* firstElt, size_t* nElts) {
  int r0 = 0, r1 = 0, rc;
  if ((rc = shard_rows(nFrames, shardRank, shardCount, &r0, &r1))) return rc;
  const size_t F = (size_t)std::max(nFrames, 0);
  if (firstElt) *firstElt = nFrames > 1 ? tri_row_start(F, r0) : 0;
  if (nElts) *nElts = nFrames > 1 ? tri_row_start(F, r1) - tri_row_start(F, r0) : 0;
  if (!outTri) return B200_OK;
  if (!crd) return fail(B200_ERR_ARG, "null coordinates");
  std::lock_guard<std::mutex> lk(g_mu);
  if ((rc = ensure_init_locked())) return rc;
  return host_tri_on_device(g_devs[0], crd, frameStrideFloats, nFramesTotal, frameIdx, nFrames, atomIdx, nAtoms, mass, fit,
                            r0, r1, outTri);
}

int b200_rms2d_tri(const float* crd, size_t frameStrideFloats, int nFramesTotal, const int* frameIdx, int nFrames,
                   const int* atomIdx, int nAtoms, const double* mass, int fit, float* outTri) {
  if (!crd || !outTri) return fail(B200_ERR_ARG, "null buffer");
  std::lock_guard<std::mutex> lk(g_mu);
  int rc;
  if ((rc = ensure_init_locked())) return rc;
  const int nd = (int)g_devs.size();
  if (nd == 1)
    return host_tri_on_device(g_devs[0], crd, frameStrideFloats, nFramesTotal, frameIdx, nFrames, atomIdx, nAtoms, mass, fit,
                              0, nFrames, outTri);
  // one host thread per device; shards are disjoint contiguous ranges of outTri
  std::vector<int> rcs(nd, 0);
  std::vector<std::thread> th;
  for (int i = 0; i < nd; ++i) {
    th.emplace_back([&, i]() {
      int r0 = 0, r1 = 0;
      int r = shard_rows(nFrames, i, nd, &r0, &r1);
      if (!r) r = host_tri_on_device(g_devs[i], crd, frameStrideFloats, nFramesTotal, frameIdx, nFrames, atomIdx, nAtoms,
                                     mass, fit, r0, r1, outTri);
      rcs[i] = r;
    });
  }
  for (auto& t : th) t.join();
  cudaSetDevice(g_devs[0].id);
  for (int r : rcs) if (r) return r;
  return B200_OK;
}

int b200_rms2d_full(const float* crdTgt, size_t strideTgt, int nTgt, const int* atomIdxTgt, const float* crdRef,
                    size_t strideRef, int nRef, const int* atomIdxRef, int nAtoms, const double* massTgt,
                    const double* massRefCentering, int fit, float* outFull) {
  if (!crdTgt || !crdRef || !outFull) return fail(B200_ERR_ARG, "null buffer");
  if (nTgt <= 0 || nRef <= 0) return B200_OK;
  std::lock_guard<std::mutex> lk(g_mu);
  int rc;
  if ((rc = ensure_init_locked())) return rc;
  Device& d = g_devs[0];
  CU(cudaSetDevice(d.id));
  int maxT = 0, maxR = 0;
  if ((rc = validate_sel(atomIdxTgt, nAtoms, strideTgt, &maxT))) return rc;
  if ((rc = validate_sel(atomIdxRef, nAtoms, strideRef, &maxR))) return rc;
  cudaStream_t st0 = d.stream[0];
  double h2d = 0.0, d2h = 0.0;
  const size_t wT = (size_t)3 * (maxT + 1), wR = (size_t)3 * (maxR + 1);
  if ((rc = upload_crd(d.crd, crdTgt, strideTgt, 0, nTgt, wT, st0, &h2d))) return rc;
  if ((rc = upload_crd(d.crdB, crdRef, strideRef, 0, nRef, wR, st0, &h2d))) return rc;
  if ((rc = upload_vec(d.idxA, atomIdxTgt, (size_t)nAtoms, st0))) return rc;
  if ((rc = upload_vec(d.idxB, atomIdxRef, (size_t)nAtoms, st0))) return rc;
  if (massTgt && (rc = upload_vec(d.massA, massTgt, (size_t)nAtoms, st0))) return rc;
  const double* mRef = massRefCentering ? massRefCentering : massTgt;
  if (mRef && (rc = upload_vec(d.massB, mRef, (size_t)nAtoms, st0))) return rc;
  if ((rc = d.scal.reserve(64))) return rc;
  double* d_total = (double*)d.scal.p;
  double* d_shift = d_total + 1;
  unsigned int* d_maxBits = (unsigned int*)(d_total + 4);
  const double* dmT = massTgt ? (const double*)d.massA.p : nullptr;
  const double* dmR = mRef ? (const double*)d.massB.p : nullptr;
  Timer tpack, tpair;
  tpack.begin(st0);
  COUNT_LAUNCH();
  mass_sum_kernel<<<1, 32, 0, st0>>>(dmT, nAtoms, d_total);
  // target: centred and weighted with its own masses; reference: centred with the reference
  // mask's masses but weighted with the TARGET masses (src/Frame.cpp:1184-1208, Analysis_Rms2d.cpp:265-266)
  PackSet A, B;
  I8Set qA, qB;
  bool useI8 = false;
  int qs = 0;
  const int engine = pair_engine();
  if (fit && engine != 1) {
    if ((rc = i8_reserve(qA, d.imgA, d.GA, d.cenA, nTgt, nAtoms))) return rc;
    if ((rc = i8_reserve(qB, d.imgB, d.GB, d.cenB, nRef, nAtoms))) return rc;
    CU(cudaMemsetAsync(d_maxBits, 0, sizeof(unsigned int), st0));
    if ((rc = i8_stats(qA, (const float*)d.crd.p, wT, nullptr, 0, 0, (const int*)d.idxA.p, nAtoms, dmT, dmT, d_maxBits, st0))) return rc;
    if ((rc = i8_stats(qB, (const float*)d.crdB.p, wR, nullptr, 0, 0, (const int*)d.idxB.p, nAtoms, dmR, dmT, d_maxBits, st0))) return rc;
    if ((rc = i8_choose_scale(d, d_maxBits, d_total, nAtoms, st0, &qs, &useI8))) return rc;
    if (useI8) {
      if ((rc = i8_clear(qA, st0)) || (rc = i8_clear(qB, st0))) return rc;
      if ((rc = i8_quant(qA, (const float*)d.crd.p, wT, nullptr, 0, 0, (const int*)d.idxA.p, nAtoms, dmT, qs, st0))) return rc;
      if ((rc = i8_quant(qB, (const float*)d.crdB.p, wR, nullptr, 0, 0, (const int*)d.idxB.p, nAtoms, dmT, qs, st0))) return rc;
    } else if (engine == 2) {
      return fail(B200_ERR_ARG, "tcgen05 int8 engine forced but the selection's extent leaves only %d fractional bits", qs);
    }
  } else if (engine == 2) {
    return fail(B200_ERR_ARG, "tcgen05 int8 engine forced but nofit RMSD runs on the FP64 engine only");
  }
  if (!useI8) {
    A.nFrames = nTgt; A.Fpad = round_up(nTgt, ROWG); A.Kpad = round_up(nAtoms, KC);
    B.nFrames = nRef; B.Fpad = round_up(nRef, ROWG); B.Kpad = A.Kpad;
    if ((rc = d.planesA.reserve(plane_doubles(A.Fpad, A.Kpad) * sizeof(double)))) return rc;
    if ((rc = d.planesB.reserve(plane_doubles(B.Fpad, B.Kpad) * sizeof(double)))) return rc;
    if ((rc = d.GA.reserve((size_t)A.Fpad * sizeof(double)))) return rc;
    if ((rc = d.GB.reserve((size_t)B.Fpad * sizeof(double)))) return rc;
    A.planes = (double*)d.planesA.p; A.G = (double*)d.GA.p;
    B.planes = (double*)d.planesB.p; B.G = (double*)d.GB.p;
    if (!fit) { COUNT_LAUNCH(); shift_kernel<<<1, 1, 0, st0>>>((const float*)d.crd.p, wT, nullptr, 0, 0, (const int*)d.idxA.p, d_shift); }
    if ((rc = run_pack((const float*)d.crd.p, wT, nullptr, 0, nTgt, 0, (const int*)d.idxA.p, nAtoms, dmT, dmT, d_shift, fit, A, st0))) return rc;
    if ((rc = run_pack((const float*)d.crdB.p, wR, nullptr, 0, nRef, 0, (const int*)d.idxB.p, nAtoms, dmR, dmT, d_shift, fit, B, st0))) return rc;
  }
  tpack.end(st0);
  g_lastEngine.store(useI8 ? 2 : 1); g_lastQs.store(useI8 ? qs : 0);
  // rows (targets) in bands; each band is a contiguous slab of outFull
  const size_t ld = (size_t)nRef;
  int bandRows = (int)std::min<size_t>(1024, std::max<size_t>(ROWG, ((size_t)32 << 20) / (4 * ld) / ROWG * ROWG));
  const bool pinnedOut = host_ptr_is_pinned(outFull);
  const size_t maxChunk = (size_t)std::min(bandRows, nTgt) * ld;
  CU(cudaEventRecord(d.done[0], st0));
  for (int s = 1; s < NSLOT; ++s) CU(cudaStreamWaitEvent(d.stream[s], d.done[0], 0));
  for (int s = 0; s < NSLOT; ++s) {
    if ((rc = d.outChunk[s].reserve(maxChunk * sizeof(float)))) return rc;
    if (!pinnedOut && (rc = d.outStage[s].reserve(maxChunk * sizeof(float)))) return rc;
  }
  struct Pending { size_t base = 0, n = 0; bool live = false; } pend[NSLOT];
  auto retire = [&](int s) -> int {
    if (!pend[s].live) return B200_OK;
    CU(cudaEventSynchronize(d.done[s]));
    if (!pinnedOut) std::memcpy(outFull + pend[s].base, d.outStage[s].p, pend[s].n * sizeof(float));
    pend[s].live = false;
    return B200_OK;
  };
  int band = 0; long nLaunch = 0;
  for (int i0 = 0; i0 < nTgt; i0 += bandRows, ++band) {
    const int s = band % NSLOT;
    if ((rc = retire(s))) return rc;
    const int i1 = std::min(nTgt, i0 + bandRows);
    cudaStream_t st = d.stream[s];
    const size_t base = (size_t)i0 * ld, n = (size_t)(i1 - i0) * ld;
    tpair.begin(st);
    // kernel indexes out[i*ld + j]; shift the pointer so row i0 lands at the chunk start
    if (useI8)
      rc = run_pair_i8_band(d, qA, qB, i0, i1, false, qs, d_total, (float*)d.outChunk[s].p - base, 0, ld, nullptr, st);
    else
      rc = run_pair_band(A, B, i0 / ROWG, (i1 + ROWG - 1) / ROWG - i0 / ROWG, false, fit != 0, d_total,
                         (float*)d.outChunk[s].p - base, 0, ld, st);
    if (rc) return rc;
    tpair.end(st);
    ++nLaunch;
    float* dst = pinnedOut ? outFull + base : (float*)d.outStage[s].p;
    CU(cudaMemcpyAsync(dst, d.outChunk[s].p, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    d2h += (double)n * sizeof(float);
    CU(cudaEventRecord(d.done[s], st));
    pend[s].base = base; pend[s].n = n; pend[s].live = true;
  }
  for (int s = 0; s < NSLOT; ++s) if ((rc = retire(s))) return rc;
  for (int s = 0; s < NSLOT; ++s) CU(cudaStreamSynchronize(d.stream[s]));
  add_stats(tpack.resolve(), 2, tpair.resolve(), nLaunch, (double)nTgt * (double)nRef, h2d, d2h);
  return B200_OK;
}

int b200_dev_rms2d_tri(const float* d_crd, size_t frameStrideFloats, const int* d_frameIdx, int nFrames,
                       const int* d_atomIdx, int nAtoms, const double* d_mass, int fit, int shardRank, int shardCount,
                       float* d_outTri, void* stream) {
  if (!d_crd || !d_outTri || nAtoms <= 0) return fail(B200_ERR_ARG, "bad argument");
  std::lock_guard<std::mutex> lk(g_mu);
  int rc;
  if ((rc = ensure_init_locked())) return rc;
  int dev = 0;
  CU(cudaGetDevice(&dev));
  Device* d = nullptr;
  for (auto& x : g_devs) if (x.id == dev) d = &x;
  if (!d) return fail(B200_ERR_STATE, "current device %d was not initialised by b200_init", dev);
  int r0 = 0, r1 = 0;
  if ((rc = shard_rows(nFrames, shardRank, shardCount, &r0, &r1))) return rc;
  Timer tpack, tpair;
  cudaStream_t st = (cudaStream_t)stream;
  rc = dev_rms2d_tri(*d, d_crd, frameStrideFloats, d_frameIdx, 0, nFrames, d_atomIdx, nAtoms, d_mass, fit, r0, r1, d_outTri,
                     0, st, 4096, g_profiling ? &tpack : nullptr, g_profiling ? &tpair : nullptr);
  if (rc) return rc;
  if (g_profiling) {
    CU(cudaStreamSynchronize(st));
    const long nl = (long)tpair.ev.size();
    const size_t F = (size_t)nFrames;
    add_stats(tpack.resolve(), 1, tpair.resolve(), nl, nFrames > 1 ? (double)(tri_row_start(F, r1) - tri_row_start(F, r0)) : 0.0, 0, 0);
  }
  return B200_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ one-vs-many
struct b200_1vN {
  Device* dev = nullptr;
  int nAtoms = 0, fit = 1, wantRot = 0;
  std::vector<int> atomIdx;
  DevBuf refw, refsum, idx, in[NSLOT], rms, rot, trans, ws;
  PinBuf stage[NSLOT];
  cudaStream_t st = nullptr;         // single in-order stream (+ events for slot reuse)
  cudaEvent_t slotFree[NSLOT] = {nullptr, nullptr, nullptr};
  long pushed = 0, flushed = 0, cap = 0;
  int slot = 0;
  long bestFrame = -1;
  double bestVal = 0.0;
  Timer timer;
  double h2d = 0.0, d2h = 0.0;
  long launches = 0;
};

static int onevn_grow(b200_1vN* h, long need) {
  if (need <= h->cap) return B200_OK;
  long ncap = std::max<long>(need, std::max<long>(h->cap * 2, 1 << 16));
  DevBuf nr, no, nt;
  int rc;
  if ((rc = nr.reserve((size_t)ncap * 8))) return rc;
  if (h->wantRot) { if ((rc = no.reserve((size_t)ncap * 72))) return rc; if ((rc = nt.reserve((size_t)ncap * 24))) return rc; }
  const long live0 = h->flushed, live1 = h->pushed;
  if (live1 > live0) {
    CU(cudaMemcpyAsync((double*)nr.p + live0, (double*)h->rms.p + live0, (size_t)(live1 - live0) * 8, cudaMemcpyDeviceToDevice, h->st));
    if (h->wantRot) {
      CU(cudaMemcpyAsync((double*)no.p + 9 * live0, (double*)h->rot.p + 9 * live0, (size_t)(live1 - live0) * 72, cudaMemcpyDeviceToDevice, h->st));
      CU(cudaMemcpyAsync((double*)nt.p + 3 * live0, (double*)h->trans.p + 3 * live0, (size_t)(live1 - live0) * 24, cudaMemcpyDeviceToDevice, h->st));
    }
  }
  CU(cudaStreamSynchronize(h->st));
  h->rms.release(); h->rot.release(); h->trans.release();
  h->rms = nr; h->rot = no; h->trans = nt;
  h->cap = ncap;
  return B200_OK;
}

/// Launches the one-vs-many kernels for `nFrames` frames at d_crd on stream st: chunk table, TMA streaming kernel +
/// per-frame finish (sorted selections, 16-byte aligned base), then the general gather kernel, which returns at once
/// when the streaming variant did the work (the choice is made on the device: no host round trip).
/// ws: workspace of at least onevn_ws_bytes() bytes.
static size_t onevn_ws_bytes(size_t stride, int nFrames) {
  const size_t maxChunks = stride / 3 / (ONEVN_S_CHUNK_BYTES / 24) + 2;   // (double frames: the smaller chunk)
  return 64 + (maxChunks + 1) * sizeof(int) + 64 + (size_t)nFrames * ONEVN_REC * sizeof(double);
}
template <typename T>
static int onevn_run(int numSMs, const void* d_crd, size_t stride, int nFrames, const int* d_atomIdx, int nAtoms,
                     const double* refw, const double* refsum, int fit, double* rmsd, double* rot, double* trans,
                     void* ws, cudaStream_t st, const int* d_frameIdx = nullptr, long srcBase = 0) {
  constexpr int APC = ONEVN_S_CHUNK_BYTES / (3 * (int)sizeof(T));
  const int maxChunks = (int)std::min<size_t>(stride / 3 / APC + 2, 1u << 20);
  int* hdr = (int*)ws;
  int* kLo = hdr + 16;
  double* rec = (double*)((char*)ws + ((64 + (size_t)(maxChunks + 1) * sizeof(int) + 63) & ~(size_t)63));
  const bool aligned = (((uintptr_t)d_crd) & 15) == 0;
  const char* env = getenv("B200_1VN_STREAM");
  const bool stream = aligned && !(env && atoi(env) == 0);
  if (stream) {
    static bool attr[64][2] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr[dev & 63][sizeof(T) == 8]) {
      CU(cudaFuncSetAttribute(onevn_stream_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, ONEVN_S_SMEM_BYTES));
      attr[dev & 63][sizeof(T) == 8] = true;
    }
    COUNT_LAUNCH();
    onevn_chunks_kernel<<<1, 256, 0, st>>>(d_atomIdx, nAtoms, APC, maxChunks, hdr, kLo);
    OneVNStreamArgs sa;
    sa.crd = d_crd; sa.stride = stride; sa.frameIdx = d_frameIdx; sa.srcBase = srcBase; sa.nFrames = nFrames;
    sa.atomIdx = d_atomIdx; sa.nAtoms = nAtoms;
    sa.refw = refw; sa.hdr = hdr; sa.kLo = kLo; sa.fit = fit; sa.rec = rec;
    const int nGroups = (nFrames + ONEVN_FB - 1) / ONEVN_FB;
    COUNT_LAUNCH();
    onevn_stream_kernel<T><<<std::min(nGroups, numSMs > 0 ? numSMs : 148), ONEVN_S_THREADS, ONEVN_S_SMEM_BYTES, st>>>(sa);
    COUNT_LAUNCH();
    onevn_finish_kernel<<<(nFrames + 127) / 128, 128, 0, st>>>(rec, hdr, nFrames, refsum, fit, rmsd, rot, trans);
  }
  OneVNArgs a;
  a.crd = d_crd; a.stride = stride; a.frameIdx = d_frameIdx; a.srcBase = srcBase; a.nFrames = nFrames;
  a.atomIdx = d_atomIdx; a.nAtoms = nAtoms;
  a.refw = refw; a.refsum = refsum; a.skipIf = stream ? hdr : nullptr; a.fit = fit;
  a.rmsd = rmsd; a.rot = rot; a.trans = trans;
  COUNT_LAUNCH();
  onevn_kernel<T><<<(nFrames + ONEVN_FB - 1) / ONEVN_FB, ONEVN_THREADS, 0, st>>>(a);
  CU(cudaGetLastError());
  return B200_OK;
}

template <typename T>
static int onevn_launch(b200_1vN* h, const void* d_crd, size_t stride, int nFrames, const int* d_atomIdx, long outOffset) {
  int rc;
  if ((rc = h->ws.reserve(onevn_ws_bytes(stride, nFrames)))) return rc;
  h->timer.begin(h->st);
  rc = onevn_run<T>(h->dev->numSMs, d_crd, stride, nFrames, d_atomIdx, h->nAtoms, (const double*)h->refw.p,
                    (const double*)h->refsum.p, h->fit, (double*)h->rms.p + outOffset,
                    h->wantRot ? (double*)h->rot.p + 9 * outOffset : nullptr,
                    h->wantRot ? (double*)h->trans.p + 3 * outOffset : nullptr, h->ws.p, h->st);
  h->timer.end(h->st);
  if (rc) return rc;
  h->launches++;
  return B200_OK;
}

extern "C" {

int b200_rmsd_1vN_begin(const double* refSelected, const int* atomIdx, int nAtoms, const double* mass, int fit, int wantRot,
                        b200_1vN** handle) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  *handle = nullptr;
  if (!refSelected || !atomIdx || nAtoms <= 0) return fail(B200_ERR_ARG, "bad reference / selection");
  std::lock_guard<std::mutex> lk(g_mu);
  int rc;
  if ((rc = ensure_init_locked())) return rc;
  b200_1vN* h = new (std::nothrow) b200_1vN();
  if (!h) return fail(B200_ERR_NOMEM, "out of memory");
  h->dev = &g_devs[0];
  CU(cudaSetDevice(h->dev->id));
  h->nAtoms = nAtoms; h->fit = fit ? 1 : 0; h->wantRot = (wantRot && fit) ? 1 : 0;
  h->atomIdx.assign(atomIdx, atomIdx + nAtoms);
  for (int k = 0; k < nAtoms; ++k) if (atomIdx[k] < 0) { delete h; return fail(B200_ERR_ARG, "negative atom index"); }
  CU(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
  for (int s = 0; s < NSLOT; ++s) CU(cudaEventCreateWithFlags(&h->slotFree[s], cudaEventDisableTiming));
  DevBuf ref, m;
  if ((rc = upload_vec(ref, refSelected, (size_t)3 * nAtoms, h->st))) { delete h; return rc; }
  if (mass && (rc = upload_vec(m, mass, (size_t)nAtoms, h->st))) { delete h; return rc; }
  if ((rc = upload_vec(h->idx, atomIdx, (size_t)nAtoms, h->st))) { delete h; return rc; }
  if ((rc = h->refw.reserve((size_t)nAtoms * 32))) { delete h; return rc; }
  if ((rc = h->refsum.reserve(64))) { delete h; return rc; }
  COUNT_LAUNCH();
  onevn_setup_kernel<<<1, 256, 0, h->st>>>((const double*)ref.p, mass ? (const double*)m.p : nullptr, nAtoms, (double*)h->refw.p,
                                          (double*)h->refsum.p);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->st));
  ref.release(); m.release();
  *handle = h;
  return B200_OK;
}

}  // extern "C"

template <typename T>
static int onevn_push(b200_1vN* h, const T* src, size_t stride, int nFrames) {
  if (!h) return fail(B200_ERR_STATE, "null handle");
  if (nFrames <= 0) return B200_OK;
  if (!src) return fail(B200_ERR_ARG, "null frames");
  std::lock_guard<std::mutex> lk(g_mu);
  CU(cudaSetDevice(h->dev->id));
  int rc;
  int maxAtom = 0;
  for (int a : h->atomIdx) maxAtom = std::max(maxAtom, a);
  if ((size_t)3 * ((size_t)maxAtom + 1) > stride) return fail(B200_ERR_ARG, "atom index %d outside frame stride %zu", maxAtom, stride);
  if ((rc = onevn_grow(h, h->pushed + nFrames))) return rc;
  const bool pinned = host_ptr_is_pinned(src);
  const int N = h->nAtoms;
  // chunk so that a slot holds <= 64 MB
  const size_t width = (size_t)3 * ((size_t)maxAtom + 1);
  const size_t perFrame = pinned ? width * sizeof(T) : (size_t)3 * N * sizeof(T);
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)nFrames, ((size_t)64 << 20) / perFrame));
  for (int f0 = 0; f0 < nFrames; f0 += chunk) {
    const int nf = std::min(chunk, nFrames - f0);
    const int s = h->slot;
    h->slot = (h->slot + 1) % NSLOT;
    CU(cudaEventSynchronize(h->slotFree[s]));  // previous user of this slot is done (no-op when never recorded)
    if ((rc = h->in[s].reserve((size_t)nf * perFrame))) return rc;
    if (pinned) {
      // direct DMA of the needed span; gather by atomIdx on the device
      CU(cudaMemcpy2DAsync(h->in[s].p, width * sizeof(T), src + (size_t)f0 * stride, stride * sizeof(T), width * sizeof(T),
                           (size_t)nf, cudaMemcpyHostToDevice, h->st));
      if ((rc = onevn_launch<T>(h, h->in[s].p, width, nf, (const int*)h->idx.p, h->pushed))) return rc;
    } else {
      // pageable source (e.g. a cpptraj Frame that is reused): gather selected atoms into pinned staging now
      if ((rc = h->stage[s].reserve((size_t)nf * perFrame))) return rc;
      T* stg = (T*)h->stage[s].p;
      for (int f = 0; f < nf; ++f) {
        const T* fr = src + (size_t)(f0 + f) * stride;
        T* o = stg + (size_t)f * 3 * N;
        for (int k = 0; k < N; ++k) {
          const size_t a3 = (size_t)3 * h->atomIdx[k];
          o[3 * k] = fr[a3]; o[3 * k + 1] = fr[a3 + 1]; o[3 * k + 2] = fr[a3 + 2];
        }
      }
      CU(cudaMemcpyAsync(h->in[s].p, stg, (size_t)nf * perFrame, cudaMemcpyHostToDevice, h->st));
      if ((rc = onevn_launch<T>(h, h->in[s].p, (size_t)3 * N, nf, nullptr, h->pushed))) return rc;
    }
    CU(cudaEventRecord(h->slotFree[s], h->st));
    h->h2d += (double)nf * perFrame;
    h->pushed += nf;
  }
  return B200_OK;
}

extern "C" {

int b200_rmsd_1vN_push_f64(b200_1vN* h, const double* xyz, size_t frameStrideDoubles, int nFrames) {
  return onevn_push<double>(h, xyz, frameStrideDoubles, nFrames);
}
int b200_rmsd_1vN_push_f32(b200_1vN* h, const float* crd, size_t frameStrideFloats, int nFrames) {
  return onevn_push<float>(h, crd, frameStrideFloats, nFrames);
}
long b200_rmsd_1vN_pending(const b200_1vN* h) { return h ? h->pushed - h->flushed : 0; }

int b200_rmsd_1vN_flush(b200_1vN* h, double* rmsdOut, double* rotOut, double* transOut, long* argminFrame) {
  if (!h) return fail(B200_ERR_STATE, "null handle");
  std::lock_guard<std::mutex> lk(g_mu);
  CU(cudaSetDevice(h->dev->id));
  const long n = h->pushed - h->flushed;
  if (n > 0) {
    if (!rmsdOut) return fail(B200_ERR_ARG, "null rmsdOut");
    CU(cudaMemcpyAsync(rmsdOut, (double*)h->rms.p + h->flushed, (size_t)n * 8, cudaMemcpyDeviceToHost, h->st));
    h->d2h += (double)n * 8;
    if (rotOut && h->wantRot) { CU(cudaMemcpyAsync(rotOut, (double*)h->rot.p + 9 * h->flushed, (size_t)n * 72, cudaMemcpyDeviceToHost, h->st)); h->d2h += (double)n * 72; }
    if (transOut && h->wantRot) { CU(cudaMemcpyAsync(transOut, (double*)h->trans.p + 3 * h->flushed, (size_t)n * 24, cudaMemcpyDeviceToHost, h->st)); h->d2h += (double)n * 24; }
  }
  CU(cudaStreamSynchronize(h->st));
  for (long i = 0; i < n; ++i) {
    if (h->bestFrame < 0 || rmsdOut[i] < h->bestVal) { h->bestVal = rmsdOut[i]; h->bestFrame = h->flushed + i; }
  }
  if (argminFrame) *argminFrame = h->bestFrame;
  {
    const double ms = h->timer.resolve();
    std::lock_guard<std::mutex> sl(g_statMu);
    g_stats.onevn_ms += ms; g_stats.onevn_launches += h->launches; g_stats.frames_1vN += (double)n;
    g_stats.h2d_bytes += h->h2d; g_stats.d2h_bytes += h->d2h;
    h->launches = 0; h->h2d = 0; h->d2h = 0;
  }
  h->flushed = h->pushed;
  return B200_OK;
}

int b200_rmsd_1vN_end(b200_1vN* h) {
  if (!h) return B200_OK;
  std::lock_guard<std::mutex> lk(g_mu);
  cudaSetDevice(h->dev->id);
  if (h->st) { cudaStreamSynchronize(h->st); cudaStreamDestroy(h->st); }
  for (int s = 0; s < NSLOT; ++s) { if (h->slotFree[s]) cudaEventDestroy(h->slotFree[s]); h->in[s].release(); h->stage[s].release(); }
  h->refw.release(); h->refsum.release(); h->idx.release(); h->rms.release(); h->rot.release(); h->trans.release(); h->ws.release();
  h->timer.resolve();
  delete h;
  return B200_OK;
}

int b200_dev_rmsd_1vN(const float* d_crd, size_t frameStrideFloats, int nFrames, const int* d_atomIdx, int nAtoms,
                      const double* d_refSelected, const double* d_mass, int fit, double* d_rmsdOut, double* d_rotOut,
                      double* d_transOut, void* stream) {
  if (!d_crd || !d_refSelected || !d_rmsdOut || nAtoms <= 0) return fail(B200_ERR_ARG, "bad argument");
  if (nFrames <= 0) return B200_OK;
  std::lock_guard<std::mutex> lk(g_mu);
  int rc;
  if ((rc = ensure_init_locked())) return rc;
  int dev = 0;
  CU(cudaGetDevice(&dev));
  Device* d = nullptr;
  for (auto& x : g_devs) if (x.id == dev) d = &x;
  if (!d) return fail(B200_ERR_STATE, "current device %d was not initialised by b200_init", dev);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t refBytes = ((size_t)nAtoms * 32 + 64 + 63) & ~(size_t)63;
  if ((rc = d->onevnWs.reserve(refBytes + onevn_ws_bytes(frameStrideFloats, nFrames)))) return rc;
  double* refw = (double*)d->onevnWs.p;
  double* refsum = refw + (size_t)4 * nAtoms;
  COUNT_LAUNCH();
  onevn_setup_kernel<<<1, 256, 0, st>>>(d_refSelected, d_mass, nAtoms, refw, refsum);
  Timer t;
  if (g_profiling) t.begin(st);
  rc = onevn_run<float>(d->numSMs, d_crd, frameStrideFloats, nFrames, d_atomIdx, nAtoms, refw, refsum, fit ? 1 : 0, d_rmsdOut,
                        fit ? d_rotOut : nullptr, fit ? d_transOut : nullptr, (char*)d->onevnWs.p + refBytes, st);
  if (g_profiling) t.end(st);
  if (rc) return rc;
  if (g_profiling) {
    CU(cudaStreamSynchronize(st));
    const double ms = t.resolve();
    std::lock_guard<std::mutex> sl(g_statMu);
    g_stats.onevn_ms += ms; g_stats.onevn_launches += 1; g_stats.frames_1vN += (double)nFrames;
  }
  return B200_OK;
}


int b200_rmsd_frames_to_centroids(const float* crd, size_t frameStrideFloats, int nFramesTotal, const int* frameIdx, int nFrames,
                                  const int* atomIdx, int nAtoms, const double* mass, int fit, const double* centroids,
                                  int nCentroids, double* distOut, int* closestOut, double* closestDistOut) {
  if (!crd || !centroids || nCentroids <= 0) return fail(B200_ERR_ARG, "bad argument");
  if (nFrames <= 0) return B200_OK;
  std::lock_guard<std::mutex> lk(g_mu);
  int rc;
  if ((rc = ensure_init_locked())) return rc;
  Device& d = g_devs[0];
  CU(cudaSetDevice(d.id));
  int maxAtom = 0;
  if ((rc = validate_sel(atomIdx, nAtoms, frameStrideFloats, &maxAtom))) return rc;
  int sLo = 0, sHi = nFrames;
  if (frameIdx) {
    sLo = nFramesTotal; sHi = 0;
    for (int f = 0; f < nFrames; ++f) {
      if (frameIdx[f] < 0 || frameIdx[f] >= nFramesTotal) return fail(B200_ERR_ARG, "frameIdx[%d]=%d out of range", f, frameIdx[f]);
      sLo = std::min(sLo, frameIdx[f]); sHi = std::max(sHi, frameIdx[f] + 1);
    }
  } else if (nFrames > nFramesTotal) {
    return fail(B200_ERR_ARG, "nFrames %d > nFramesTotal %d", nFrames, nFramesTotal);
  }
  cudaStream_t st = d.stream[0];
  double h2d = 0.0;
  const size_t width = (size_t)3 * ((size_t)maxAtom + 1);
  if ((rc = upload_crd(d.crd, crd, frameStrideFloats, sLo, sHi, width, st, &h2d))) return rc;
  if ((rc = upload_vec(d.idxA, atomIdx, (size_t)nAtoms, st))) return rc;
  if (mass && (rc = upload_vec(d.massA, mass, (size_t)nAtoms, st))) return rc;
  if (frameIdx && (rc = upload_vec(d.frameIdx, frameIdx, (size_t)nFrames, st))) return rc;
  DevBuf cen;
  if ((rc = upload_vec(cen, centroids, (size_t)nCentroids * 3 * (size_t)nAtoms, st))) return rc;
  // workspace: refw + refsum | one-vs-many workspace | dist [K][nFrames] | outputs
  const size_t refBytes = ((size_t)nAtoms * 32 + 64 + 63) & ~(size_t)63;
  const size_t wsBytes = (onevn_ws_bytes(width, nFrames) + 63) & ~(size_t)63;
  const size_t distBytes = (size_t)nCentroids * (size_t)nFrames * sizeof(double);
  const size_t outBytes = distBytes + (size_t)nFrames * (sizeof(int) + sizeof(double)) + 64;
  if ((rc = d.onevnWs.reserve(refBytes + wsBytes + distBytes + outBytes))) { cen.release(); return rc; }
  char* base = (char*)d.onevnWs.p;
  double* refw = (double*)base;
  double* refsum = refw + (size_t)4 * nAtoms;
  void* ws = base + refBytes;
  double* dist = (double*)(base + refBytes + wsBytes);
  double* dOutT = (double*)((char*)dist + distBytes);
  double* dClosestDist = dOutT + (size_t)nCentroids * nFrames;
  int* dClosest = (int*)(dClosestDist + nFrames);
  const double* d_mass = mass ? (const double*)d.massA.p : nullptr;
  Timer t;
  t.begin(st);
  for (int k = 0; k < nCentroids; ++k) {
    COUNT_LAUNCH();
    onevn_setup_kernel<<<1, 256, 0, st>>>((const double*)cen.p + (size_t)k * 3 * nAtoms, d_mass, nAtoms, refw, refsum);
    if ((rc = onevn_run<float>(d.numSMs, d.crd.p, width, nFrames, (const int*)d.idxA.p, nAtoms, refw, refsum, fit ? 1 : 0,
                               dist + (size_t)k * nFrames, nullptr, nullptr, ws, st,
                               frameIdx ? (const int*)d.frameIdx.p : nullptr, (long)sLo))) { cen.release(); return rc; }
  }
  COUNT_LAUNCH();
  centroid_argmin_kernel<<<(nFrames + 255) / 256, 256, 0, st>>>(dist, nFrames, nCentroids, distOut ? dOutT : nullptr, dClosest,
                                                                 dClosestDist);
  t.end(st);
  double d2h = 0.0;
  if (distOut) { CU(cudaMemcpyAsync(distOut, dOutT, distBytes, cudaMemcpyDeviceToHost, st)); d2h += (double)distBytes; }
  if (closestOut) { CU(cudaMemcpyAsync(closestOut, dClosest, (size_t)nFrames * sizeof(int), cudaMemcpyDeviceToHost, st)); d2h += 4.0 * nFrames; }
  if (closestDistOut) { CU(cudaMemcpyAsync(closestDistOut, dClosestDist, (size_t)nFrames * sizeof(double), cudaMemcpyDeviceToHost, st)); d2h += 8.0 * nFrames; }
  CU(cudaStreamSynchronize(st));
  cen.release();
  {
    const double ms = t.resolve();
    std::lock_guard<std::mutex> sl(g_statMu);
    g_stats.onevn_ms += ms; g_stats.onevn_launches += nCentroids; g_stats.frames_1vN += (double)nFrames * nCentroids;
    g_stats.h2d_bytes += h2d; g_stats.d2h_bytes += d2h;
  }
  return B200_OK;
}

int b200_debug_i8_clocks(long long* out, int ctas) {
  // Arms (out == NULL) or reads back (out != NULL) the per-CTA cycle counters of pair_i8_kernel: 16 per CTA.
  std::lock_guard<std::mutex> lk(g_mu);
  if (ensure_init_locked()) return B200_ERR_NO_DEVICE;
  if (!out) {
    if (!g_dbgClk) CU(cudaMalloc(&g_dbgClk, 16 * 1024 * sizeof(long long)));
    CU(cudaMemset(g_dbgClk, 0, 16 * 1024 * sizeof(long long)));
    return B200_OK;
  }
  if (!g_dbgClk) return fail(B200_ERR_STATE, "clock counters not armed");
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(out, g_dbgClk, (size_t)16 * std::min(ctas, 1024) * sizeof(long long), cudaMemcpyDeviceToHost));
  CU(cudaFree(g_dbgClk));
  g_dbgClk = nullptr;
  return B200_OK;
}

int b200_set_pair_engine(int engine) {
  if (engine < 0 || engine > 2) return fail(B200_ERR_ARG, "engine must be 0 (auto), 1 (fp64) or 2 (tcgen05 int8)");
  g_engine = engine;
  return B200_OK;
}
int b200_set_i8_cta_group(int ctaGroup) {
  if (ctaGroup != 1 && ctaGroup != 2) return fail(B200_ERR_ARG, "cta group must be 1 or 2");
  g_i8Cg = ctaGroup;
  return B200_OK;
}
int b200_get_i8_cta_group(void) { return i8_cta_group(); }
int b200_last_pair_engine(int* fractionalBits) {
  if (fractionalBits) *fractionalBits = g_lastQs.load();
  return g_lastEngine.load();
}

int b200_debug_i8(const float* crd, size_t frameStrideFloats, int nFrames, const int* atomIdx, int nAtoms, const double* mass,
                  unsigned char* imageOut, size_t imageCap, size_t* imageBytes, double* GOut, double* SOut, float* outTri,
                  int* qsOut) {
  if (!crd || nFrames < 2) return fail(B200_ERR_ARG, "bad argument");
  std::lock_guard<std::mutex> lk(g_mu);
  int rc, maxAtom = 0;
  if ((rc = ensure_init_locked())) return rc;
  if ((rc = validate_sel(atomIdx, nAtoms, frameStrideFloats, &maxAtom))) return rc;
  Device& d = g_devs[0];
  CU(cudaSetDevice(d.id));
  cudaStream_t st = d.stream[0];
  double h2d = 0.0;
  const size_t width = (size_t)3 * ((size_t)maxAtom + 1);
  if ((rc = upload_crd(d.crd, crd, frameStrideFloats, 0, nFrames, width, st, &h2d))) return rc;
  if ((rc = upload_vec(d.idxA, atomIdx, (size_t)nAtoms, st))) return rc;
  if (mass && (rc = upload_vec(d.massA, mass, (size_t)nAtoms, st))) return rc;
  const double* d_mass = mass ? (const double*)d.massA.p : nullptr;
  if ((rc = d.scal.reserve(64))) return rc;
  double* d_total = (double*)d.scal.p;
  unsigned int* d_maxBits = (unsigned int*)(d_total + 4);
  COUNT_LAUNCH();
  mass_sum_kernel<<<1, 32, 0, st>>>(d_mass, nAtoms, d_total);
  I8Set q;
  if ((rc = i8_reserve(q, d.imgA, d.GA, d.cenA, nFrames, nAtoms))) return rc;
  CU(cudaMemsetAsync(d_maxBits, 0, sizeof(unsigned int), st));
  if ((rc = i8_stats(q, (const float*)d.crd.p, width, nullptr, 0, 0, (const int*)d.idxA.p, nAtoms, d_mass, d_mass, d_maxBits, st))) return rc;
  int qs = 0; bool ok = false;
  if ((rc = i8_choose_scale(d, d_maxBits, d_total, nAtoms, st, &qs, &ok))) return rc;
  if (qsOut) *qsOut = qs;
  if ((rc = i8_clear(q, st))) return rc;
  if ((rc = i8_quant(q, (const float*)d.crd.p, width, nullptr, 0, 0, (const int*)d.idxA.p, nAtoms, d_mass, qs, st))) return rc;
  const size_t ib = i8_image_bytes(q.nRg, q.nC);
  if (imageBytes) *imageBytes = ib;
  if (imageOut) {
    if (imageCap < ib) return fail(B200_ERR_ARG, "image buffer too small: %zu < %zu", imageCap, ib);
    CU(cudaMemcpyAsync(imageOut, q.image, ib, cudaMemcpyDeviceToHost, st));
  }
  if (GOut) CU(cudaMemcpyAsync(GOut, q.G, (size_t)nFrames * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (outTri) {
    const size_t nTri = (size_t)nFrames * (nFrames - 1) / 2;
    double* d_S = nullptr;
    if (SOut) {
      if ((rc = d.dbgS.reserve((size_t)nFrames * nFrames * 9 * sizeof(double)))) return rc;
      d_S = (double*)d.dbgS.p;
      CU(cudaMemsetAsync(d_S, 0, (size_t)nFrames * nFrames * 9 * sizeof(double), st));
    }
    if ((rc = d.outChunk[0].reserve(nTri * sizeof(float)))) return rc;
    if ((rc = run_pair_i8_band(d, q, q, 0, nFrames, true, qs, d_total, (float*)d.outChunk[0].p, 0, 0, d_S, st))) return rc;
    CU(cudaMemcpyAsync(outTri, d.outChunk[0].p, nTri * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (SOut) CU(cudaMemcpyAsync(SOut, d_S, (size_t)nFrames * nFrames * 9 * sizeof(double), cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(st));
  return B200_OK;
}


double b200_measure_i8_mma_peak(void) { return b200_measure_i8_mma_peak_variant(0); }

double b200_measure_i8_mma_peak_variant(int variant) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (ensure_init_locked()) return -1.0;
  cudaDeviceProp prop;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return -1.0;
  const int blocks = prop.multiProcessorCount, iters = 20000;
  const int smem = 3 * I8_BLK_BYTES;
  if (cudaFuncSetAttribute(i8_mma_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1.0;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  i8_mma_peak_kernel<<<blocks, 128, smem>>>(256, nullptr, variant);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a);
    i8_mma_peak_kernel<<<blocks, 128, smem>>>(iters, nullptr, variant);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    best = std::min(best, ms);
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  if (cudaGetLastError() != cudaSuccess) return -1.0;
  return 2.0 * 128 * 256 * 32 * (double)iters * blocks / (best * 1e-3) / 1e12;   // TOP/s
}

double b200_debug_i8_mma_latency(int nMma, int ctas) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (ensure_init_locked()) return -1.0;
  if (nMma < 0 || ctas < 1 || ctas > 148) return -1.0;
  const int smem = 3 * I8_BLK_BYTES;
  if (cudaFuncSetAttribute(i8_mma_latency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return -1.0;
  long long* d_out = nullptr;
  if (cudaMalloc(&d_out, 148 * sizeof(long long)) != cudaSuccess) return -1.0;
  i8_mma_latency_kernel<<<ctas, 128, smem>>>(nMma, 200, d_out);
  long long h[148];
  double avg = -1.0;
  if (cudaDeviceSynchronize() == cudaSuccess && cudaMemcpy(h, d_out, ctas * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
    avg = 0.0;
    for (int i = 0; i < ctas; ++i) avg += (double)h[i];
    avg /= ctas;
  }
  cudaFree(d_out);
  return avg;
}

double b200_measure_fp64_mma_peak(int variant) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (ensure_init_locked()) return -1.0;
  cudaDeviceProp prop;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return -1.0;
  const int blocks = prop.multiProcessorCount * 4, threads = 128, iters = 4096;
  double* sink = nullptr;
  if (cudaMalloc(&sink, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return -1.0;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  double flopPerWarpIter = 0.0;
  auto launch = [&](int it) {
    switch (variant) {
      case 0: fp64_mma_peak_kernel<0><<<blocks, threads>>>(sink, it); break;
      case 1: fp64_mma_peak_kernel<1><<<blocks, threads>>>(sink, it); break;
      case 2: fp64_mma_peak_kernel<2><<<blocks, threads>>>(sink, it); break;
      case 3: fp64_mma_peak_kernel<3><<<blocks, threads>>>(sink, it); break;
      default: fp64_mma_peak_kernel<4><<<blocks, threads>>>(sink, it); break;
    }
  };
  switch (variant) {
    case 0: flopPerWarpIter = 12.0 * 2 * 2.0 * 8 * 8 * 4; break;
    case 1: flopPerWarpIter = 12.0 * 2.0 * 16 * 8 * 4; break;
    case 2: flopPerWarpIter = 12.0 * 2.0 * 16 * 8 * 8; break;
    case 3: flopPerWarpIter = 12.0 * 2.0 * 16 * 8 * 16; break;
    default: flopPerWarpIter = 12.0 * 4 * 2.0 * 32; break;
  }
  launch(64);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a);
    launch(iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    best = std::min(best, ms);
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  cudaFree(sink);
  if (cudaGetLastError() != cudaSuccess) return -1.0;
  const double warps = (double)blocks * threads / 32.0;
  return flopPerWarpIter * warps * iters / (best * 1e-3) / 1e12;
}

}  // extern "C"
