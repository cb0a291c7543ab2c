// b200_rmsd.cu -- host side of the C ABI declared in include/b200_rmsd.h.
//
// No torch, no CPU fallback: every compute entry point needs an sm_100 device.
// Device memory, streams and pinned staging are owned here; cpptraj owns the
// host buffers it passes in.
#include "b200_rmsd.h"         // (include/ of this repository; src/cuda_b200/ in a cpptraj tree)
#include "b200_rmsd_debug.h"
#include "rmsd_kernels.cuh"
#include "pair_i8.cuh"
#include "hieragglo.cuh"
#include "avgcorr.cuh"
#include "host_util.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
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
// B200_ALLOC_TRACE=1: what a call spent in cudaMalloc / cudaMallocHost (the fixed cost of a process's first call)
struct AllocTrace {
  std::atomic<long long> devNs{0}, pinNs{0}, devBytes{0}, pinBytes{0};
  bool on = getenv("B200_ALLOC_TRACE") != nullptr;
  static long long now() { return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
  void report(const char* what, long long callNs) {
    if (!on) return;
    fprintf(stderr, "B200_ALLOC_TRACE %s: %.4f s in the call; cudaMalloc %.4f s (%.1f MB), cudaMallocHost %.4f s (%.1f MB)\n", what, callNs * 1e-9,
            devNs.exchange(0) * 1e-9, devBytes.exchange(0) / 1e6, pinNs.exchange(0) * 1e-9, pinBytes.exchange(0) / 1e6);
  }
};
AllocTrace g_allocTrace;
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return B200_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    // grow a little beyond the request to avoid re-allocation churn
    size_t want = bytes + bytes / 8 + 256;
    const long long t0 = g_allocTrace.on ? AllocTrace::now() : 0;
    cudaError_t e = cudaMalloc(&p, want);
    if (g_allocTrace.on) { g_allocTrace.devNs += AllocTrace::now() - t0; g_allocTrace.devBytes += (long long)want; }
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
    const long long t0 = g_allocTrace.on ? AllocTrace::now() : 0;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (g_allocTrace.on) { g_allocTrace.pinNs += AllocTrace::now() - t0; g_allocTrace.pinBytes += (long long)bytes; }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(B200_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e)); }
    cap = bytes;
    return B200_OK;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

constexpr int NSLOT = 5;   // streams per device: stream 0 uploads / packs (pipelined path), all of them carry bands + downloads
constexpr int NIN = 2;     // pinned ring slots for staged uploads of pageable COORDS

struct Device {
  int id = -1;
  cudaStream_t stream[NSLOT] = {};
  cudaStream_t copyIn = nullptr;          // uploads of the pipelined host path: DMA back to back, never behind a kernel
  cudaEvent_t done[NSLOT] = {};
  // workspaces (grow-only)
  DevBuf crd, crdB, idxA, idxB, frameIdx, massA, massB, planesA, planesB, GA, GB, scal, onevnWs;
  DevBuf imgA, imgB, cenA, cenB, dbgS;   // tcgen05 int8 path: operand images, frame centres
  DevBuf acP, acRms, acMisc;              // rmsavgcorr: prefix sums of the selected coordinates, RMSDs of a window batch, per-window data
  DevBuf haTri, haD, haS, haMisc;         // hierarchical clustering: uploaded cache triangle, n x n cluster distances, linkage sums, per-cluster state
  PinBuf hostScal;                        // pinned slot for the few scalars read back per call
  int numSMs = 0;
  DevBuf outChunk[NSLOT];
  PinBuf outStage[NSLOT];                 // pageable result buffers: a band lands here, the copy pool moves it on
  PinBuf inStage[NIN];                    // pageable COORDS: the copy pool gathers rows here, then one H2D copy
  cudaEvent_t inFree[NIN] = {};
  bool inUsed[NIN] = {};
  int inNext = 0;
  // COORDS kept on the device between calls (b200_coords_resident_begin/end): the leading resWidth floats of all frames
  const float* resHost = nullptr; size_t resStride = 0, resWidth = 0; int resFrames = 0; DevBuf resBuf;
  // pairwise cache kept on the device between calls (b200_cache_resident_begin/end): the triangle of cacheN frames
  const float* cacheHost = nullptr; int cacheN = 0; DevBuf cacheBuf;
  CopyPool pool;                          // staging of pageable inputs (and everything else on the calling thread)
  CopyPool poolOut;                       // delivery of pageable results (driven by OutRing's own thread)
  void destroy() {
    if (id < 0) return;
    cudaSetDevice(id);
    for (int s = 0; s < NSLOT; ++s) {
      if (stream[s]) cudaStreamDestroy(stream[s]);
      if (done[s]) cudaEventDestroy(done[s]);
      stream[s] = nullptr; done[s] = nullptr;
      outChunk[s].release(); outStage[s].release();
    }
    if (copyIn) cudaStreamDestroy(copyIn);
    copyIn = nullptr;
    for (int b = 0; b < NIN; ++b) {
      if (inFree[b]) cudaEventDestroy(inFree[b]);
      inFree[b] = nullptr; inUsed[b] = false;
      inStage[b].release();
    }
    DevBuf* all[] = {&crd, &crdB, &idxA, &idxB, &frameIdx, &massA, &massB, &planesA, &planesB, &GA, &GB, &scal, &onevnWs,
                     &imgA, &imgB, &cenA, &cenB, &dbgS, &haTri, &haD, &haS, &haMisc, &acP, &acRms, &acMisc};
    for (DevBuf* b : all) b->release();
    hostScal.release(); resBuf.release(); resHost = nullptr; cacheBuf.release(); cacheHost = nullptr; cacheN = 0;
    pool.stop(); poolOut.stop();
    id = -1;
  }
};

std::mutex g_mu;               // serialises public entry points (re-entrant across sequential calls)
/// The initialised devices (a Device owns threads and mutexes: held by pointer, handed out by reference).
struct DevList {
  std::vector<std::unique_ptr<Device>> v;
  struct It {
    std::vector<std::unique_ptr<Device>>::iterator p;
    Device& operator*() const { return **p; }
    It& operator++() { ++p; return *this; }
    bool operator!=(const It& o) const { return p != o.p; }
  };
  It begin() { return It{v.begin()}; }
  It end() { return It{v.end()}; }
  size_t size() const { return v.size(); }
  bool empty() const { return v.empty(); }
  Device& operator[](size_t i) { return *v[i]; }
  void clear() { v.clear(); }
  void resize(size_t n) { v.clear(); for (size_t i = 0; i < n; ++i) v.emplace_back(new Device()); }
};
DevList g_devs;
bool g_inited = false;

// ------------------------------------------------------------------ stats
std::mutex g_statMu;
b200_stats g_stats = {};
bool g_profiling = false;
std::atomic<long> g_launches{0};
struct Timer;

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

// CUDA events around the streaming kernel of the one-vs-many path alone (the dominant kernel of that path); resolved by
// b200_get_stats after a device synchronisation.
std::mutex g_streamMu;
Timer g_streamTimer;
long g_streamLaunches = 0;
// the whole one-vs-many pass of the device-pointer API (asynchronous: its events are resolved later as well)
Timer g_devPassTimer;
long g_devPasses = 0;
double g_devPassFrames = 0.0;

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
  for (int b = 0; b < NIN; ++b) CU(cudaEventCreateWithFlags(&d.inFree[b], cudaEventDisableTiming));
  CU(cudaStreamCreateWithFlags(&d.copyIn, cudaStreamNonBlocking));
  return B200_OK;
}

// ------------------------------------------------------------------ launch helpers
template <int VAR, bool FIT, bool TRI>
int launch_pair_t(const PairArgs& a, dim3 grid, cudaStream_t st) {
  static std::atomic<bool> attr[64];   // per device; several host threads (one per device) may get here at once
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr[dev & 63].load(std::memory_order_acquire)) {
    CU(cudaFuncSetAttribute(pair_kernel<VAR, FIT, TRI>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_BYTES));
    attr[dev & 63].store(true, std::memory_order_release);
  }
  COUNT_LAUNCH();
  pair_kernel<VAR, FIT, TRI><<<grid, PAIR_THREADS, PAIR_SMEM_BYTES, st>>>(a);
  CU(cudaGetLastError());
  return B200_OK;
}

// Lazily initialised knobs are atomics: b200_rms2d_tri drives one host thread per device and they all read them.
std::atomic<int> g_variant{-1};  // MMA shape variant; env B200_MMA_VARIANT overrides (0..3)
int mma_variant() {
  int v = g_variant.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("B200_MMA_VARIANT");
    v = e ? atoi(e) : 3;
    if (v < 0 || v > 3) v = 3;
    g_variant.store(v, std::memory_order_relaxed);
  }
  return v;
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
std::atomic<int> g_engine{-1};
int pair_engine() {
  int v = g_engine.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("B200_PAIR_ENGINE");
    v = e ? atoi(e) : 0;
    if (v < 0 || v > 2) v = 0;
    g_engine.store(v, std::memory_order_relaxed);
  }
  return v;
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

int i8_stats(const I8Set& S, const void* d_crd, size_t stride, const int* d_frameIdx, long srcBase, int f0,
             const int* d_atomIdx, int nAtoms, const double* d_centerMass, const double* d_covMass,
             unsigned int* d_maxBits, cudaStream_t st, int fEnd = -1, bool srcIsDouble = false) {
  if (fEnd < 0) fEnd = S.nFrames;
  I8StatsArgs a;
  a.crd = d_crd; a.stride = stride; a.frameIdx = d_frameIdx; a.srcBase = srcBase; a.nFrames = fEnd; a.f0 = f0;
  a.atomIdx = d_atomIdx; a.nAtoms = nAtoms; a.centerMass = d_centerMass; a.covMass = d_covMass;
  a.centers = S.centers; a.maxAbsBits = d_maxBits;
  const int nb = (fEnd - f0 + 7) / 8;
  if (nb <= 0) return B200_OK;
  COUNT_LAUNCH();
  if (srcIsDouble) i8_stats_kernel<double><<<nb, 256, 0, st>>>(a);
  else i8_stats_kernel<float><<<nb, 256, 0, st>>>(a);
  CU(cudaGetLastError());
  return B200_OK;
}

/// Zero image and G (padding atoms, rows and frames must read as zeros) before the first i8_quant of a set.
int i8_clear(const I8Set& S, cudaStream_t st) {
  CU(cudaMemsetAsync(S.image, 0, i8_image_bytes(S.nRg, S.nC), st));
  CU(cudaMemsetAsync(S.G, 0, (size_t)S.nRg * I8_FR_PER_RG * sizeof(double), st));
  return B200_OK;
}

int i8_quant(const I8Set& S, const void* d_crd, size_t stride, const int* d_frameIdx, long srcBase, int f0,
             const int* d_atomIdx, int nAtoms, const double* d_covMass, int qs, cudaStream_t st, int fEnd = -1,
             bool srcIsDouble = false) {
  if (fEnd < 0) fEnd = S.nFrames;
  I8QuantArgs a;
  a.crd = d_crd; a.stride = stride; a.frameIdx = d_frameIdx; a.srcBase = srcBase; a.nFrames = fEnd; a.f0 = f0;
  a.atomIdx = d_atomIdx; a.nAtoms = nAtoms; a.nC = S.nC; a.covMass = d_covMass; a.centers = S.centers;
  a.scale = std::ldexp(1.0, qs); a.invScale2 = std::ldexp(1.0, -2 * qs);
  a.image = S.image; a.G = S.G;
  const int nb = (fEnd - f0 + 7) / 8;
  if (nb <= 0) return B200_OK;
  COUNT_LAUNCH();
  if (srcIsDouble) i8_quant_kernel<double><<<nb, 256, 0, st>>>(a);
  else i8_quant_kernel<float><<<nb, 256, 0, st>>>(a);
  CU(cudaGetLastError());
  return B200_OK;
}

// Fixed-point scale pinned by the caller (b200_set_fixed_point_bits; 0 = automatic).  One process per GPU: the ranks
// agree on min(qs) among themselves and pin it, so that every shard of one matrix is rounded to the same grid.
std::atomic<int> g_fixedQs{0};
constexpr int I8_MAX_ATOMS = 130000;   // int32 TMEM accumulators: one digit product is <= 2^14, the sum stays exact below 2^17 atoms
                                       // (130,000: the B-digit fold of the epilogue then also fits 48 bits)

/// Meeting point of the per-device host threads of one multi-device call: every thread contributes the extent of the
/// frames it holds and all of them continue with the maximum (device 0 holds every frame), so that all shards use one
/// grid and one engine whatever the device count.
struct ExtentShare {
  std::mutex mu;
  std::condition_variable cv;
  int parties = 1, arrived = 0;
  bool failed = false;
  float mx = 0.f;
  /// \return false when another party failed before the meeting (the caller gives up too)
  bool meet(float mine, float* all) {
    std::unique_lock<std::mutex> lk(mu);
    mx = std::max(mx, mine);
    ++arrived;
    cv.notify_all();
    cv.wait(lk, [&] { return arrived >= parties || failed; });
    *all = mx;
    return !failed;
  }
  void abandon() {   // a party that cannot reach the meeting (error path) releases the others
    std::lock_guard<std::mutex> lk(mu);
    failed = true;
    cv.notify_all();
  }
  void leave() {     // a party with nothing to do (empty shard)
    std::lock_guard<std::mutex> lk(mu);
    ++arrived;
    cv.notify_all();
  }
};
/// Releases the other parties if this one returns before the meeting.
struct ShareGuard {
  ExtentShare* s;
  bool settled = false;
  explicit ShareGuard(ExtentShare* sh) : s(sh) {}
  ~ShareGuard() { if (s && !settled) s->abandon(); }
};

/// qs (fractional bits) for a largest centred, sqrt(m)-scaled coordinate mx; < 0 when none is usable.
inline int i8_bits_for_extent(double mx, double headroom) {
  int q = 30;
  if (mx > 0.0) q = (int)std::floor(std::log2((double)I8_QMAX / (headroom * mx)));
  return std::min(q, 30);
}

/// Reads back (max |coordinate|, total mass) after the stats kernel and picks the number of
/// fractional bits.  *eligible = false when the grid would be too coarse for the contract.
int i8_choose_scale(Device& d, const unsigned int* d_maxBits, const double* d_total, int nAtoms, cudaStream_t st,
                    int* qs, bool* eligible, ExtentShare* share = nullptr) {
  int rc;
  *eligible = false; *qs = 0;
  if ((rc = d.hostScal.reserve(64))) { if (share) share->abandon(); return rc; }
  unsigned int* hBits = (unsigned int*)d.hostScal.p;
  double* hTotal = (double*)((char*)d.hostScal.p + 8);
  cudaError_t ce = cudaMemcpyAsync(hBits, d_maxBits, sizeof(unsigned int), cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(hTotal, d_total, sizeof(double), cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) {
    if (share) share->abandon();
    return fail(B200_ERR_CUDA, "%s:%d: reading back the extent -> %s", __FILE__, __LINE__, cudaGetErrorString(ce));
  }
  float mx;
  std::memcpy(&mx, hBits, 4);
  if (share && !share->meet(std::isfinite(mx) ? mx : INFINITY, &mx)) return fail(B200_ERR_STATE, "another device of this call failed");
  const double total = *hTotal;
  // total mass < SMALL: every RMSD is -1 (src/Frame.cpp:1160-1163); the FP64 engine writes that
  if (!(total >= 1e-14) || !std::isfinite(mx) || nAtoms >= I8_MAX_ATOMS) return B200_OK;
  int q = i8_bits_for_extent((double)mx, 1.0);
  const int pinned = g_fixedQs.load(std::memory_order_relaxed);
  if (pinned > 0) {
    if (pinned > q) return fail(B200_ERR_ARG, "pinned fixed-point scale (%d bits) too fine for this selection's extent (%d bits fit)", pinned, q);
    q = pinned;
  }
  if (q < 0) return B200_OK;
  *qs = q;
  *eligible = i8_worst_error(q, nAtoms, total) <= I8_MAX_WORST_ERROR;
  return B200_OK;
}

// MMA CTA group of the tcgen05 kernel: 2 = CTA pairs (tcgen05.mma.cta_group::2, 28 x 28 tiles), 1 = single CTA
// (14 x 28 tiles).  Env B200_I8_CTA_GROUP or b200_set_i8_cta_group().
std::atomic<int> g_i8Cg{-1};
int i8_cta_group() {
  int v = g_i8Cg.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("B200_I8_CTA_GROUP");
    v = e ? atoi(e) : 2;
    if (v != 1 && v != 2) v = 2;
    g_i8Cg.store(v, std::memory_order_relaxed);
  }
  return v;
}

template <bool TRI, int CG, bool DBG>
int launch_pair_i8_t(const PairI8Args& a, int grid, cudaStream_t st) {
  static std::atomic<bool> attr[64];
  int dev = 0;
  cudaGetDevice(&dev);
  constexpr int smem = i8_smem_bytes<CG>();
  if (!attr[dev & 63].load(std::memory_order_acquire)) {
    CU(cudaFuncSetAttribute(pair_i8_kernel<TRI, CG, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr[dev & 63].store(true, std::memory_order_release);
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
  a.dbgRing = nullptr;
  if (a.dbgMode == 7) {
    static long long* ring = nullptr;
    if (!ring) CU(cudaMalloc(&ring, (size_t)160 * 4 * 10 * 448 * sizeof(long long)));
    a.dbgRing = ring;
  }
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
                Timer* tpack, TriPlan& plan, ExtentShare* share = nullptr) {
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
    if ((rc = i8_choose_scale(d, d_maxBits, d_total, nAtoms, st, &plan.qs, &ok, share))) return rc;
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
    // nofit: a common origin, taken from the LAST frame -- the one frame every shard of a matrix holds, so that all
    // shards (and any device count) subtract the same origin and produce the same bits
    if (!fit) { COUNT_LAUNCH(); shift_kernel<<<1, 1, 0, st>>>(d_crd, stride, d_frameIdx, srcBase, nFrames - 1, d_atomIdx, d_shift); }
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

// Rows [fLo,fHi) of host COORDS (the leading `width` floats of each frame) -> d_dst (row pitch `width`), on stream st.
// Pinned source: one strided DMA.  Pageable source (cpptraj's std::vector<float>): the copy pool gathers <= 16 MiB of
// rows at a time into one of two pinned ring slots, each followed by its own contiguous H2D copy, so the host-side
// gather of piece k+1 overlaps the transfer of piece k (the driver's own pageable path moves ~10 GB/s, this ~50).
int upload_rows(Device& d, float* d_dst, const float* crd, size_t stride, int fLo, int fHi, size_t width, bool pinnedIn,
                cudaStream_t st, double* h2dBytes) {
  const size_t rows = (size_t)std::max(0, fHi - fLo);
  if (!rows) return B200_OK;
  *h2dBytes += (double)(rows * width * sizeof(float));
  const float* src = crd + (size_t)fLo * stride;
  if (pinnedIn) {
    // frames whose selected span is the whole frame are contiguous on both sides: one linear DMA instead of a pitched one
    // (cfg2, 12 MB chunks while downloads are in flight: 43 instead of 37 GB/s, e2e 5.19 -> 4.75 ms; B200_UPLOAD_2D=1: old path)
    static const bool linear = !getenv("B200_UPLOAD_2D");
    if (width == stride && linear)
      CU(cudaMemcpyAsync(d_dst, src, rows * width * sizeof(float), cudaMemcpyHostToDevice, st));
    else
      CU(cudaMemcpy2DAsync(d_dst, width * sizeof(float), src, stride * sizeof(float), width * sizeof(float), rows,
                           cudaMemcpyHostToDevice, st));
    return B200_OK;
  }
  const size_t rowBytes = width * sizeof(float);
  const size_t rpp = std::max<size_t>(1, ((size_t)16 << 20) / rowBytes);
  int rc;
  for (size_t r0 = 0; r0 < rows; r0 += rpp) {
    const size_t nr = std::min(rpp, rows - r0);
    const int b = d.inNext;
    d.inNext = (d.inNext + 1) % NIN;
    if (d.inUsed[b]) CU(cudaEventSynchronize(d.inFree[b]));   // the transfer that last read this slot is done
    if ((rc = d.inStage[b].reserve(std::min(rpp, rows) * rowBytes))) return rc;
    d.pool.copy2d(d.inStage[b].p, rowBytes, src + r0 * stride, stride * sizeof(float), rowBytes, nr);
    CU(cudaMemcpyAsync(d_dst + r0 * width, d.inStage[b].p, nr * rowBytes, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(d.inFree[b], st));
    d.inUsed[b] = true;
  }
  return B200_OK;
}

/// Frames [fLo,fHi) of host COORDS into a (grown) device buffer.
int upload_crd(Device& d, DevBuf& buf, const float* crd, size_t stride, int fLo, int fHi, size_t widthFloats, cudaStream_t st,
               double* h2dBytes) {
  int rc;
  if ((rc = buf.reserve((size_t)std::max(1, fHi - fLo) * widthFloats * sizeof(float)))) return rc;
  return upload_rows(d, (float*)buf.p, crd, stride, fLo, fHi, widthFloats, host_ptr_is_pinned(crd + (size_t)fLo * stride), st, h2dBytes);
}

/// Device copy of frames [fLo,fHi) of host COORDS: the resident copy when the caller announced one
/// (b200_coords_resident_begin: no transfer, rows indexed by absolute frame number), else an upload into `buf`.
/// Frame f is row f - *srcBase of *d_ptr, row pitch *pitch floats.
int get_coords(Device& d, DevBuf& buf, const float* crd, size_t stride, int nFramesTotal, int fLo, int fHi, size_t width,
               cudaStream_t st, double* h2dBytes, const float** d_ptr, size_t* pitch, long* srcBase) {
  if (d.resHost == crd && d.resStride == stride && d.resFrames == nFramesTotal && width <= d.resWidth && d.resBuf.p) {
    *d_ptr = (const float*)d.resBuf.p; *pitch = d.resWidth; *srcBase = 0;
    return B200_OK;
  }
  int rc;
  if ((rc = upload_crd(d, buf, crd, stride, fLo, fHi, width, st, h2dBytes))) return rc;
  *d_ptr = (const float*)buf.p; *pitch = width; *srcBase = (long)fLo;
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

/// Bands of result elements on their way to the caller: slot s holds [base, base+n) of the output array until its
/// download has completed.  A pinned destination receives the DMA directly.  A pageable one (cpptraj's new float[],
/// never touched: every page faults on first write) is filled from the slot's pinned stage by the output copy pool,
/// driven by a delivery thread of its own, so that the thread that stages the uploads and launches the kernels
/// never waits for a page fault: uploads, kernels, downloads and both host-side copies overlap.
struct OutRing {
  Device& d;
  float* out;          // caller's array (element 0 of the whole matrix)
  bool pinned;
  struct Pending { size_t base = 0, n = 0; bool live = false; } pend[NSLOT];
  double d2h = 0.0;
  // delivery thread (pageable destinations only)
  std::thread worker;
  std::mutex mu;
  std::condition_variable cv;
  std::vector<int> queue;      // slots whose download has been queued, in order
  size_t qHead = 0;
  bool quit = false, failed = false;
  // pageable destination: the range this call will fill, populated by the delivery thread's pool before the first
  // band arrives (from the END when `fromEnd`: the pipelined path delivers the last rows first)
  char* popLo = nullptr;
  size_t popBytes = 0;
  bool fromEnd = false;
  OutRing(Device& dev, float* o, bool pin, float* first = nullptr, size_t elts = 0, bool lastRowsFirst = false)
      : d(dev), out(o), pinned(pin), popLo((char*)first), popBytes(elts * sizeof(float)), fromEnd(lastRowsFirst) {
    if (!pinned) {
      if (first) advise_hugepages(first, popBytes);
      worker = std::thread([this] { deliver(); });
    }
  }
  ~OutRing() {
    if (worker.joinable()) {
      { std::lock_guard<std::mutex> lk(mu); quit = true; }
      cv.notify_all();
      worker.join();
    }
  }
  void deliver() {
    cudaSetDevice(d.id);
    if (popLo && popBytes && !getenv("B200_NO_POPULATE")) {
      const size_t piece = (size_t)4 << 20, np = (popBytes + piece - 1) / piece;
      std::atomic<bool> supported{true};
      d.poolOut.run(np, [&](size_t k) {
        if (!supported.load(std::memory_order_relaxed)) return;
        const size_t a = (fromEnd ? np - 1 - k : k) * piece;
        if (!populate_pages(popLo + a, std::min(piece, popBytes - a))) supported.store(false, std::memory_order_relaxed);
      });
    }
    for (;;) {
      int s;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return quit || qHead < queue.size(); });
        if (qHead >= queue.size()) return;   // quit and nothing left
        s = queue[qHead];
      }
      const bool ok = cudaEventSynchronize(d.done[s]) == cudaSuccess;
      if (ok) d.poolOut.copy(out + pend[s].base, d.outStage[s].p, pend[s].n * sizeof(float));
      {
        std::lock_guard<std::mutex> lk(mu);
        if (!ok) failed = true;
        pend[s].live = false;
        ++qHead;
      }
      cv.notify_all();
    }
  }
  int reserve(int s, size_t maxElts) {
    int rc;
    if ((rc = d.outChunk[s].reserve(std::max<size_t>(maxElts, 1) * sizeof(float)))) return rc;
    if (!pinned && (rc = d.outStage[s].reserve(std::max<size_t>(maxElts, 1) * sizeof(float)))) return rc;
    return B200_OK;
  }
  /// Wait until slot s's band has been delivered; the slot may be refilled afterwards.
  int retire(int s) {
    if (pinned) {
      if (!pend[s].live) return B200_OK;
      CU(cudaEventSynchronize(d.done[s]));
      pend[s].live = false;
      return B200_OK;
    }
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return !pend[s].live || failed; });
    if (failed) return fail(B200_ERR_CUDA, "a download of the result failed");
    return B200_OK;
  }
  /// Queue the download of slot s's chunk buffer (n elements, destined for out[base..]) on stream st.
  int download(int s, size_t base, size_t n, cudaStream_t st) {
    if (n) {
      float* dst = pinned ? out + base : (float*)d.outStage[s].p;
      CU(cudaMemcpyAsync(dst, d.outChunk[s].p, n * sizeof(float), cudaMemcpyDeviceToHost, st));
      d2h += (double)n * sizeof(float);
    }
    CU(cudaEventRecord(d.done[s], st));
    if (pinned) { pend[s].base = base; pend[s].n = n; pend[s].live = true; return B200_OK; }
    {
      std::lock_guard<std::mutex> lk(mu);
      pend[s].base = base; pend[s].n = n; pend[s].live = true;
      queue.push_back(s);
    }
    cv.notify_all();
    return B200_OK;
  }
  int drain() {
    int rc;
    for (int s = 0; s < NSLOT; ++s) if ((rc = retire(s))) return rc;
    return B200_OK;
  }
};

// Pipelined host path of the tcgen05 engine (no frame list).  Frames are uploaded from the END of the trajectory
// towards the start, chunk by chunk.  Rows [a, b) of the triangle pair only with frames >= a, so as soon as a chunk
// has landed and is quantised its band of rows can be computed and its slice of the triangle copied back while the
// next chunk uploads: H2D, compute and D2H overlap (PCIe is full duplex) instead of upload -> compute/download.
// Pageable COORDS and result buffers (what cpptraj passes) go through pinned ring slots and the copy pool.
// The fixed-point scale must be known before the first chunk is quantised: it is taken from the extent of the TOP
// chunk of the trajectory -- the same frames whatever the shard, so all shards of a matrix agree on it -- with 25 %
// headroom (or it is the scale pinned by b200_set_fixed_point_bits); every later chunk keeps feeding the running
// maximum, and if at the end the headroom turned out too small (or the scale leaves too few fractional bits) *done
// stays false and the caller runs the two-pass path.
int host_tri_pipelined_i8(Device& d, const float* crd, size_t stride, int nFrames, const int* atomIdx, int nAtoms,
                          const double* mass, int row0, int row1, float* outTri, bool* done) {
  *done = false;
  int maxAtom = 0, rc;
  if ((rc = validate_sel(atomIdx, nAtoms, stride, &maxAtom))) return rc;
  if (nAtoms >= I8_MAX_ATOMS) return B200_OK;
  double total = (double)nAtoms;
  if (mass) { total = 0.0; for (int k = 0; k < nAtoms; ++k) total += mass[k]; }
  if (!(total >= 1e-14)) return B200_OK;      // (every RMSD is -1: the FP64 engine writes that)
  const size_t F = (size_t)nFrames;
  const size_t width = (size_t)3 * ((size_t)maxAtom + 1);
  const int f0 = (row0 / I8_TILE_J) * I8_TILE_J;           // first frame needed (tile aligned)
  // chunk = band: <= 40 MB of output, a multiple of the 28-frame tile
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
  // Chunk list, last chunk of the trajectory first.  The top chunk [T0, F) is a function of F alone; below it the
  // chunks grow from small ones at the shard's first frame (nothing overlaps the first upload and the last download).
  struct Chunk { int fa, fb; };
  std::vector<Chunk> chunks;
  {
    std::vector<int> bnd;   // ascending boundaries from f0
    const int small = std::max(I8_TILE_J, C / 4 / I8_TILE_J * I8_TILE_J);
    const int T0 = (nFrames - f0 > 2 * small) ? (nFrames - small) / I8_TILE_J * I8_TILE_J : f0;
    int at = f0, step = small;
    while (at < T0) {
      bnd.push_back(at);
      at += step;
      if (step < C) step = std::min(C, step + small);
    }
    bnd.push_back(T0);
    bnd.push_back(nFrames);
    for (size_t i = bnd.size() - 1; i > 0; --i)
      if (bnd[i] > bnd[i - 1]) chunks.push_back({bnd[i - 1], bnd[i]});
  }
  size_t maxChunk = 0;
  for (const Chunk& c : chunks) {
    const int lo = std::max(c.fa, row0), hi = std::min(c.fb, row1);
    if (hi > lo) maxChunk = std::max(maxChunk, tri_row_start(F, hi) - tri_row_start(F, lo));
  }
  float* outFirst = outTri + tri_row_start(F, (size_t)row0);   // (the shard's own range: the base may lie outside the caller's buffer)
  const size_t outElts = tri_row_start(F, (size_t)row1) - tri_row_start(F, (size_t)row0);
  const bool pinnedIn = host_ptr_is_pinned(crd + (size_t)f0 * stride);
  OutRing ring(d, outTri, host_ptr_is_pinned(outFirst), outFirst, outElts, true);
  // (ring slots are reserved at first use: in a fresh process pinning slot k+1 overlaps the work queued for slot k)
  std::vector<cudaEvent_t> evs;
  cudaStream_t sCopy = d.copyIn;
  // (every exit: the uploads queued ahead have left the caller's buffer and the device copy is quiescent)
  auto cleanup = [&]() { cudaStreamSynchronize(sCopy); for (cudaEvent_t e : evs) cudaEventDestroy(e); evs.clear(); };
  const float* d_crd = (const float*)d.crd.p;
  const long srcBase = (long)f0;      // row r of the device copy is frame f0 + r
  Timer tpair;
  double h2d = (double)nAtoms * 4 + (mass ? (double)nAtoms * 8 : 0);
  int qs = g_fixedQs.load(std::memory_order_relaxed), band = 0;
  const bool pinnedScale = qs > 0;
  if (pinnedScale && i8_worst_error(qs, nAtoms, total) > I8_MAX_WORST_ERROR) return B200_OK;
  long nLaunch = 0;
  // B200_PIPE_TRACE=1: a time line of the pipeline on stderr (events after every upload, band kernel and download)
  static const bool trace = getenv("B200_PIPE_TRACE") != nullptr;
  struct TraceEv { cudaEvent_t e; int chunk; char what; };
  std::vector<TraceEv> tev;
  cudaEvent_t tStart = nullptr;
  auto mark = [&](cudaStream_t strm, int chunk, char what) {
    if (!trace) return;
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, strm); tev.push_back({e, chunk, what});
  };
  if (trace) { cudaEventCreate(&tStart); cudaEventRecord(tStart, sIn); }
  // Uploads go through a stream of their own, chunk after chunk without a kernel in between (on the stream that also
  // carries the statistics / quantisation kernels the next copy waited for those, and those for SMs the band kernels
  // hold: measured, the 120 MB of cfg2 took 4.1 ms instead of 2.2).  Pinned COORDS: every copy is queued up front;
  // pageable: the copy pool stages one chunk ahead of the kernels.
  std::vector<cudaEvent_t> evUp(chunks.size(), nullptr);
  size_t nextUp = 0;
  {   // order after whatever the caller's earlier calls left on sIn (selection and mass uploads above are on sIn)
    cudaEvent_t e0;
    CU(cudaEventCreateWithFlags(&e0, cudaEventDisableTiming));
    evs.push_back(e0);
    CU(cudaEventRecord(e0, sIn));
    CU(cudaStreamWaitEvent(sCopy, e0, 0));
  }
  auto enqueue_upload = [&](size_t k) -> int {
    const int fa = chunks[k].fa, fb = chunks[k].fb;
    int r = upload_rows(d, (float*)d.crd.p + (size_t)(fa - f0) * width, crd, stride, fa, fb, width, pinnedIn, sCopy, &h2d);
    if (r) return r;
    CU(cudaEventCreateWithFlags(&evUp[k], cudaEventDisableTiming));
    evs.push_back(evUp[k]);
    CU(cudaEventRecord(evUp[k], sCopy));
    mark(sCopy, (int)k, 'U');
    return B200_OK;
  };
  for (size_t k = 0; k < chunks.size(); ++k) {
    const int fa = chunks[k].fa, fb = chunks[k].fb;
    const size_t upTo = pinnedIn ? chunks.size() : std::min(chunks.size(), k + 2);
    for (; nextUp < upTo; ++nextUp)
      if ((rc = enqueue_upload(nextUp))) { cleanup(); return rc; }
    CU(cudaStreamWaitEvent(sIn, evUp[k], 0));
    if ((rc = i8_stats(q, d_crd, width, nullptr, srcBase, fa, (const int*)d.idxA.p, nAtoms, d_mass, d_mass, d_maxBits, sIn, fb))) { cleanup(); return rc; }
    if (k == 0 && !pinnedScale) {
      // scale from the top chunk, with headroom for the frames not seen yet
      if ((rc = d.hostScal.reserve(64))) { cleanup(); return rc; }
      unsigned int* hBits = (unsigned int*)d.hostScal.p;
      CU(cudaMemcpyAsync(hBits, d_maxBits, sizeof(unsigned int), cudaMemcpyDeviceToHost, sIn));
      CU(cudaStreamSynchronize(sIn));
      float mx;
      std::memcpy(&mx, hBits, 4);
      if (!std::isfinite(mx) || !(mx > 0.f)) { cleanup(); return B200_OK; }
      qs = i8_bits_for_extent((double)mx, 1.25);
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
    if ((rc = ring.retire(s)) || (rc = ring.reserve(s, maxChunk))) { cleanup(); return rc; }
    cudaStream_t st = d.stream[s];
    CU(cudaStreamWaitEvent(st, ev, 0));
    const size_t base = tri_row_start(F, lo), n = tri_row_start(F, hi) - base;
    tpair.begin(st);
    if ((rc = run_pair_i8_band(d, q, q, lo, hi, true, qs, d_total, (float*)d.outChunk[s].p, base, 0, nullptr, st))) { cleanup(); return rc; }
    tpair.end(st);
    mark(st, (int)k, 'K');
    ++nLaunch;
    if ((rc = ring.download(s, base, n, st))) { cleanup(); return rc; }
    mark(st, (int)k, 'D');
  }
  // did the scale hold for every frame?
  {
    if ((rc = d.hostScal.reserve(64))) { cleanup(); return rc; }
    unsigned int* hBits = (unsigned int*)d.hostScal.p;
    CU(cudaMemcpyAsync(hBits, d_maxBits, sizeof(unsigned int), cudaMemcpyDeviceToHost, sIn));
    CU(cudaStreamSynchronize(sIn));
    CU(cudaStreamSynchronize(sCopy));
    if ((rc = ring.drain())) { cleanup(); return rc; }
    for (int s = 1; s < NSLOT; ++s) CU(cudaStreamSynchronize(d.stream[s]));
    cleanup();
    if (trace) {
      fprintf(stderr, "pipeline trace (ms after the first enqueue; U upload done, K band kernel done, D download done; rows of the chunk):\n");
      for (const TraceEv& t : tev) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, tStart, t.e);
        fprintf(stderr, "  %c chunk %2d [%5d,%5d) %.3f\n", t.what, t.chunk, chunks[t.chunk].fa, chunks[t.chunk].fb, ms);
        cudaEventDestroy(t.e);
      }
      cudaEventDestroy(tStart);
    }
    float mx;
    std::memcpy(&mx, hBits, 4);
    if (!((double)mx * std::ldexp(1.0, qs) <= (double)I8_QMAX)) {
      if (pinnedScale) return fail(B200_ERR_ARG, "pinned fixed-point scale (%d bits) too fine for this selection's extent", qs);
      return B200_OK;   // *done == false: the two-pass path recomputes
    }
  }
  const double pairs = (double)outElts;
  add_stats(0.0, 1, tpair.resolve(), nLaunch, pairs, h2d, ring.d2h);
  g_lastEngine.store(2); g_lastQs.store(qs);
  *done = true;
  return B200_OK;
}

// One shard of the triangle on one device, host buffers.  `share`: meeting point of the device threads of a
// multi-device call (nullable).
int host_tri_on_device(Device& d, const float* crd, size_t stride, int nFramesTotal, const int* frameIdx, int nFrames,
                       const int* atomIdx, int nAtoms, const double* mass, int fit, int row0, int row1, float* outTri,
                       ExtentShare* share = nullptr) {
  ShareGuard guard(share);
  CU(cudaSetDevice(d.id));
  if (row1 <= row0 || nFrames < 2) {
    if (share) { share->leave(); guard.settled = true; }
    return B200_OK;
  }
  int maxAtom = 0, rc;
  if ((rc = validate_sel(atomIdx, nAtoms, stride, &maxAtom))) return rc;
  if (frameIdx) {   // a frame list that lists every frame in order (cluster without a sieve) is no list
    bool identity = nFrames <= nFramesTotal;
    for (int f = 0; identity && f < nFrames; ++f) identity = frameIdx[f] == f;
    if (identity) frameIdx = nullptr;
  }
  bool pipelineTried = false;
  {
    const char* e = getenv("B200_HOST_PIPELINE");
    if (fit && !frameIdx && pair_engine() != 1 && nFrames >= 1024 && nFrames <= nFramesTotal && !(e && atoi(e) == 0)) {
      bool done = false;
      pipelineTried = true;
      if ((rc = host_tri_pipelined_i8(d, crd, stride, nFrames, atomIdx, nAtoms, mass, row0, row1, outTri, &done))) return rc;
      if (done) {
        if (share) { share->leave(); guard.settled = true; }
        return B200_OK;
      }
    }
  }
  // (every device thread of a call takes the same decisions up to here; one that falls back after a failed pipelined
  //  attempt must not wait for the others, which may have succeeded: it leaves the meeting and chooses on its own)
  if (share && pipelineTried) { share->leave(); guard.settled = true; share = nullptr; }
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
  double h2d = 0.0;
  const size_t width = (size_t)3 * ((size_t)maxAtom + 1);
  if ((rc = d.crd.reserve((size_t)(sHi - sLo) * width * sizeof(float)))) return rc;
  if ((rc = upload_rows(d, (float*)d.crd.p, crd, stride, sLo, sHi, width, host_ptr_is_pinned(crd + (size_t)sLo * stride), st0, &h2d))) return rc;
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
                        fit, row0, st0, &tpack, plan, share))) return rc;
  guard.settled = true;   // (met, or nobody meets: fit == 0 / FP64 engine forced)
  CU(cudaEventRecord(d.done[0], st0));
  for (int s = 1; s < NSLOT; ++s) CU(cudaStreamWaitEvent(d.stream[s], d.done[0], 0));

  // band size: <= ~40 MB of output per band, multiple of 32 rows
  const size_t F = (size_t)nFrames;
  int bandRows = (int)std::min<size_t>(2048, std::max<size_t>(ROWG, ((size_t)40 << 20) / (4 * F) / ROWG * ROWG));
  float* outFirst = outTri + tri_row_start(F, (size_t)row0);   // (the shard's own range: the base may lie outside the caller's buffer)
  const size_t outElts = tri_row_start(F, (size_t)row1) - tri_row_start(F, (size_t)row0);
  OutRing ring(d, outTri, host_ptr_is_pinned(outFirst), outFirst, outElts, false);
  size_t maxChunk = 0;
  for (int i0 = row0; i0 < row1; i0 += bandRows) {
    const int i1 = std::min(row1, i0 + bandRows);
    maxChunk = std::max(maxChunk, tri_row_start(F, i1) - tri_row_start(F, i0));
  }
  int band = 0;
  long nLaunch = 0;
  for (int i0 = row0; i0 < row1; i0 += bandRows, ++band) {
    const int s = band % NSLOT;
    if ((rc = ring.retire(s)) || (rc = ring.reserve(s, maxChunk))) return rc;
    const int i1 = std::min(row1, i0 + bandRows);
    const size_t base = tri_row_start(F, i0), n = tri_row_start(F, i1) - base;
    cudaStream_t st = d.stream[s];
    tpair.begin(st);
    if ((rc = run_tri_band(d, plan, i0, i1, fit != 0, (float*)d.outChunk[s].p, base, st))) return rc;
    tpair.end(st);
    ++nLaunch;
    if ((rc = ring.download(s, base, n, st))) return rc;
  }
  if ((rc = ring.drain())) return rc;
  for (int s = 0; s < NSLOT; ++s) CU(cudaStreamSynchronize(d.stream[s]));
  add_stats(tpack.resolve(), 1, tpair.resolve(), nLaunch, (double)outElts, h2d, ring.d2h);
  return B200_OK;
}

/// Runs fn(deviceIndex) on one host thread per initialised device (on the caller's thread when there is one device).
template <typename F>
int for_each_device(F&& fn) {
  const int nd = (int)g_devs.size();
  if (nd == 1) return fn(0);
  std::vector<int> rcs(nd, 0);
  std::vector<std::thread> th;
  for (int i = 0; i < nd; ++i) th.emplace_back([&, i]() { rcs[i] = fn(i); });
  for (auto& t : th) t.join();
  cudaSetDevice(g_devs[0].id);
  for (int r : rcs) if (r) return r;
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
  for (Device& d : g_devs) d.destroy();
  g_devs.clear();
  g_devs.resize(want);
  for (int i = 0; i < want; ++i) {
    int rc = init_device(g_devs[i], ids[i]);
    if (rc) { for (Device& d : g_devs) d.destroy(); g_devs.clear(); g_inited = false; return rc; }
  }
  for (Device& d : g_devs) {
    // fresh pages make the result side the heavier one: it gets the larger share of the host threads
    const int nt = host_threads_per_device(want);
    d.pool.ensure(std::max(1, nt * 3 / 8));
    d.poolOut.ensure(std::max(1, nt - nt * 3 / 8));
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

int b200_warmup(void) {
  // What the first call of a process would otherwise pay for inside its timed region and that costs little here: the
  // device chunk buffers and the lazy loading of the main kernels' code.  (Pinning the ~190 MB of staging slots ahead
  // of time was measured too: page-locking runs at ~0.3 GB/s on the B200 box, 0.7 s -- longer than cpptraj needs to
  // read cfg2's trajectory; reserved at first use it overlaps the work already queued.)  Meant for a background
  // thread at command set-up; everything here is grow-only and reused by the calls that follow.
  std::lock_guard<std::mutex> lk(g_mu);
  int rc;
  if ((rc = ensure_init_locked())) return rc;
  for (Device& d : g_devs) {
    CU(cudaSetDevice(d.id));
    for (int s = 1; s < NSLOT; ++s)
      if ((rc = d.outChunk[s].reserve((size_t)40 << 20))) return rc;
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, pair_i8_kernel<true, 2, false>));
    CU(cudaFuncGetAttributes(&fa, pair_i8_kernel<false, 2, false>));
    CU(cudaFuncGetAttributes(&fa, i8_stats_kernel<float>));
    CU(cudaFuncGetAttributes(&fa, i8_quant_kernel<float>));
    CU(cudaFuncGetAttributes(&fa, onevn_stream2_kernel<float>));
    CU(cudaFuncGetAttributes(&fa, onevn_stream2_kernel<double>));
  }
  cudaSetDevice(g_devs[0].id);
  return B200_OK;
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
  for (Device& d : g_devs) d.destroy();
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
int b200_set_mma_variant(int v) { if (v < 0 || v > 3) return fail(B200_ERR_ARG, "variant must be 0..3"); g_variant.store(v); return B200_OK; }
void b200_reset_stats(void) {
  {
    std::lock_guard<std::mutex> tl(g_streamMu);
    if (!g_streamTimer.ev.empty() || !g_devPassTimer.ev.empty()) { cudaDeviceSynchronize(); g_streamTimer.resolve(); g_devPassTimer.resolve(); }
    g_streamLaunches = 0; g_devPasses = 0; g_devPassFrames = 0.0;
  }
  std::lock_guard<std::mutex> lk(g_statMu);
  g_stats = b200_stats();
  g_launches.store(0);
}
void b200_get_stats(b200_stats* out) {
  if (!out) return;
  {
    std::lock_guard<std::mutex> tl(g_streamMu);
    if (!g_streamTimer.ev.empty() || !g_devPassTimer.ev.empty()) {
      cudaDeviceSynchronize();   // (the one-vs-many kernels run on the current device's streams)
      const double ms = g_streamTimer.resolve(), msPass = g_devPassTimer.resolve();
      std::lock_guard<std::mutex> lk(g_statMu);
      g_stats.onevn_stream_ms += ms; g_stats.onevn_stream_launches += g_streamLaunches;
      g_stats.onevn_ms += msPass; g_stats.onevn_launches += g_devPasses; g_stats.frames_1vN += g_devPassFrames;
      g_streamLaunches = 0; g_devPasses = 0; g_devPassFrames = 0.0;
    }
  }
  std::lock_guard<std::mutex> lk(g_statMu);
  *out = g_stats;
  out->kernel_launches = g_launches.load();
}

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
  // one host thread per device; shards are disjoint contiguous ranges of outTri; the threads agree on one
  // fixed-point grid (and engine) for the whole matrix
  ExtentShare share;
  share.parties = nd;
  const long long t0 = g_allocTrace.on ? AllocTrace::now() : 0;
  rc = for_each_device([&](int i) -> int {
    int r0 = 0, r1 = 0;
    int r = shard_rows(nFrames, i, nd, &r0, &r1);
    if (r) { if (nd > 1) share.abandon(); return r; }
    return host_tri_on_device(g_devs[i], crd, frameStrideFloats, nFramesTotal, frameIdx, nFrames, atomIdx, nAtoms, mass, fit,
                              r0, r1, outTri, nd > 1 ? &share : nullptr);
  });
  g_allocTrace.report("b200_rms2d_tri", g_allocTrace.on ? AllocTrace::now() - t0 : 0);
  return rc;
}

}  // extern "C"

namespace {
/// Target rows [t0,t1) of the full matrix on one device.
int host_full_on_device(Device& d, const float* crdTgt, size_t strideTgt, int nTgt, const int* atomIdxTgt, const float* crdRef,
                        size_t strideRef, int nRef, const int* atomIdxRef, int nAtoms, const double* massTgt,
                        const double* massRefCentering, int fit, int t0, int t1, float* outFull, ExtentShare* share) {
  ShareGuard guard(share);
  CU(cudaSetDevice(d.id));
  if (t1 <= t0) {
    if (share) { share->leave(); guard.settled = true; }
    return B200_OK;
  }
  int maxT = 0, maxR = 0, rc;
  if ((rc = validate_sel(atomIdxTgt, nAtoms, strideTgt, &maxT))) return rc;
  if ((rc = validate_sel(atomIdxRef, nAtoms, strideRef, &maxR))) return rc;
  cudaStream_t st0 = d.stream[0];
  double h2d = 0.0;
  const size_t wT = (size_t)3 * (maxT + 1), wR = (size_t)3 * (maxR + 1);
  const int nT = t1 - t0;   // this device's targets: frames t0.. of crdTgt are rows 0.. of the device copy
  if ((rc = upload_crd(d, d.crd, crdTgt, strideTgt, t0, t1, wT, st0, &h2d))) return rc;
  if ((rc = upload_crd(d, d.crdB, crdRef, strideRef, 0, nRef, wR, st0, &h2d))) return rc;
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
    if ((rc = i8_reserve(qA, d.imgA, d.GA, d.cenA, nT, nAtoms))) return rc;
    if ((rc = i8_reserve(qB, d.imgB, d.GB, d.cenB, nRef, nAtoms))) return rc;
    CU(cudaMemsetAsync(d_maxBits, 0, sizeof(unsigned int), st0));
    if ((rc = i8_stats(qA, (const float*)d.crd.p, wT, nullptr, 0, 0, (const int*)d.idxA.p, nAtoms, dmT, dmT, d_maxBits, st0))) return rc;
    if ((rc = i8_stats(qB, (const float*)d.crdB.p, wR, nullptr, 0, 0, (const int*)d.idxB.p, nAtoms, dmR, dmT, d_maxBits, st0))) return rc;
    if ((rc = i8_choose_scale(d, d_maxBits, d_total, nAtoms, st0, &qs, &useI8, share))) return rc;
    if (useI8) {
      if ((rc = i8_clear(qA, st0)) || (rc = i8_clear(qB, st0))) return rc;
      if ((rc = i8_quant(qA, (const float*)d.crd.p, wT, nullptr, 0, 0, (const int*)d.idxA.p, nAtoms, dmT, qs, st0))) return rc;
      if ((rc = i8_quant(qB, (const float*)d.crdB.p, wR, nullptr, 0, 0, (const int*)d.idxB.p, nAtoms, dmT, qs, st0))) return rc;
    } else if (engine == 2) {
      return fail(B200_ERR_ARG, "tcgen05 int8 engine forced but the selection does not qualify (%d fractional bits, %d atoms)", qs, nAtoms);
    }
  } else if (engine == 2) {
    return fail(B200_ERR_ARG, "tcgen05 int8 engine forced but nofit RMSD runs on the FP64 engine only");
  }
  guard.settled = true;
  if (!useI8) {
    A.nFrames = nT; A.Fpad = round_up(nT, ROWG); A.Kpad = round_up(nAtoms, KC);
    B.nFrames = nRef; B.Fpad = round_up(nRef, ROWG); B.Kpad = A.Kpad;
    if ((rc = d.planesA.reserve(plane_doubles(A.Fpad, A.Kpad) * sizeof(double)))) return rc;
    if ((rc = d.planesB.reserve(plane_doubles(B.Fpad, B.Kpad) * sizeof(double)))) return rc;
    if ((rc = d.GA.reserve((size_t)A.Fpad * sizeof(double)))) return rc;
    if ((rc = d.GB.reserve((size_t)B.Fpad * sizeof(double)))) return rc;
    A.planes = (double*)d.planesA.p; A.G = (double*)d.GA.p;
    B.planes = (double*)d.planesB.p; B.G = (double*)d.GB.p;
    // nofit: a common origin for both sets (the first selected atom of the first REFERENCE frame: every device has it)
    if (!fit) { COUNT_LAUNCH(); shift_kernel<<<1, 1, 0, st0>>>((const float*)d.crdB.p, wR, nullptr, 0, 0, (const int*)d.idxB.p, d_shift); }
    if ((rc = run_pack((const float*)d.crd.p, wT, nullptr, 0, nT, 0, (const int*)d.idxA.p, nAtoms, dmT, dmT, d_shift, fit, A, st0))) return rc;
    if ((rc = run_pack((const float*)d.crdB.p, wR, nullptr, 0, nRef, 0, (const int*)d.idxB.p, nAtoms, dmR, dmT, d_shift, fit, B, st0))) return rc;
  }
  tpack.end(st0);
  g_lastEngine.store(useI8 ? 2 : 1); g_lastQs.store(useI8 ? qs : 0);
  // rows (targets) in bands; each band is a contiguous slab of outFull
  const size_t ld = (size_t)nRef;
  int bandRows = (int)std::min<size_t>(1024, std::max<size_t>(ROWG, ((size_t)32 << 20) / (4 * ld) / ROWG * ROWG));
  float* outFirst = outFull + (size_t)t0 * ld;
  OutRing ring(d, outFull, host_ptr_is_pinned(outFirst), outFirst, (size_t)nT * ld, false);
  const size_t maxChunk = (size_t)std::min(bandRows, nT) * ld;
  CU(cudaEventRecord(d.done[0], st0));
  for (int s = 1; s < NSLOT; ++s) CU(cudaStreamWaitEvent(d.stream[s], d.done[0], 0));
  for (int s = 0; s < NSLOT; ++s) if ((rc = ring.reserve(s, maxChunk))) return rc;
  int band = 0; long nLaunch = 0;
  for (int i0 = 0; i0 < nT; i0 += bandRows, ++band) {
    const int s = band % NSLOT;
    if ((rc = ring.retire(s))) return rc;
    const int i1 = std::min(nT, i0 + bandRows);
    cudaStream_t st = d.stream[s];
    const size_t base = (size_t)i0 * ld, n = (size_t)(i1 - i0) * ld;
    tpair.begin(st);
    // kernel indexes out[i*ld + j] with i the device-local target row; shift the pointer so row i0 lands at the chunk start
    if (useI8)
      rc = run_pair_i8_band(d, qA, qB, i0, i1, false, qs, d_total, (float*)d.outChunk[s].p - base, 0, ld, nullptr, st);
    else
      rc = run_pair_band(A, B, i0 / ROWG, (i1 + ROWG - 1) / ROWG - i0 / ROWG, false, fit != 0, d_total,
                         (float*)d.outChunk[s].p - base, 0, ld, st);
    if (rc) return rc;
    tpair.end(st);
    ++nLaunch;
    if ((rc = ring.download(s, (size_t)t0 * ld + base, n, st))) return rc;
  }
  if ((rc = ring.drain())) return rc;
  for (int s = 0; s < NSLOT; ++s) CU(cudaStreamSynchronize(d.stream[s]));
  add_stats(tpack.resolve(), 2, tpair.resolve(), nLaunch, (double)nT * (double)nRef, h2d, ring.d2h);
  return B200_OK;
}
}  // namespace

extern "C" {

int b200_rms2d_full(const float* crdTgt, size_t strideTgt, int nTgt, const int* atomIdxTgt, const float* crdRef,
                    size_t strideRef, int nRef, const int* atomIdxRef, int nAtoms, const double* massTgt,
                    const double* massRefCentering, int fit, float* outFull) {
  if (!crdTgt || !crdRef || !outFull) return fail(B200_ERR_ARG, "null buffer");
  if (nTgt <= 0 || nRef <= 0) return B200_OK;
  std::lock_guard<std::mutex> lk(g_mu);
  int rc;
  if ((rc = ensure_init_locked())) return rc;
  // target rows in equal, 32-row aligned shares over the devices (every pair costs the same here)
  const int nd = (int)g_devs.size();
  const int per = round_up((nTgt + nd - 1) / nd, ROWG);
  ExtentShare share;
  share.parties = nd;
  return for_each_device([&](int i) -> int {
    const int t0 = std::min(nTgt, i * per), t1 = std::min(nTgt, (i + 1) * per);
    return host_full_on_device(g_devs[i], crdTgt, strideTgt, nTgt, atomIdxTgt, crdRef, strideRef, nRef, atomIdxRef, nAtoms,
                               massTgt, massRefCentering, fit, t0, t1, outFull, nd > 1 ? &share : nullptr);
  });
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
  for (Device& x : g_devs) if (x.id == dev) d = &x;
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
// A streaming handle spreads its chunks round-robin over the initialised devices ("lanes"): every lane has its own
// stream, input ring and result arrays; a segment list remembers where the results of each pushed chunk live so that
// flush() returns them in push order (DataSet_double::Add is append-only, src/DataSet_double.cpp:14-20).
constexpr int ONEVN_NIN = 2;   // input ring slots per lane
struct OneVNLane {
  Device* dev = nullptr;
  DevBuf refw, refsum, ref, mass, idx, in[ONEVN_NIN], rms, rot, trans, ws;
  PinBuf stage[ONEVN_NIN];
  cudaStream_t st = nullptr;         // single in-order stream (+ events for slot reuse)
  cudaEvent_t slotFree[ONEVN_NIN] = {};
  cudaEvent_t copied = nullptr;      // last H2D copy that read the caller's (pinned) buffer
  bool touched = false;
  int slot = 0;
  long cap = 0, used = 0;            // result slots allocated / holding unflushed results
  Timer timer;
  long launches = 0;
};
struct OneVNSegment { int lane; long laneOff, n; };
struct b200_1vN {
  int nAtoms = 0, fit = 1, wantRot = 0, maxAtom = 0;
  bool identity = false, hasMass = false;
  std::vector<int> atomIdx;
  std::vector<std::unique_ptr<OneVNLane>> lanes;
  std::vector<OneVNSegment> segs;    // unflushed chunks in push order
  int nextLane = 0;
  long pushed = 0, flushed = 0;
  long bestFrame = -1;
  double bestVal = 0.0;
  double h2d = 0.0, d2h = 0.0;
};

static void onevn_destroy(b200_1vN* h) {
  if (!h) return;
  for (auto& lp : h->lanes) {
    OneVNLane& L = *lp;
    if (L.dev) cudaSetDevice(L.dev->id);
    if (L.st) { cudaStreamSynchronize(L.st); cudaStreamDestroy(L.st); }
    for (int s = 0; s < ONEVN_NIN; ++s) { if (L.slotFree[s]) cudaEventDestroy(L.slotFree[s]); L.in[s].release(); L.stage[s].release(); }
    if (L.copied) cudaEventDestroy(L.copied);
    DevBuf* all[] = {&L.refw, &L.refsum, &L.ref, &L.mass, &L.idx, &L.rms, &L.rot, &L.trans, &L.ws};
    for (DevBuf* b : all) b->release();
    L.timer.resolve();
  }
  if (!g_devs.empty()) cudaSetDevice(g_devs[0].id);
  delete h;
}

/// Room for `need` unflushed results on a lane (the live ones are kept).
static int onevn_grow(b200_1vN* h, OneVNLane& L, long need) {
  if (need <= L.cap) return B200_OK;
  long ncap = std::max<long>(need, std::max<long>(L.cap * 2, 1 << 16));
  DevBuf nr, no, nt;
  int rc;
  if ((rc = nr.reserve((size_t)ncap * 8))) return rc;
  if (h->wantRot) {
    if ((rc = no.reserve((size_t)ncap * 72))) { nr.release(); return rc; }
    if ((rc = nt.reserve((size_t)ncap * 24))) { nr.release(); no.release(); return rc; }
  }
  if (L.used > 0) {
    CU(cudaMemcpyAsync(nr.p, L.rms.p, (size_t)L.used * 8, cudaMemcpyDeviceToDevice, L.st));
    if (h->wantRot) {
      CU(cudaMemcpyAsync(no.p, L.rot.p, (size_t)L.used * 72, cudaMemcpyDeviceToDevice, L.st));
      CU(cudaMemcpyAsync(nt.p, L.trans.p, (size_t)L.used * 24, cudaMemcpyDeviceToDevice, L.st));
    }
  }
  CU(cudaStreamSynchronize(L.st));
  L.rms.release(); L.rot.release(); L.trans.release();
  L.rms = nr; L.rot = no; L.trans = nt;
  L.cap = ncap;
  return B200_OK;
}

/// Launches the one-vs-many kernels for `nFrames` frames at d_crd on stream st: chunk table, a TMA streaming kernel +
/// per-frame finish (sorted selections, 16-byte aligned base), then the general gather kernel, which returns at once
/// when a streaming variant did the work (the choice is made on the device: no host round trip).
/// Fitted RMSD of a dense enough selection: streaming variant 2 (chunk-major, partial records per frame and part);
/// no-fit, or selections so sparse that the partial records would rival the frame data: variant 1.
/// refw: 4 N doubles (rx, ry, rz, m) followed by 4 N doubles (m rx, m ry, m rz, m), as onevn_setup_kernel writes them.
/// ws: workspace of at least onevn_ws_bytes() bytes.
static size_t onevn_max_chunks(size_t stride, int chunkAtoms) { return std::min<size_t>(stride / 3 / (size_t)chunkAtoms + 2, 1u << 20); }
template <typename T>
static bool onevn_use_v2(size_t stride, int nAtoms, int fit) {
  if (!fit) return false;
  if (const char* e = getenv("B200_1VN_V2")) return atoi(e) != 0;
  const size_t parts = onevn_max_chunks(stride, ONEVN2_CHUNK_BYTES / (3 * (int)sizeof(T)));   // one record per frame and chunk
  return parts * 128 <= (size_t)nAtoms * 3 * sizeof(T) / 4 && parts <= 256;   // partial records (128 B each) <= a quarter of the bytes read per frame
}
static size_t onevn_ws_bytes(size_t stride, int nFrames) {
  const size_t maxChunks = onevn_max_chunks(stride, ONEVN2_CHUNK_BYTES / 24);   // (double frames: the smaller chunk)
  const size_t head = 64 + (maxChunks + 1) * sizeof(int) + 64 + 8 * maxChunks * sizeof(double) + 64;
  // variant 2 is only chosen when its 2 * maxChunks records per frame stay small (onevn_use_v2): bound them by 256 parts
  const size_t parts = std::min<size_t>(2 * maxChunks, 256);
  return head + (size_t)nFrames * std::max<size_t>(ONEVN_REC, parts * 16) * sizeof(double);
}
template <typename T>
static int onevn_run(int numSMs, const void* d_crd, size_t stride, int nFrames, const int* d_atomIdx, int nAtoms,
                     const double* refw, const double* refsum, int fit, double* rmsd, double* rot, double* trans,
                     void* ws, cudaStream_t st, const int* d_frameIdx = nullptr, long srcBase = 0,
                     const double* setupRef = nullptr, const double* setupMass = nullptr) {
  // setupRef != nullptr: refw / refsum are still to be derived from this reference (3 nAtoms doubles) and these masses;
  // that happens inside the one preparation launch below
  const bool v2 = onevn_use_v2<T>(stride, nAtoms, fit);
  const int APC = v2 ? ONEVN2_CHUNK_BYTES / (3 * (int)sizeof(T)) : ONEVN_S_CHUNK_BYTES / (3 * (int)sizeof(T));
  const int maxChunks = (int)onevn_max_chunks(stride, APC);
  int* hdr = (int*)ws;
  int* kLo = hdr + 16;
  double* partSum = (double*)((char*)ws + ((64 + (size_t)(maxChunks + 1) * sizeof(int) + 63) & ~(size_t)63));
  double* rec = partSum + 8 * (size_t)maxChunks + 8;
  const bool aligned = (((uintptr_t)d_crd) & 15) == 0;
  const char* env = getenv("B200_1VN_STREAM");
  const bool stream = aligned && !(env && atoi(env) == 0);
  if (stream) {
    static std::atomic<bool> attr[64][4];
    int dev = 0;
    cudaGetDevice(&dev);
    const int slot = (sizeof(T) == 8 ? 1 : 0) + (v2 ? 2 : 0);
    if (!attr[dev & 63][slot].load(std::memory_order_acquire)) {
      if (v2) CU(cudaFuncSetAttribute(onevn_stream2_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, ONEVN2_SMEM_BYTES));
      else CU(cudaFuncSetAttribute(onevn_stream_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, ONEVN_S_SMEM_BYTES));
      attr[dev & 63][slot].store(true, std::memory_order_release);
    }
    COUNT_LAUNCH();
    onevn_prep_kernel<<<1, 1024, 0, st>>>(setupRef, setupMass, nAtoms, const_cast<double*>(refw), const_cast<double*>(refsum),
                                          const_cast<double*>(refw) + (size_t)4 * nAtoms, d_atomIdx, APC, maxChunks, hdr, kLo,
                                          v2 ? partSum : nullptr);
    const int sms = numSMs > 0 ? numSMs : 148;
    if (v2) {
      OneVN2Args sa;
      sa.crd = d_crd; sa.stride = stride; sa.frameIdx = d_frameIdx; sa.srcBase = srcBase; sa.nFrames = nFrames;
      sa.atomIdx = d_atomIdx; sa.nAtoms = nAtoms; sa.refmw = refw + (size_t)4 * nAtoms; sa.hdr = hdr; sa.kLo = kLo; sa.rec = rec;
      COUNT_LAUNCH();
      { std::lock_guard<std::mutex> tl(g_streamMu); if (g_streamTimer.begin(st)) ++g_streamLaunches; }
      onevn_stream2_kernel<T><<<std::max(sms, maxChunks), ONEVN2_THREADS, ONEVN2_SMEM_BYTES, st>>>(sa);
      { std::lock_guard<std::mutex> tl(g_streamMu); g_streamTimer.end(st); }
      COUNT_LAUNCH();
      onevn_finish2_kernel<<<(nFrames + 127) / 128, 128, 0, st>>>(rec, hdr, kLo, partSum, nFrames, refsum, rmsd, rot, trans);
    } else {
      OneVNStreamArgs sa;
      sa.crd = d_crd; sa.stride = stride; sa.frameIdx = d_frameIdx; sa.srcBase = srcBase; sa.nFrames = nFrames;
      sa.atomIdx = d_atomIdx; sa.nAtoms = nAtoms;
      sa.refw = refw; sa.hdr = hdr; sa.kLo = kLo; sa.fit = fit; sa.rec = rec;
      const int nGroups = (nFrames + ONEVN_FB - 1) / ONEVN_FB;
      COUNT_LAUNCH();
      { std::lock_guard<std::mutex> tl(g_streamMu); if (g_streamTimer.begin(st)) ++g_streamLaunches; }
      onevn_stream_kernel<T><<<std::min(nGroups, sms), ONEVN_S_THREADS, ONEVN_S_SMEM_BYTES, st>>>(sa);
      { std::lock_guard<std::mutex> tl(g_streamMu); g_streamTimer.end(st); }
      COUNT_LAUNCH();
      onevn_finish_kernel<<<(nFrames + 127) / 128, 128, 0, st>>>(rec, hdr, nFrames, refsum, fit, rmsd, rot, trans);
    }
  }
  else if (setupRef) {
    COUNT_LAUNCH();
    onevn_setup_kernel<<<1, 256, 0, st>>>(setupRef, setupMass, nAtoms, const_cast<double*>(refw), const_cast<double*>(refsum),
                                          const_cast<double*>(refw) + (size_t)4 * nAtoms);
  }
  OneVNArgs a;
  a.crd = d_crd; a.stride = stride; a.frameIdx = d_frameIdx; a.srcBase = srcBase; a.nFrames = nFrames;
  a.atomIdx = d_atomIdx; a.nAtoms = nAtoms;
  a.refw = refw; a.refsum = refsum; a.skipIf = stream ? hdr : nullptr; a.fit = fit;
  a.rmsd = rmsd; a.rot = rot; a.trans = trans;
  COUNT_LAUNCH();
  {   // the general kernel returns at once when the streaming path took the call: a small grid keeps that cheap
    const int nGroups = (nFrames + ONEVN_FB - 1) / ONEVN_FB;
    const int cap = 8 * (numSMs > 0 ? numSMs : 148);
    onevn_kernel<T><<<stream ? std::min(nGroups, cap) : nGroups, ONEVN_THREADS, 0, st>>>(a);
  }
  CU(cudaGetLastError());
  return B200_OK;
}

template <typename T>
static int onevn_launch(b200_1vN* h, OneVNLane& L, const void* d_crd, size_t stride, int nFrames, const int* d_atomIdx, long outOffset) {
  int rc;
  if ((rc = L.ws.reserve(onevn_ws_bytes(stride, nFrames)))) return rc;
  L.timer.begin(L.st);
  rc = onevn_run<T>(L.dev->numSMs, d_crd, stride, nFrames, d_atomIdx, h->nAtoms, (const double*)L.refw.p,
                    (const double*)L.refsum.p, h->fit, (double*)L.rms.p + outOffset,
                    h->wantRot ? (double*)L.rot.p + 9 * outOffset : nullptr,
                    h->wantRot ? (double*)L.trans.p + 3 * outOffset : nullptr, L.ws.p, L.st);
  L.timer.end(L.st);
  if (rc) return rc;
  L.launches++;
  return B200_OK;
}

/// (Re)loads the reference of every lane: refw = (rx, ry, rz, m) per atom, refsum = its moments.
static int onevn_load_ref(b200_1vN* h, const double* refSelected) {
  int rc;
  for (auto& lp : h->lanes) {
    OneVNLane& L = *lp;
    CU(cudaSetDevice(L.dev->id));
    if ((rc = upload_vec(L.ref, refSelected, (size_t)3 * h->nAtoms, L.st))) return rc;
    COUNT_LAUNCH();
    onevn_setup_kernel<<<1, 256, 0, L.st>>>((const double*)L.ref.p, h->hasMass ? (const double*)L.mass.p : nullptr, h->nAtoms,
                                           (double*)L.refw.p, (double*)L.refsum.p, (double*)L.refw.p + (size_t)4 * h->nAtoms);
    CU(cudaGetLastError());
  }
  // the caller's reference array may change as soon as we return (reftraj / previous): wait for the uploads
  for (auto& lp : h->lanes) { CU(cudaSetDevice(lp->dev->id)); CU(cudaStreamSynchronize(lp->st)); }
  return B200_OK;
}

extern "C" {

int b200_rmsd_1vN_begin(const double* refSelected, const int* atomIdx, int nAtoms, const double* mass, int fit, int wantRot,
                        b200_1vN** handle) {
  if (!handle) return fail(B200_ERR_ARG, "null handle");
  *handle = nullptr;
  if (!refSelected || !atomIdx || nAtoms <= 0) return fail(B200_ERR_ARG, "bad reference / selection");
  for (int k = 0; k < nAtoms; ++k) if (atomIdx[k] < 0) return fail(B200_ERR_ARG, "negative atom index");
  std::lock_guard<std::mutex> lk(g_mu);
  int rc;
  if ((rc = ensure_init_locked())) return rc;
  b200_1vN* h = new (std::nothrow) b200_1vN();
  if (!h) return fail(B200_ERR_NOMEM, "out of memory");
  h->nAtoms = nAtoms; h->fit = fit ? 1 : 0; h->wantRot = (wantRot && fit) ? 1 : 0; h->hasMass = mass != nullptr;
  h->atomIdx.assign(atomIdx, atomIdx + nAtoms);
  h->identity = true;
  for (int k = 0; k < nAtoms; ++k) { h->maxAtom = std::max(h->maxAtom, atomIdx[k]); if (atomIdx[k] != k) h->identity = false; }
  auto setup = [&]() -> int {
    for (Device& d : g_devs) {
      h->lanes.emplace_back(new OneVNLane());
      OneVNLane& L = *h->lanes.back();
      L.dev = &d;
      CU(cudaSetDevice(d.id));
      CU(cudaStreamCreateWithFlags(&L.st, cudaStreamNonBlocking));
      for (int s = 0; s < ONEVN_NIN; ++s) CU(cudaEventCreateWithFlags(&L.slotFree[s], cudaEventDisableTiming));
      CU(cudaEventCreateWithFlags(&L.copied, cudaEventDisableTiming));
      int r;
      if (mass && (r = upload_vec(L.mass, mass, (size_t)nAtoms, L.st))) return r;
      if ((r = upload_vec(L.idx, atomIdx, (size_t)nAtoms, L.st))) return r;
      if ((r = L.refw.reserve((size_t)nAtoms * 64))) return r;   // (rx, ry, rz, m) and (m rx, m ry, m rz, m)
      if ((r = L.refsum.reserve(64))) return r;
    }
    return onevn_load_ref(h, refSelected);
  };
  rc = setup();
  cudaSetDevice(g_devs[0].id);
  if (rc) { onevn_destroy(h); return rc; }
  *handle = h;
  return B200_OK;
}

int b200_rmsd_1vN_set_ref(b200_1vN* h, const double* refSelected) {
  if (!h) return fail(B200_ERR_STATE, "null handle");
  if (!refSelected) return fail(B200_ERR_ARG, "null reference");
  std::lock_guard<std::mutex> lk(g_mu);
  const int rc = onevn_load_ref(h, refSelected);   // in stream order: frames pushed so far keep the old reference
  cudaSetDevice(g_devs[0].id);
  return rc;
}

}  // extern "C"

template <typename T>
static int onevn_push(b200_1vN* h, const T* src, size_t stride, int nFrames) {
  if (!h) return fail(B200_ERR_STATE, "null handle");
  if (nFrames <= 0) return B200_OK;
  if (!src) return fail(B200_ERR_ARG, "null frames");
  std::lock_guard<std::mutex> lk(g_mu);
  int rc;
  if ((size_t)3 * ((size_t)h->maxAtom + 1) > stride)
    return fail(B200_ERR_ARG, "atom index %d outside frame stride %zu", h->maxAtom, stride);
  const bool pinned = host_ptr_is_pinned(src);
  const int N = h->nAtoms;
  // chunk so that a slot holds <= 64 MB; several lanes: at least one chunk each when there is enough to share
  const size_t width = (size_t)3 * ((size_t)h->maxAtom + 1);
  const size_t perFrame = pinned ? width * sizeof(T) : (size_t)3 * N * sizeof(T);
  const int nl = (int)h->lanes.size();
  int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)nFrames, ((size_t)64 << 20) / perFrame));
  if (nl > 1 && nFrames >= 64 * nl) chunk = std::min(chunk, (nFrames + nl - 1) / nl);
  for (auto& lp : h->lanes) lp->touched = false;
  for (int f0 = 0; f0 < nFrames; f0 += chunk) {
    const int nf = std::min(chunk, nFrames - f0);
    const int li = h->nextLane;
    h->nextLane = (h->nextLane + 1) % nl;
    OneVNLane& L = *h->lanes[li];
    CU(cudaSetDevice(L.dev->id));
    if ((rc = onevn_grow(h, L, L.used + nf))) return rc;
    const int s = L.slot;
    L.slot = (L.slot + 1) % ONEVN_NIN;
    CU(cudaEventSynchronize(L.slotFree[s]));  // previous user of this slot is done (no-op when never recorded)
    if ((rc = L.in[s].reserve((size_t)nf * perFrame))) return rc;
    if (pinned) {
      // direct DMA of the needed span; gather by atomIdx on the device
      if (width == stride)   // whole frames: contiguous on both sides, one linear DMA
        CU(cudaMemcpyAsync(L.in[s].p, src + (size_t)f0 * stride, (size_t)nf * width * sizeof(T), cudaMemcpyHostToDevice, L.st));
      else
        CU(cudaMemcpy2DAsync(L.in[s].p, width * sizeof(T), src + (size_t)f0 * stride, stride * sizeof(T), width * sizeof(T),
                             (size_t)nf, cudaMemcpyHostToDevice, L.st));
      CU(cudaEventRecord(L.copied, L.st));
      L.touched = true;
      if ((rc = onevn_launch<T>(h, L, L.in[s].p, width, nf, (const int*)L.idx.p, L.used))) return rc;
    } else {
      // pageable source (e.g. cpptraj Frames that are reused): the copy pool packs the selected atoms into pinned staging
      if ((rc = L.stage[s].reserve((size_t)nf * perFrame))) return rc;
      T* stg = (T*)L.stage[s].p;
      const T* base = src + (size_t)f0 * stride;
      if (h->identity) {
        L.dev->pool.copy2d(stg, (size_t)3 * N * sizeof(T), base, stride * sizeof(T), (size_t)3 * N * sizeof(T), (size_t)nf);
      } else {
        const int* ai = h->atomIdx.data();
        const int fpp = std::max(1, (int)(((size_t)1 << 20) / ((size_t)3 * N * sizeof(T))));
        L.dev->pool.run((size_t)(nf + fpp - 1) / fpp, [=](size_t pc) {
          const int fEnd = std::min(nf, (int)(pc + 1) * fpp);
          for (int f = (int)pc * fpp; f < fEnd; ++f) {
            const T* fr = base + (size_t)f * stride;
            T* o = stg + (size_t)f * 3 * N;
            for (int k = 0; k < N; ++k) {
              const size_t a3 = (size_t)3 * ai[k];
              o[3 * k] = fr[a3]; o[3 * k + 1] = fr[a3 + 1]; o[3 * k + 2] = fr[a3 + 2];
            }
          }
        });
      }
      CU(cudaMemcpyAsync(L.in[s].p, stg, (size_t)nf * perFrame, cudaMemcpyHostToDevice, L.st));
      if ((rc = onevn_launch<T>(h, L, L.in[s].p, (size_t)3 * N, nf, nullptr, L.used))) return rc;
    }
    CU(cudaEventRecord(L.slotFree[s], L.st));
    h->segs.push_back({li, L.used, (long)nf});
    L.used += nf;
    h->h2d += (double)nf * perFrame;
    h->pushed += nf;
  }
  // the call is blocking as far as the caller's buffer is concerned: its last bytes have left for the device
  for (auto& lp : h->lanes)
    if (lp->touched) { CU(cudaSetDevice(lp->dev->id)); CU(cudaEventSynchronize(lp->copied)); }
  cudaSetDevice(g_devs[0].id);
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
  const long n = h->pushed - h->flushed;
  if (n > 0 && !rmsdOut) return fail(B200_ERR_ARG, "null rmsdOut");
  long at = 0;
  for (const OneVNSegment& sg : h->segs) {
    OneVNLane& L = *h->lanes[sg.lane];
    CU(cudaSetDevice(L.dev->id));
    CU(cudaMemcpyAsync(rmsdOut + at, (double*)L.rms.p + sg.laneOff, (size_t)sg.n * 8, cudaMemcpyDeviceToHost, L.st));
    h->d2h += (double)sg.n * 8;
    if (rotOut && h->wantRot) { CU(cudaMemcpyAsync(rotOut + 9 * at, (double*)L.rot.p + 9 * sg.laneOff, (size_t)sg.n * 72, cudaMemcpyDeviceToHost, L.st)); h->d2h += (double)sg.n * 72; }
    if (transOut && h->wantRot) { CU(cudaMemcpyAsync(transOut + 3 * at, (double*)L.trans.p + 3 * sg.laneOff, (size_t)sg.n * 24, cudaMemcpyDeviceToHost, L.st)); h->d2h += (double)sg.n * 24; }
    at += sg.n;
  }
  double ms = 0.0;
  long launches = 0;
  for (auto& lp : h->lanes) {
    CU(cudaSetDevice(lp->dev->id));
    CU(cudaStreamSynchronize(lp->st));
    ms += lp->timer.resolve(); launches += lp->launches;
    lp->launches = 0; lp->used = 0;
  }
  cudaSetDevice(g_devs[0].id);
  h->segs.clear();
  for (long i = 0; i < n; ++i) {
    if (h->bestFrame < 0 || rmsdOut[i] < h->bestVal) { h->bestVal = rmsdOut[i]; h->bestFrame = h->flushed + i; }
  }
  if (argminFrame) *argminFrame = h->bestFrame;
  {
    std::lock_guard<std::mutex> sl(g_statMu);
    g_stats.onevn_ms += ms; g_stats.onevn_launches += launches; g_stats.frames_1vN += (double)n;
    g_stats.h2d_bytes += h->h2d; g_stats.d2h_bytes += h->d2h;
    h->h2d = 0; h->d2h = 0;
  }
  h->flushed = h->pushed;
  return B200_OK;
}

int b200_rmsd_1vN_end(b200_1vN* h) {
  if (!h) return B200_OK;
  std::lock_guard<std::mutex> lk(g_mu);
  onevn_destroy(h);
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
  for (Device& x : g_devs) if (x.id == dev) d = &x;
  if (!d) return fail(B200_ERR_STATE, "current device %d was not initialised by b200_init", dev);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t refBytes = ((size_t)nAtoms * 64 + 64 + 63) & ~(size_t)63;
  if ((rc = d->onevnWs.reserve(refBytes + onevn_ws_bytes(frameStrideFloats, nFrames)))) return rc;
  double* refw = (double*)d->onevnWs.p;
  double* refsum = refw + (size_t)8 * nAtoms;
  // (profiling: the events are only recorded here and resolved by b200_get_stats -- the call stays asynchronous)
  if (g_profiling) { std::lock_guard<std::mutex> tl(g_streamMu); if (g_devPassTimer.begin(st)) { ++g_devPasses; g_devPassFrames += (double)nFrames; } }
  rc = onevn_run<float>(d->numSMs, d_crd, frameStrideFloats, nFrames, d_atomIdx, nAtoms, refw, refsum, fit ? 1 : 0, d_rmsdOut,
                        fit ? d_rotOut : nullptr, fit ? d_transOut : nullptr, (char*)d->onevnWs.p + refBytes, st, nullptr, 0,
                        d_refSelected, d_mass);
  if (g_profiling) { std::lock_guard<std::mutex> tl(g_streamMu); g_devPassTimer.end(st); }
  return rc;
}


}  // extern "C"

namespace {
/// List positions [p0,p1) of the frames against all K centroids on one device.
/// K >= 3 and a fitted RMSD the tcgen05 engine qualifies for: ONE frames x centroids contraction (the frames are read
/// once, quantised once; the centroids are the column operand) -- the full-matrix mode of the pair engine.  Otherwise
/// K streaming one-vs-many passes over the resident frames (K * 12 N bytes per frame; cheaper than quantising for
/// K <= 2, and the only nofit / FP64 path).
int host_centroids_on_device(Device& d, const float* crd, size_t stride, int nFramesTotal, const int* frameIdx, int p0, int p1,
                             const int* atomIdx, int nAtoms, const double* mass, int fit, const double* centroids, int K,
                             double* distOut, int* closestOut, double* closestDistOut, ExtentShare* share) {
  ShareGuard guard(share);
  CU(cudaSetDevice(d.id));
  const int nP = p1 - p0;
  if (nP <= 0) {
    if (share) { share->leave(); guard.settled = true; }
    return B200_OK;
  }
  int maxAtom = 0, rc;
  if ((rc = validate_sel(atomIdx, nAtoms, stride, &maxAtom))) return rc;
  int sLo = p0, sHi = p1;
  if (frameIdx) {
    sLo = nFramesTotal; sHi = 0;
    for (int f = p0; f < p1; ++f) {
      if (frameIdx[f] < 0 || frameIdx[f] >= nFramesTotal) return fail(B200_ERR_ARG, "frameIdx[%d]=%d out of range", f, frameIdx[f]);
      sLo = std::min(sLo, frameIdx[f]); sHi = std::max(sHi, frameIdx[f] + 1);
    }
  } else if (p1 > nFramesTotal) {
    return fail(B200_ERR_ARG, "frame %d > nFramesTotal %d", p1, nFramesTotal);
  }
  cudaStream_t st = d.stream[0];
  double h2d = 0.0, d2h = 0.0;
  size_t width = (size_t)3 * ((size_t)maxAtom + 1);
  const float* d_frames = nullptr;
  long rowBase = 0;
  if ((rc = get_coords(d, d.crd, crd, stride, nFramesTotal, sLo, sHi, width, st, &h2d, &d_frames, &width, &rowBase))) return rc;
  // list position 0 when there is no frame list: frame p0; with a list: row = frameIdx[.] - rowBase
  const float* d_list0 = frameIdx ? d_frames : d_frames + (size_t)((long)p0 - rowBase) * width;
  if ((rc = upload_vec(d.idxA, atomIdx, (size_t)nAtoms, st))) return rc;
  if (mass && (rc = upload_vec(d.massA, mass, (size_t)nAtoms, st))) return rc;
  if (frameIdx && (rc = upload_vec(d.frameIdx, frameIdx + p0, (size_t)nP, st))) return rc;
  if ((rc = upload_vec(d.crdB, centroids, (size_t)K * 3 * (size_t)nAtoms, st))) return rc;
  h2d += (double)K * 24.0 * nAtoms;
  const double* d_mass = mass ? (const double*)d.massA.p : nullptr;
  const int* d_fidx = frameIdx ? (const int*)d.frameIdx.p : nullptr;
  // outputs (device): frame-major table, nearest centroid and its distance
  const size_t tabBytes = ((size_t)K * (size_t)nP * sizeof(double) + 63) & ~(size_t)63;
  if ((rc = d.planesB.reserve(tabBytes + (size_t)nP * (sizeof(int) + sizeof(double)) + 128))) return rc;
  double* dOutT = (double*)d.planesB.p;
  double* dClosestDist = (double*)((char*)d.planesB.p + tabBytes);
  int* dClosest = (int*)(dClosestDist + nP);
  Timer t;
  bool contracted = false;
  const int engine = pair_engine();
  if (fit && K >= 3 && engine != 1) {
    // ---- one contraction on the tcgen05 engine
    if ((rc = d.scal.reserve(64))) return rc;
    double* d_total = (double*)d.scal.p;
    unsigned int* d_maxBits = (unsigned int*)(d_total + 4);
    COUNT_LAUNCH();
    mass_sum_kernel<<<1, 32, 0, st>>>(d_mass, nAtoms, d_total);
    I8Set qA, qB;
    if ((rc = i8_reserve(qA, d.imgA, d.GA, d.cenA, nP, nAtoms))) return rc;
    if ((rc = i8_reserve(qB, d.imgB, d.GB, d.cenB, K, nAtoms))) return rc;
    CU(cudaMemsetAsync(d_maxBits, 0, sizeof(unsigned int), st));
    const long baseA = frameIdx ? rowBase : 0;
    if ((rc = i8_stats(qA, d_list0, width, d_fidx, baseA, 0, (const int*)d.idxA.p, nAtoms, d_mass, d_mass, d_maxBits, st))) return rc;
    // centroids: double rows of 3 N, already gathered; cpptraj keeps them centred (Metric_RMS.cpp:66-81), so the
    // centring here moves them by rounding noise only
    if ((rc = i8_stats(qB, d.crdB.p, (size_t)3 * nAtoms, nullptr, 0, 0, nullptr, nAtoms, d_mass, d_mass, d_maxBits, st, -1, true))) return rc;
    int qs = 0;
    bool ok = false;
    if ((rc = i8_choose_scale(d, d_maxBits, d_total, nAtoms, st, &qs, &ok, share))) return rc;
    guard.settled = true;
    if (ok) {
      if ((rc = i8_clear(qA, st)) || (rc = i8_clear(qB, st))) return rc;
      if ((rc = i8_quant(qA, d_list0, width, d_fidx, baseA, 0, (const int*)d.idxA.p, nAtoms, d_mass, qs, st))) return rc;
      if ((rc = i8_quant(qB, d.crdB.p, (size_t)3 * nAtoms, nullptr, 0, 0, nullptr, nAtoms, d_mass, qs, st, -1, true))) return rc;
      if ((rc = d.outChunk[0].reserve((size_t)nP * K * sizeof(float)))) return rc;
      t.begin(st);
      if ((rc = run_pair_i8_band(d, qA, qB, 0, nP, false, qs, d_total, (float*)d.outChunk[0].p, 0, (size_t)K, nullptr, st))) return rc;
      COUNT_LAUNCH();
      centroid_argmin_rows_kernel<<<(nP + 255) / 256, 256, 0, st>>>((const float*)d.outChunk[0].p, nP, K, distOut ? dOutT : nullptr,
                                                                    dClosest, dClosestDist);
      t.end(st);
      contracted = true;
      g_lastEngine.store(2); g_lastQs.store(qs);
    } else if (engine == 2) {
      return fail(B200_ERR_ARG, "tcgen05 int8 engine forced but the selection does not qualify (%d fractional bits, %d atoms)", qs, nAtoms);
    }
  } else if (share) {
    share->leave(); guard.settled = true;
  }
  if (!contracted) {
    // ---- K streaming passes
    const size_t refBytes = ((size_t)nAtoms * 64 + 64 + 63) & ~(size_t)63;
    const size_t wsBytes = (onevn_ws_bytes(width, nP) + 63) & ~(size_t)63;
    const size_t distBytes = (size_t)K * (size_t)nP * sizeof(double);
    if ((rc = d.onevnWs.reserve(refBytes + wsBytes + distBytes))) return rc;
    char* base = (char*)d.onevnWs.p;
    double* refw = (double*)base;
    double* refsum = refw + (size_t)8 * nAtoms;
    void* ws = base + refBytes;
    double* dist = (double*)(base + refBytes + wsBytes);
    t.begin(st);
    for (int k = 0; k < K; ++k) {
      if ((rc = onevn_run<float>(d.numSMs, d_list0, width, nP, (const int*)d.idxA.p, nAtoms, refw, refsum, fit ? 1 : 0,
                                 dist + (size_t)k * nP, nullptr, nullptr, ws, st, d_fidx, rowBase,
                                 (const double*)d.crdB.p + (size_t)k * 3 * nAtoms, d_mass))) return rc;
    }
    COUNT_LAUNCH();
    centroid_argmin_kernel<<<(nP + 255) / 256, 256, 0, st>>>(dist, nP, K, distOut ? dOutT : nullptr, dClosest, dClosestDist);
    t.end(st);
    g_lastEngine.store(1); g_lastQs.store(0);
  }
  if (distOut) { CU(cudaMemcpyAsync(distOut + (size_t)p0 * K, dOutT, (size_t)K * nP * sizeof(double), cudaMemcpyDeviceToHost, st)); d2h += 8.0 * K * nP; }
  if (closestOut) { CU(cudaMemcpyAsync(closestOut + p0, dClosest, (size_t)nP * sizeof(int), cudaMemcpyDeviceToHost, st)); d2h += 4.0 * nP; }
  if (closestDistOut) { CU(cudaMemcpyAsync(closestDistOut + p0, dClosestDist, (size_t)nP * sizeof(double), cudaMemcpyDeviceToHost, st)); d2h += 8.0 * nP; }
  CU(cudaStreamSynchronize(st));
  {
    const double ms = t.resolve();
    std::lock_guard<std::mutex> sl(g_statMu);
    if (contracted) { g_stats.pair_ms += ms; g_stats.pair_launches += 1; g_stats.pairs += (double)nP * K; }
    else { g_stats.onevn_ms += ms; g_stats.onevn_launches += K; g_stats.frames_1vN += (double)nP * K; }
    g_stats.h2d_bytes += h2d; g_stats.d2h_bytes += d2h;
  }
  return B200_OK;
}
}  // namespace

extern "C" {

int b200_rmsd_frames_to_centroids(const float* crd, size_t frameStrideFloats, int nFramesTotal, const int* frameIdx, int nFrames,
                                  const int* atomIdx, int nAtoms, const double* mass, int fit, const double* centroids,
                                  int nCentroids, double* distOut, int* closestOut, double* closestDistOut) {
  if (!crd || !centroids || nCentroids <= 0) return fail(B200_ERR_ARG, "bad argument");
  if (nFrames <= 0) return B200_OK;
  std::lock_guard<std::mutex> lk(g_mu);
  int rc;
  if ((rc = ensure_init_locked())) return rc;
  // frames in equal contiguous shares over the devices (small jobs stay on one)
  const bool resident = g_devs[0].resHost == crd && g_devs[0].resBuf.p != nullptr;   // (the resident copy lives on device 0)
  const int nd = (nFrames >= 4096 && !resident) ? (int)g_devs.size() : 1;
  const int per = (nFrames + nd - 1) / nd;
  ExtentShare share;
  share.parties = nd;
  auto job = [&](int i) -> int {
    const int p0 = std::min(nFrames, i * per), p1 = std::min(nFrames, (i + 1) * per);
    return host_centroids_on_device(g_devs[i], crd, frameStrideFloats, nFramesTotal, frameIdx, p0, p1, atomIdx, nAtoms, mass, fit,
                                    centroids, nCentroids, distOut, closestOut, closestDistOut, nd > 1 ? &share : nullptr);
  };
  if (nd == 1) return job(0);
  return for_each_device(job);
}

int b200_coords_resident_begin(const float* crd, size_t frameStrideFloats, int nFramesTotal, const int* atomIdx, int nAtoms) {
  if (!crd || nFramesTotal <= 0) return fail(B200_ERR_ARG, "bad argument");
  std::lock_guard<std::mutex> lk(g_mu);
  int rc, maxAtom = 0;
  if ((rc = ensure_init_locked())) return rc;
  if ((rc = validate_sel(atomIdx, nAtoms, frameStrideFloats, &maxAtom))) return rc;
  Device& d = g_devs[0];
  CU(cudaSetDevice(d.id));
  const size_t width = (size_t)3 * ((size_t)maxAtom + 1);
  d.resHost = nullptr;
  double h2d = 0.0;
  if ((rc = upload_crd(d, d.resBuf, crd, frameStrideFloats, 0, nFramesTotal, width, d.stream[0], &h2d))) return rc;
  CU(cudaStreamSynchronize(d.stream[0]));
  d.resHost = crd; d.resStride = frameStrideFloats; d.resWidth = width; d.resFrames = nFramesTotal;
  { std::lock_guard<std::mutex> sl(g_statMu); g_stats.h2d_bytes += h2d; }
  return B200_OK;
}

int b200_coords_resident_end(const float* crd) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_devs.empty()) return B200_OK;
  Device& d = g_devs[0];
  if (crd == nullptr || d.resHost == crd) {
    cudaSetDevice(d.id);
    d.resBuf.release();
    d.resHost = nullptr; d.resWidth = 0; d.resFrames = 0;
  }
  return B200_OK;
}

int b200_rmsd_build_centroids(const float* crd, size_t frameStrideFloats, int nFramesTotal, const int* frames, const int* offsets,
                              int nClusters, const int* atomIdx, int nAtoms, const double* mass, int fit, double* centroidsOut) {
  if (!crd || !frames || !offsets || !centroidsOut || nClusters <= 0) return fail(B200_ERR_ARG, "bad argument");
  std::lock_guard<std::mutex> lk(g_mu);
  int rc, maxAtom = 0;
  if ((rc = ensure_init_locked())) return rc;
  if ((rc = validate_sel(atomIdx, nAtoms, frameStrideFloats, &maxAtom))) return rc;
  const int nList = offsets[nClusters];
  if (offsets[0] != 0 || nList < 0) return fail(B200_ERR_ARG, "bad offsets");
  if (nList == 0) return B200_OK;
  int sLo = nFramesTotal, sHi = 0;
  for (int k = 0; k < nClusters; ++k) if (offsets[k + 1] < offsets[k]) return fail(B200_ERR_ARG, "offsets not ascending");
  for (int p = 0; p < nList; ++p) {
    if (frames[p] < 0 || frames[p] >= nFramesTotal) return fail(B200_ERR_ARG, "frames[%d]=%d out of range", p, frames[p]);
    sLo = std::min(sLo, frames[p]); sHi = std::max(sHi, frames[p] + 1);
  }
  Device& d = g_devs[0];
  CU(cudaSetDevice(d.id));
  cudaStream_t st = d.stream[0];
  double h2d = 0.0;
  const size_t width = (size_t)3 * ((size_t)maxAtom + 1);
  const float* d_frames = nullptr;
  size_t pitch = width;
  long rowBase = 0;
  if ((rc = get_coords(d, d.crd, crd, frameStrideFloats, nFramesTotal, sLo, sHi, width, st, &h2d, &d_frames, &pitch, &rowBase))) return rc;
  if ((rc = upload_vec(d.idxA, atomIdx, (size_t)nAtoms, st))) return rc;
  if (mass && (rc = upload_vec(d.massA, mass, (size_t)nAtoms, st))) return rc;
  if ((rc = upload_vec(d.frameIdx, frames, (size_t)nList, st))) return rc;
  if ((rc = upload_vec(d.idxB, offsets, (size_t)nClusters + 1, st))) return rc;
  const size_t outBytes = (size_t)nClusters * 3 * (size_t)nAtoms * sizeof(double);
  if ((rc = d.planesB.reserve(outBytes))) return rc;
  CU(cudaMemsetAsync(d.planesB.p, 0, outBytes, st));
  CentroidArgs a;
  a.crd = d_frames; a.stride = pitch; a.srcBase = rowBase; a.frames = (const int*)d.frameIdx.p;
  a.offsets = (const int*)d.idxB.p; a.atomIdx = (const int*)d.idxA.p; a.nAtoms = nAtoms;
  a.mass = mass ? (const double*)d.massA.p : nullptr; a.fit = fit ? 1 : 0; a.out = (double*)d.planesB.p;
  COUNT_LAUNCH();
  const size_t centSmem = (size_t)3 * (size_t)nAtoms * sizeof(double);
  if (centSmem <= (size_t)200 * 1024) {   // the running sum in shared memory
    static std::atomic<size_t> granted[64];
    if (granted[d.id & 63].load(std::memory_order_acquire) < centSmem) {
      CU(cudaFuncSetAttribute(centroid_build_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      granted[d.id & 63].store((size_t)200 * 1024, std::memory_order_release);
    }
    centroid_build_kernel<true><<<nClusters, CENT_THREADS, centSmem, st>>>(a);
  } else {
    centroid_build_kernel<false><<<nClusters, CENT_THREADS, 0, st>>>(a);
  }
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(centroidsOut, d.planesB.p, outBytes, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  {
    std::lock_guard<std::mutex> sl(g_statMu);
    g_stats.h2d_bytes += h2d; g_stats.d2h_bytes += (double)outBytes;
  }
  return B200_OK;
}

// ---- rmsavgcorr: RMSD of running-averaged coordinates, all window sizes (avgcorr.cuh)
/// The window sizes windows[0..nWindows) on one device (prefix sums built there).
static int avgcorr_on_device(Device& d, const float* crd, size_t frameStrideFloats, int nFrames, const int* atomIdx, int nAtoms,
                             int maxAtom, const double* mass, const double* refSelected, const int* windows, int nWindows,
                             double* avgOut, double* sdOut) {
  int rc;
  if (nWindows == 0) return B200_OK;
  CU(cudaSetDevice(d.id));
  cudaStream_t st = d.stream[0];
  double h2d = 0.0;
  const size_t width = (size_t)3 * ((size_t)maxAtom + 1);
  const float* d_frames = nullptr;
  size_t pitch = width;
  long rowBase = 0;
  if ((rc = get_coords(d, d.crd, crd, frameStrideFloats, nFrames, 0, nFrames, width, st, &h2d, &d_frames, &pitch, &rowBase))) return rc;
  if ((rc = upload_vec(d.idxA, atomIdx, (size_t)nAtoms, st))) return rc;
  if (mass && (rc = upload_vec(d.massA, mass, (size_t)nAtoms, st))) return rc;
  std::vector<double> refPlanes;   // the fixed reference plane-major (x | y | z), like a row of the prefix sums
  if (refSelected) {
    refPlanes.resize((size_t)3 * nAtoms);
    for (int k = 0; k < nAtoms; ++k)
      for (int c = 0; c < 3; ++c) refPlanes[(size_t)c * nAtoms + k] = refSelected[3 * (size_t)k + c];
    if ((rc = upload_vec(d.planesB, refPlanes.data(), (size_t)3 * nAtoms, st))) return rc;
  }
  double totalMass = 0.0;   // in atom order, as Frame::RMSD_CenteredRef sums it (src/Frame.cpp:1150-1158)
  for (int k = 0; k < nAtoms; ++k) totalMass += mass ? mass[k] : 1.0;
  const int ld = 3 * nAtoms;
  if ((rc = d.acP.reserve(((size_t)nFrames + 1) * ld * sizeof(double)))) return rc;
  if ((rc = d.scal.reserve(64))) return rc;
  COUNT_LAUNCH();
  avgcorr_shift_kernel<<<1, 32, 0, st>>>(d_frames, pitch, rowBase, (const int*)d.idxA.p, nAtoms, (double*)d.scal.p);
  COUNT_LAUNCH();
  avgcorr_prefix_kernel<<<(ld + 127) / 128, 128, 0, st>>>(d_frames, pitch, rowBase, (const int*)d.idxA.p, nAtoms, nFrames,
                                                          (const double*)d.scal.p, (double*)d.acP.p);
  CU(cudaGetLastError());
  // windows in batches of bounded RMSD count (one RMSD per window and averaged frame)
  const long long maxItems = (long long)32 << 20;
  std::vector<long long> off;
  std::vector<double> hAvg, hSd;
  for (int w0 = 0; w0 < nWindows;) {
    off.assign(1, 0);
    int w1 = w0, maxPer = 0;
    while (w1 < nWindows && w1 - w0 < 65535) {
      const long long items = (long long)nFrames - windows[w1] + 1;
      if (w1 > w0 && off.back() + items > maxItems) break;
      off.push_back(off.back() + items);
      maxPer = std::max<long long>(maxPer, items);
      ++w1;
    }
    const int nW = w1 - w0;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t oWin = 0, oOff = up((size_t)nW * sizeof(int)), oInfo = oOff + up(((size_t)nW + 1) * sizeof(long long)),
                 oAvg = oInfo + up((size_t)nW * 8 * sizeof(double)), oSd = oAvg + up((size_t)nW * sizeof(double)),
                 total = oSd + up((size_t)nW * sizeof(double));
    if ((rc = d.acMisc.reserve(total))) return rc;
    if ((rc = d.acRms.reserve((size_t)off.back() * sizeof(double)))) return rc;
    char* base = (char*)d.acMisc.p;
    CU(cudaMemcpyAsync(base + oWin, windows + w0, (size_t)nW * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(base + oOff, off.data(), ((size_t)nW + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    AvgCorrArgs a;
    a.P = (const double*)d.acP.p; a.ld = ld; a.nAtoms = nAtoms; a.nFrames = nFrames;
    a.mass = mass ? (const double*)d.massA.p : nullptr;
    a.refFixed = refSelected ? (const double*)d.planesB.p : nullptr;
    a.win = (const int*)(base + oWin); a.itemOff = (const long long*)(base + oOff); a.nW = nW;
    a.refInfo = (double*)(base + oInfo); a.rms = (double*)d.acRms.p; a.totalMass = totalMass;
    COUNT_LAUNCH();
    avgcorr_ref_kernel<<<nW, 256, 0, st>>>(a);
    CU(cudaGetLastError());
    COUNT_LAUNCH();
    {
      static const int nwin = getenv("B200_AVGCORR_NWIN") ? atoi(getenv("B200_AVGCORR_NWIN")) : 2;   // window sizes per warp (measured, 10k x 1k: 0.42 / 0.31 / 0.35 s for 1 / 2 / 4)
      const dim3 blk(AVGCORR_WARPS * 32);
      const unsigned gx = (unsigned)((maxPer + AVGCORR_WARPS - 1) / AVGCORR_WARPS);
      if (nwin <= 1) avgcorr_kernel<1><<<dim3(gx, nW), blk, 0, st>>>(a);
      else if (nwin == 2) avgcorr_kernel<2><<<dim3(gx, (nW + 1) / 2), blk, 0, st>>>(a);
      else avgcorr_kernel<4><<<dim3(gx, (nW + 3) / 4), blk, 0, st>>>(a);
    }
    CU(cudaGetLastError());
    COUNT_LAUNCH();
    avgcorr_reduce_kernel<<<nW, 256, 0, st>>>(a.rms, a.itemOff, nW, (double*)(base + oAvg), (double*)(base + oSd));
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(avgOut + w0, base + oAvg, (size_t)nW * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(sdOut + w0, base + oSd, (size_t)nW * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));   // (the staging vectors and the per-batch buffers are reused)
    w0 = w1;
  }
  {
    std::lock_guard<std::mutex> sl(g_statMu);
    g_stats.h2d_bytes += h2d; g_stats.d2h_bytes += (double)nWindows * 16.0;
  }
  return B200_OK;
}

int b200_rmsavgcorr(const float* crd, size_t frameStrideFloats, int nFrames, const int* atomIdx, int nAtoms, const double* mass,
                    const double* refSelected, const int* windows, int nWindows, double* avgOut, double* sdOut) {
  if (!crd || !windows || !avgOut || !sdOut) return fail(B200_ERR_ARG, "null argument");
  if (nFrames < 1 || nWindows < 0) return fail(B200_ERR_ARG, "bad frame or window count");
  for (int w = 0; w < nWindows; ++w)
    if (windows[w] < 1 || windows[w] > nFrames) return fail(B200_ERR_ARG, "windows[%d]=%d outside 1..%d", w, windows[w], nFrames);
  std::lock_guard<std::mutex> lk(g_mu);
  int rc, maxAtom = 0;
  if ((rc = ensure_init_locked())) return rc;
  if ((rc = validate_sel(atomIdx, nAtoms, frameStrideFloats, &maxAtom))) return rc;
  if (nWindows == 0) return B200_OK;
  // Window sizes are independent: with several devices, device k takes windows k, k + nd, ... (the cost of a window size
  // falls with its size: interleaving balances) and builds its own prefix sums.
  const int nd = (nWindows >= 64 * (int)g_devs.size()) ? (int)g_devs.size() : 1;
  if (nd == 1) return avgcorr_on_device(g_devs[0], crd, frameStrideFloats, nFrames, atomIdx, nAtoms, maxAtom, mass, refSelected,
                                        windows, nWindows, avgOut, sdOut);
  std::vector<std::vector<int>> win(nd);
  std::vector<std::vector<double>> av(nd), sd(nd);
  for (int w = 0; w < nWindows; ++w) win[w % nd].push_back(windows[w]);
  for (int k = 0; k < nd; ++k) { av[k].resize(win[k].size()); sd[k].resize(win[k].size()); }
  rc = for_each_device([&](int k) {
    return avgcorr_on_device(g_devs[k], crd, frameStrideFloats, nFrames, atomIdx, nAtoms, maxAtom, mass, refSelected,
                             win[k].data(), (int)win[k].size(), av[k].data(), sd[k].data());
  });
  if (rc) return rc;
  for (int w = 0; w < nWindows; ++w) { avgOut[w] = av[w % nd][w / nd]; sdOut[w] = sd[w % nd][w / nd]; }
  return B200_OK;
}

/// Device copy of a cache triangle of n frames: the resident one when the caller announced it (b200_cache_resident_begin),
/// else an upload into the device's scratch (rows of 1 Mi floats + a remainder: pinned -> DMA, pageable -> staged by the
/// copy pool).
static int get_cache(Device& d, const float* tri, int n, cudaStream_t st, double* h2d, const float** d_tri, DevBuf* into = nullptr) {
  if (!into && d.cacheHost == tri && d.cacheN == n && d.cacheBuf.p) { *d_tri = (const float*)d.cacheBuf.p; return B200_OK; }
  DevBuf& buf = into ? *into : d.haTri;
  const size_t nElt = (size_t)n * (size_t)(n - 1) / 2;
  int rc;
  if ((rc = buf.reserve(std::max<size_t>(nElt, 1) * sizeof(float)))) return rc;
  const size_t W = (size_t)1 << 20, rows = nElt / W, rem = nElt - rows * W;
  const bool pinned = host_ptr_is_pinned(tri);
  if (rows && (rc = upload_rows(d, (float*)buf.p, tri, W, 0, (int)rows, W, pinned, st, h2d))) return rc;
  if (rem && (rc = upload_rows(d, (float*)buf.p + rows * W, tri + rows * W, rem, 0, 1, rem, pinned, st, h2d))) return rc;
  *d_tri = (const float*)buf.p;
  return B200_OK;
}

int b200_cache_resident_begin(const float* tri, int nCached) {
  if (!tri || nCached < 2) return fail(B200_ERR_ARG, "bad cache");
  std::lock_guard<std::mutex> lk(g_mu);
  int rc;
  if ((rc = ensure_init_locked())) return rc;
  Device& d = g_devs[0];
  CU(cudaSetDevice(d.id));
  d.cacheHost = nullptr; d.cacheN = 0;
  double h2d = 0.0;
  const float* d_tri = nullptr;
  if ((rc = get_cache(d, tri, nCached, d.stream[0], &h2d, &d_tri, &d.cacheBuf))) return rc;
  CU(cudaStreamSynchronize(d.stream[0]));
  d.cacheHost = tri; d.cacheN = nCached;
  std::lock_guard<std::mutex> sl(g_statMu);
  g_stats.h2d_bytes += h2d;
  return B200_OK;
}

int b200_cache_resident_end(const float* tri) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_devs.empty()) return B200_OK;
  Device& d = g_devs[0];
  if (tri && d.cacheHost != tri) return B200_OK;
  cudaSetDevice(d.id);
  d.cacheHost = nullptr; d.cacheN = 0;
  d.cacheBuf.release();
  return B200_OK;
}

int b200_cache_cluster_sums(const float* tri, int nCached, const int* members, const int* offsets, int nClusters,
                            double* cumOut, double* upOut, double* up2Out) {
  if (!tri || !members || !offsets || !cumOut || nClusters <= 0 || nCached < 2) return fail(B200_ERR_ARG, "bad argument");
  std::lock_guard<std::mutex> lk(g_mu);
  int rc;
  if ((rc = ensure_init_locked())) return rc;
  const int total = offsets[nClusters];
  if (offsets[0] != 0 || total < 0) return fail(B200_ERR_ARG, "bad offsets");
  if (total == 0) return B200_OK;
  std::vector<int> clusterOf((size_t)total);
  for (int c = 0; c < nClusters; ++c) {
    if (offsets[c + 1] < offsets[c]) return fail(B200_ERR_ARG, "offsets not ascending");
    for (int p = offsets[c]; p < offsets[c + 1]; ++p) {
      if (members[p] < 0 || members[p] >= nCached) return fail(B200_ERR_ARG, "members[%d]=%d outside the cache", p, members[p]);
      clusterOf[p] = c;
    }
  }
  Device& d = g_devs[0];
  CU(cudaSetDevice(d.id));
  cudaStream_t st = d.stream[0];
  double h2d = 0.0;
  const float* d_tri = nullptr;
  if ((rc = get_cache(d, tri, nCached, st, &h2d, &d_tri))) return rc;
  if ((rc = upload_vec(d.frameIdx, members, (size_t)total, st))) return rc;
  if ((rc = upload_vec(d.idxB, offsets, (size_t)nClusters + 1, st))) return rc;
  if ((rc = upload_vec(d.idxA, clusterOf.data(), (size_t)total, st))) return rc;
  if ((rc = d.planesB.reserve((size_t)3 * total * sizeof(double)))) return rc;
  double* dc = (double*)d.planesB.p;
  COUNT_LAUNCH();
  cache_cluster_sums_kernel<<<(total + 127) / 128, 128, 0, st>>>(d_tri, nCached, (const int*)d.frameIdx.p, (const int*)d.idxB.p,
                                                                (const int*)d.idxA.p, total, dc, upOut ? dc + total : nullptr,
                                                                up2Out ? dc + 2 * (size_t)total : nullptr);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(cumOut, dc, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (upOut) CU(cudaMemcpyAsync(upOut, dc + total, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (up2Out) CU(cudaMemcpyAsync(up2Out, dc + 2 * (size_t)total, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  std::lock_guard<std::mutex> sl(g_statMu);
  g_stats.h2d_bytes += h2d; g_stats.d2h_bytes += (double)total * 8.0 * (1 + (upOut != nullptr) + (up2Out != nullptr));
  return B200_OK;
}

int b200_cache_cluster_links(const float* tri, int nCached, const int* label, int nClusters, double* minOut, double* maxOut,
                             double* sumOut, long long* countOut) {
  if (!tri || !label || !minOut || !maxOut || !sumOut || !countOut || nClusters <= 0 || nCached < 2) return fail(B200_ERR_ARG, "bad argument");
  if (nClusters > 4096) return fail(B200_ERR_ARG, "more than 4096 clusters");
  std::lock_guard<std::mutex> lk(g_mu);
  int rc;
  if ((rc = ensure_init_locked())) return rc;
  for (int f = 0; f < nCached; ++f)
    if (label[f] >= nClusters) return fail(B200_ERR_ARG, "label[%d]=%d >= %d clusters", f, label[f], nClusters);
  Device& d = g_devs[0];
  CU(cudaSetDevice(d.id));
  cudaStream_t st = d.stream[0];
  double h2d = 0.0;
  const float* d_tri = nullptr;
  if ((rc = get_cache(d, tri, nCached, st, &h2d, &d_tri))) return rc;
  if ((rc = upload_vec(d.frameIdx, label, (size_t)nCached, st))) return rc;
  const size_t K2 = (size_t)nClusters * nClusters;
  std::vector<LinkCell> cells(K2);
  for (LinkCell& c : cells) { c.mn = 0xffffffffu; c.mx = 0u; c.cnt = 0ull; c.sum = 0.0; }
  if ((rc = upload_vec(d.planesB, cells.data(), K2, st))) return rc;
  const size_t smem = K2 * sizeof(LinkCell) <= 40960 ? K2 * sizeof(LinkCell) : 0;
  COUNT_LAUNCH();
  cache_cluster_links_kernel<<<std::min(nCached - 1, d.numSMs * 8), 256, smem, st>>>(d_tri, nCached, (const int*)d.frameIdx.p, nClusters,
                                                                                   (LinkCell*)d.planesB.p);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(cells.data(), d.planesB.p, K2 * sizeof(LinkCell), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  auto unord = [](unsigned int u) { unsigned int b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u; float f; std::memcpy(&f, &b, 4); return (double)f; };
  for (size_t x = 0; x < K2; ++x) {
    countOut[x] = (long long)cells[x].cnt;
    sumOut[x] = cells[x].sum;
    minOut[x] = cells[x].cnt ? unord(cells[x].mn) : 0.0;
    maxOut[x] = cells[x].cnt ? unord(cells[x].mx) : 0.0;
  }
  std::lock_guard<std::mutex> sl(g_statMu);
  g_stats.h2d_bytes += h2d; g_stats.d2h_bytes += (double)(K2 * sizeof(LinkCell));
  return B200_OK;
}

// ---- cluster: hierarchical agglomerative clustering on the cache triangle (hieragglo.cuh)
int b200_hieragglo(const float* tri, int nFrames, int linkage, int targetClusters, double epsilon,
                   int* mergeInto, int* mergeFrom, float* findMin, int* nCalls, int* nMerges) {
  if (!tri || !mergeInto || !mergeFrom || !findMin || !nCalls || !nMerges) return fail(B200_ERR_ARG, "null argument");
  if (linkage < 0 || linkage > 2) return fail(B200_ERR_ARG, "linkage must be 0 (single), 1 (average) or 2 (complete)");
  if (nFrames < 0) return fail(B200_ERR_ARG, "nFrames < 0");
  *nCalls = 0; *nMerges = 0;
  std::lock_guard<std::mutex> lk(g_mu);
  int rc;
  if ((rc = ensure_init_locked())) return rc;
  if (nFrames < 2) return B200_OK;   // nothing to merge
  Device& d = g_devs[0];
  CU(cudaSetDevice(d.id));
  cudaStream_t st = d.stream[0];
  const size_t n = (size_t)nFrames, nElt = n * (n - 1) / 2;
  double h2d = 0.0;
  const float* d_tri = nullptr;
  if ((rc = get_cache(d, tri, nFrames, st, &h2d, &d_tri))) return rc;
  // symmetric n x n working matrices (rows contiguous): 4 bytes per entry, + 8 for the sums of average linkage
  if ((rc = d.haD.reserve(n * n * sizeof(float)))) return rc;
  if (linkage == 1 && (rc = d.haS.reserve(n * n * sizeof(double)))) return rc;
  // per-cluster state, 256-byte aligned pieces of one allocation
  auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
  size_t off = 0;
  const size_t oKeys = off;    off += up(n * sizeof(ha_u64));
  const size_t oClosest = off; off += up(n * sizeof(int));
  const size_t oCmin = off;    off += up(n * sizeof(float));
  const size_t oNfr = off;     off += up(n * sizeof(int));
  const size_t oVnew = off;    off += up(n * sizeof(float));
  const size_t oOold = off;    off += up(n * sizeof(float));
  const size_t oSnew = off;    off += up(n * sizeof(double));
  const size_t oListA = off;   off += up(n * sizeof(int));
  const size_t oListB = off;   off += up(n * sizeof(int));
  const size_t oInto = off;    off += up(n * sizeof(int));
  const size_t oFrom = off;    off += up(n * sizeof(int));
  const size_t oFind = off;    off += up(n * sizeof(float));
  const size_t oIgn = off;     off += up(n);
  const size_t oLb2 = off;     off += up(n * sizeof(float));
  const size_t oRlb = off;     off += up(n * sizeof(unsigned int));
  const size_t oCtl = off;     off += up(sizeof(HaCtl));
  if ((rc = d.haMisc.reserve(off))) return rc;
  char* base = (char*)d.haMisc.p;
  HaArgs a;
  a.D = (float*)d.haD.p; a.S = linkage == 1 ? (double*)d.haS.p : nullptr;
  a.n = nFrames; a.linkage = linkage; a.target = std::max(1, targetClusters); a.eps = epsilon;
  a.closest = (int*)(base + oClosest); a.cmin = (float*)(base + oCmin); a.ign = (unsigned char*)(base + oIgn);
  a.nfr = (int*)(base + oNfr); a.vnew = (float*)(base + oVnew); a.snew = (double*)(base + oSnew); a.oold = (float*)(base + oOold);
  a.listA = (int*)(base + oListA); a.listB = (int*)(base + oListB); a.ctl = (HaCtl*)(base + oCtl);
  a.mergeInto = (int*)(base + oInto); a.mergeFrom = (int*)(base + oFrom); a.findMin = (float*)(base + oFind);
  a.rkey = (ha_u64*)(base + oKeys); a.lb2 = (float*)(base + oLb2); a.rlb = (unsigned int*)(base + oRlb);
  CU(cudaMemsetAsync(a.ign, 0, n, st));
  COUNT_LAUNCH();
  hieragglo_expand_kernel<<<std::min<int>(nFrames, d.numSMs * 16), 256, 0, st>>>(d_tri, nFrames, a.D, a.S);
  CU(cudaGetLastError());
  COUNT_LAUNCH();
  hieragglo_init_kernel<<<std::min<int>((nFrames + 7) / 8, d.numSMs * 16), 256, 0, st>>>(a);
  CU(cudaGetLastError());
  // one thread-block cluster runs every merge: CTAs per cluster from the cluster count (env B200_HA_TEAM overrides)
  int team = nFrames < 2048 ? 1 : nFrames < 8192 ? 4 : 16;   // (measured: 30,000 clusters, average linkage: 127 / 42 / 23 us per merge with 1 / 4 / 16 CTAs)
  if (const char* e = getenv("B200_HA_TEAM")) team = std::max(1, std::min(16, atoi(e)));
  void (*kern)(HaArgs) = linkage == 0 ? hieragglo_kernel<0> : linkage == 1 ? hieragglo_kernel<1> : hieragglo_kernel<2>;
  if (team > 8) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(team); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = team; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  COUNT_LAUNCH();
  CU(cudaLaunchKernelEx(&cfg, kern, a));
  HaCtl ctl;
  CU(cudaMemcpyAsync(&ctl, a.ctl, sizeof(HaCtl), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  *nCalls = ctl.nCalls; *nMerges = ctl.nMerges;
  if (getenv("B200_HA_DEBUG"))
    fprintf(stderr, "hieragglo n=%d linkage=%d team=%d merges=%d: cycles findmin %lld newrow %lld ignore-rescan %lld update %lld "
            "post-rescan %lld; rows re-scanned before %lld after %lld, tied merges %lld\n", nFrames, linkage, team, ctl.nMerges,
            ctl.clk[0], ctl.clk[1], ctl.clk[2], ctl.clk[3], ctl.clk[4], ctl.cnt[0], ctl.cnt[1], ctl.cnt[2]);
  if (ctl.nMerges > 0) {
    CU(cudaMemcpy(mergeInto, a.mergeInto, (size_t)ctl.nMerges * sizeof(int), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(mergeFrom, a.mergeFrom, (size_t)ctl.nMerges * sizeof(int), cudaMemcpyDeviceToHost));
  }
  if (ctl.nCalls > 0) CU(cudaMemcpy(findMin, a.findMin, (size_t)ctl.nCalls * sizeof(float), cudaMemcpyDeviceToHost));
  {
    std::lock_guard<std::mutex> sl(g_statMu);
    g_stats.h2d_bytes += h2d; g_stats.d2h_bytes += (double)(ctl.nMerges * 8 + ctl.nCalls * 4);
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
  g_engine.store(engine);
  return B200_OK;
}
int b200_set_fixed_point_bits(int bits) {
  if (bits < 0 || bits > 30) return fail(B200_ERR_ARG, "fractional bits must be 0 (automatic) or 1..30");
  g_fixedQs.store(bits);
  return B200_OK;
}
int b200_set_i8_cta_group(int ctaGroup) {
  if (ctaGroup != 1 && ctaGroup != 2) return fail(B200_ERR_ARG, "cta group must be 1 or 2");
  g_i8Cg.store(ctaGroup);
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
  if ((rc = upload_crd(d, d.crd, crd, frameStrideFloats, 0, nFrames, width, st, &h2d))) return rc;
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
