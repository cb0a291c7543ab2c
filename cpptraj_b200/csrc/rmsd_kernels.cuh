// rmsd_kernels.cuh -- sm_100a device code of the B200 best-fit RMSD path.
//
// Kernels (see DESIGN.md for the roofline of each):
//   pack_kernel     float32 AoS COORDS -> centred, sqrt(mass)-scaled FP64 planes in a
//                   fragment-major layout + per-frame G = sum m|x-c|^2.
//                   Replaces CompactFrameArray::GetToMaskDblPtr (src/CompactFrameArray.cpp:244-262)
//                   + Frame::CenterOnOrigin (src/Frame.cpp:1043-1055) + the target centring
//                   of Frame::RMSD_CenteredRef (src/Frame.cpp:1141-1171), once per frame
//                   instead of once per pair.
//   pair_kernel     32x32 frame-pair tiles: the nine F x N . N x F covariance products
//                   (src/Frame.cpp:1184-1208) as FP64 tensor-core MMAs fed by bulk-async
//                   (TMA engine) copies through an mbarrier pipeline, with the per-pair
//                   eigen-solve (src/Frame.cpp:1215-1268, src/Matrix_3x3.cpp:110-268)
//                   fused into the tile epilogue and the float result stored straight
//                   into cpptraj's Matrix<float> layout (src/Matrix.h:94-122).
//   onevn_kernel    one-vs-many rmsd action body (src/Action_Rmsd.cpp:361-392): single
//                   streaming pass, 13 FP64 accumulators per frame, optional U / Trans.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

// ----------------------------------------------------------------------------
// Layout of the packed planes in HBM ("fragment-major"):
//   element(frame f, atom k, plane p) lives at
//     ((rg*nKb + kb)*3 + p)*128 + r*4 + kk      [doubles]
//   rg = f/32, r = f%32, kb = k/4, kk = k%4, nKb = Kpad/4.
// One (rg,kb) block is 3 planes x 32 frames x 4 atoms = 384 doubles = 3072 B and
// contains, for every group of 8 frames, the 8x4 FP64 MMA operand fragment as
// 256 contiguous bytes (lane l of a warp reads double l).  For a fixed rg the
// blocks of consecutive kb are contiguous, so one K-chunk of a 32-frame tile is
// a single contiguous 12 KB run: one cp.async.bulk per operand per stage.
// ----------------------------------------------------------------------------
constexpr int ROWG = 32;                 // frames per row group
constexpr int KBLK = 4;                  // atoms per k-block (MMA k granularity)
constexpr int KC = 16;                   // atoms per pipeline stage
constexpr int KB_PER_CHUNK = KC / KBLK;  // 4
constexpr int BLK_DBL = 3 * ROWG * KBLK; // 384 doubles per (rg,kb) block
constexpr int RG_CHUNK_DBL = KB_PER_CHUNK * BLK_DBL;  // 1536 doubles = 12288 B
constexpr int RG_CHUNK_BYTES = RG_CHUNK_DBL * 8;
constexpr int PAIR_STAGES = 4;
constexpr int PAIR_THREADS = 128;
constexpr int PAIR_SMEM_BYTES = PAIR_STAGES * 2 * RG_CHUNK_BYTES + 64;  // + mbarriers

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
__host__ __device__ inline size_t plane_doubles(int Fpad, int Kpad) {
  return (size_t)(Fpad / ROWG) * (size_t)(Kpad / KBLK) * BLK_DBL;
}
/// Triangle offset of row i's first element (i,i+1): src/Matrix.h:110-122.
__host__ __device__ inline size_t tri_row_start(size_t n, size_t i) {
  return n * i - (i * (i + 1)) / 2;
}

// ----------------------------------------------------------------------------
// PTX helpers
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
/// Blocks until the phase with the given parity has completed.  The suspend-time hint lets the hardware park the
/// thread until the phase completes instead of returning early: without it the retry loop of ~700 waiting threads
/// took ~40 % of the issue slots of the tcgen05 pair kernel (ncu: SYNCS + BRA + YIELD).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
/// 1-D bulk asynchronous copy global -> shared through the TMA engine (SASS: UBLKCP),
/// completion signalled on an mbarrier as transaction bytes.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// FP64 tensor-core MMAs.  Fragment layouts (g = lane/4, t = lane%4):
//   m8n8k4 : a = A[g][t]; b = B[t][g]; c{0,1} = C[g][2t+{0,1}]
//   m16n8kK: a[v0+2*v1] = A[g+8*v0][t+4*v1]; b[v1] = B[t+4*v1][g];
//            c[e+2*h] = C[g+8*h][2t+e]
__device__ __forceinline__ void mma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma_1684(double* c, double a0, double a1, double b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a0), "d"(a1), "d"(b0));
}
__device__ __forceinline__ void mma_1688(double* c, const double* a, const double* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma_16816(double* c, const double* a, const double* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
      "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
        "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// ----------------------------------------------------------------------------
// Per-pair solve.
// cpptraj: rms = sqrt(2*(E0 - sqrt|mu1| - sqrt|mu2| - sig3*sqrt|mu3|)/M) with mu the
// eigenvalues of R R^T by Jacobi and sig3 the sign of det (src/Frame.cpp:1215-1268).
// sigma1+sigma2+sig3*sigma3 is the largest eigenvalue of the 4x4 quaternion key
// matrix of S, i.e. the largest root of
//     P(l) = l^4 + c2 l^2 + c1 l + c0,
//     c2 = -2 ||S||_F^2,  c1 = -8 det S,  c0 = 2 tr((S^T S)^2) - (tr S^T S)^2
// (roots are s1+s2+s3', s1-s2-s3', -s1+s2-s3', -s1-s2+s3').  Newton from the upper
// bound l0 = E0 converges monotonically; all arithmetic FP64.
// ----------------------------------------------------------------------------
struct Quartic {
  double c2, c1, c0;
};
__device__ __forceinline__ Quartic quartic_of(const double* S) {
  const double m00 = S[0] * S[0] + S[3] * S[3] + S[6] * S[6];
  const double m11 = S[1] * S[1] + S[4] * S[4] + S[7] * S[7];
  const double m22 = S[2] * S[2] + S[5] * S[5] + S[8] * S[8];
  const double m01 = S[0] * S[1] + S[3] * S[4] + S[6] * S[7];
  const double m02 = S[0] * S[2] + S[3] * S[5] + S[6] * S[8];
  const double m12 = S[1] * S[2] + S[4] * S[5] + S[7] * S[8];
  const double p1 = m00 + m11 + m22;
  const double trM2 = m00 * m00 + m11 * m11 + m22 * m22 + 2.0 * (m01 * m01 + m02 * m02 + m12 * m12);
  const double det = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) +
                     S[2] * (S[3] * S[7] - S[4] * S[6]);
  Quartic q;
  q.c2 = -2.0 * p1;
  q.c1 = -8.0 * det;
  q.c0 = 2.0 * trM2 - p1 * p1;
  return q;
}
/// Robust path for (near-)double roots of the quartic (collinear / 2-atom selections, or
/// s2 ~ -s3'): singular values of S by one-sided Jacobi (no squaring of the condition number),
/// lambda_max = s1 + s2 + sign(det S) * s3.  Rarely taken; deliberately not inlined.
__device__ __noinline__ double lambda_by_svd(const double* S) {
  double a[3][3];  // a[c] = column c of S
#pragma unroll
  for (int c = 0; c < 3; ++c) { a[c][0] = S[c]; a[c][1] = S[3 + c]; a[c][2] = S[6 + c]; }
  for (int sweep = 0; sweep < 30; ++sweep) {
    bool rotated = false;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
      const double al = a[p][0] * a[p][0] + a[p][1] * a[p][1] + a[p][2] * a[p][2];
      const double be = a[q][0] * a[q][0] + a[q][1] * a[q][1] + a[q][2] * a[q][2];
      const double ga = a[p][0] * a[q][0] + a[p][1] * a[q][1] + a[p][2] * a[q][2];
      if (ga != 0.0 && fabs(ga) > 1e-17 * sqrt(al * be)) {
        const double zeta = (be - al) / (2.0 * ga);
        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const double x = a[p][r], y = a[q][r];
          a[p][r] = c * x - sn * y;
          a[q][r] = sn * x + c * y;
        }
        rotated = true;
      }
    }
    if (!rotated) break;
  }
  double s0 = sqrt(a[0][0] * a[0][0] + a[0][1] * a[0][1] + a[0][2] * a[0][2]);
  double s1 = sqrt(a[1][0] * a[1][0] + a[1][1] * a[1][1] + a[1][2] * a[1][2]);
  double s2 = sqrt(a[2][0] * a[2][0] + a[2][1] * a[2][1] + a[2][2] * a[2][2]);
  const double smin = fmin(s0, fmin(s1, s2));
  const double det = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) +
                     S[2] * (S[3] * S[7] - S[4] * S[6]);
  return (s0 + s1 + s2) - ((det < 0.0) ? 2.0 * smin : 0.0);
}

__device__ __forceinline__ double largest_root(const Quartic& q, double e0, const double* S) {
  double lam = e0, dP = 0.0;
  bool converged = false;
#pragma unroll 1
  for (int it = 0; it < 50; ++it) {
    const double l2 = lam * lam;
    const double b = (l2 + q.c2) * lam;
    const double a = b + q.c1;
    const double P = a * lam + q.c0;
    dP = 2.0 * l2 * lam + b + a;
    if (dP == 0.0) break;
    const double d = P / dP;
    lam -= d;
    if (fabs(d) <= 1e-14 * fabs(lam)) { converged = true; break; }
  }
  // Conditioning of the root: rounding noise of P (~ 64 eps lam^4) over P' must stay below
  // ~1e-12 lam for the 1e-4 A contract; otherwise (double root) take the SVD path.
  if (!converged || fabs(dP) < 7e-3 * fabs(lam * lam * lam)) lam = lambda_by_svd(S);
  return lam;
}
/// Best-fit RMSD from the 3x3 covariance S, E0 = (Ga+Gb)/2 and total mass M.
__device__ __forceinline__ double rmsd_fit_from_cov(const double* S, double e0, double invM) {
  const Quartic q = quartic_of(S);
  const double lam = largest_root(q, e0, S);
  const double e = e0 - lam;
  return (e < 0.0) ? 0.0 : sqrt(2.0 * e * invM);  // clamp: src/Frame.cpp:1264-1268
}
/// No-fit RMSD (src/Frame.cpp:1279-1307) from G's and the covariance trace.
__device__ __forceinline__ double rmsd_nofit_from_trace(double tr, double ga, double gb, double invM) {
  const double s = ga + gb - 2.0 * tr;
  return (s < 0.0) ? 0.0 : sqrt(s * invM);
}

// ----------------------------------------------------------------------------
// pack_kernel
// One CTA per 32-frame row group, 8 warps x 4 frames.  Per frame a warp makes
// two passes over the selected atoms (second pass hits L1/L2): centre (FP64
// shuffle reduction), then centred/scaled values are written into the
// fragment-major planes; padding atoms (k >= nAtoms) and padding frames are
// written as zeros so the pair kernel needs no bounds checks in its main loop.
// ----------------------------------------------------------------------------
struct PackArgs {
  const float* crd;        // device COORDS
  size_t stride;           // floats per source frame
  const int* frameIdx;     // nullable: source frame of output frame f
  long srcBase;            // source row = frameIdx ? frameIdx[f] - srcBase : f - srcBase
  int nFrames;             // valid output frames F
  int f0;                  // first output frame of this launch (multiple of 32)
  const int* atomIdx;      // nullable => identity
  int nAtoms;
  int Kpad;
  const double* centerMass;  // nullable: weights for the centre
  const double* covMass;     // nullable: planes scaled by sqrt(covMass)
  const double* shift;       // 3 doubles (device); used when !fit
  int fit;
  double* planes;
  double* G;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) pack_kernel(PackArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rg = a.f0 / ROWG + blockIdx.x;
  const int nKb = a.Kpad / KBLK;
  double* rgBase = a.planes + (size_t)rg * nKb * BLK_DBL;
  for (int rr = 0; rr < 4; ++rr) {
    const int r = warp * 4 + rr;
    const int f = rg * ROWG + r;
    const bool valid = f < a.nFrames;
    const float* src = nullptr;
    if (valid) {
      const long row = (a.frameIdx ? (long)a.frameIdx[f] : (long)f) - a.srcBase;
      src = a.crd + (size_t)row * a.stride;
    }
    double cx = 0.0, cy = 0.0, cz = 0.0;
    if (valid) {
      if (a.fit) {
        double sx = 0.0, sy = 0.0, sz = 0.0, sm = 0.0;
        for (int k = lane; k < a.nAtoms; k += 32) {
          const int at = a.atomIdx ? a.atomIdx[k] : k;
          const double m = a.centerMass ? a.centerMass[k] : 1.0;
          const double x = (double)src[3 * (size_t)at], y = (double)src[3 * (size_t)at + 1],
                       z = (double)src[3 * (size_t)at + 2];
          sx += x * m; sy += y * m; sz += z * m; sm += m;
        }
        sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz); sm = warp_sum(sm);
        if (sm != 0.0) { cx = sx / sm; cy = sy / sm; cz = sz / sm; }
      } else {
        cx = a.shift[0]; cy = a.shift[1]; cz = a.shift[2];
      }
    }
    double g = 0.0;
    for (int k = lane; k < a.Kpad; k += 32) {
      double x = 0.0, y = 0.0, z = 0.0;
      if (valid && k < a.nAtoms) {
        const int at = a.atomIdx ? a.atomIdx[k] : k;
        const double w = a.covMass ? sqrt(a.covMass[k]) : 1.0;
        x = ((double)src[3 * (size_t)at] - cx) * w;
        y = ((double)src[3 * (size_t)at + 1] - cy) * w;
        z = ((double)src[3 * (size_t)at + 2] - cz) * w;
        g += x * x + y * y + z * z;
      }
      double* blk = rgBase + (size_t)(k >> 2) * BLK_DBL + r * KBLK + (k & 3);
      blk[0] = x;
      blk[ROWG * KBLK] = y;
      blk[2 * ROWG * KBLK] = z;
    }
    g = warp_sum(g);
    if (lane == 0) a.G[f] = g;
  }
}

/// shift[0..2] = position of the first selected atom of the first frame (no-fit mode:
/// a common origin keeps |x| small so G_i + G_j - 2 tr(S) loses no digits).
__global__ void shift_kernel(const float* crd, size_t stride, const int* frameIdx, long srcBase, int fFirst,
                             const int* atomIdx, double* shift) {
  const long row = (frameIdx ? (long)frameIdx[fFirst] : (long)fFirst) - srcBase;
  const int at = atomIdx ? atomIdx[0] : 0;
  const float* s = crd + (size_t)row * stride + 3 * (size_t)at;
  shift[0] = (double)s[0]; shift[1] = (double)s[1]; shift[2] = (double)s[2];
}
/// totalMass[0] = sum of mass (or nAtoms)
__global__ void mass_sum_kernel(const double* mass, int n, double* total) {
  double s = 0.0;
  for (int k = threadIdx.x; k < n; k += 32) s += mass ? mass[k] : 1.0;
  s = warp_sum(s);
  if (threadIdx.x == 0) total[0] = s;
}

// ----------------------------------------------------------------------------
// pair_kernel
// ----------------------------------------------------------------------------
struct PairArgs {
  const double* PA;   // planes of the row frames (i)
  const double* PB;   // planes of the column frames (j)
  const double* GA;
  const double* GB;
  int nKb;            // Kpad/4
  int nRows;          // valid i frames
  int nCols;          // valid j frames
  int rg0;            // first row group (i tiles) of this launch
  int nRgI;           // number of i row groups in this launch
  int cg0;            // first column group covered by blockIdx.x == 0
  const double* totalMass;  // device scalar
  float* out;         // TRI: base such that out[triIndex - outBase]; FULL: out[i*ldo + j]
  size_t outBase;
  size_t ldo;
};

template <int VAR, bool FIT, bool TRI>
__global__ void __launch_bounds__(PAIR_THREADS, 2) pair_kernel(PairArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* smem = reinterpret_cast<double*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + PAIR_STAGES * 2 * RG_CHUNK_BYTES);

  const int it = a.rg0 + blockIdx.y;        // i row group
  const int jt = a.cg0 + blockIdx.x;        // j row group
  if (TRI && jt < it) return;               // below the diagonal: nothing to do
  if (jt * ROWG >= a.nCols) return;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wi = warp >> 1, wj = warp & 1;
  const int nChunks = a.nKb / KB_PER_CHUNK;
  const double* gA = a.PA + (size_t)it * a.nKb * BLK_DBL;
  const double* gB = a.PB + (size_t)jt * a.nKb * BLK_DBL;

  if (tid == 0) {
    for (int s = 0; s < PAIR_STAGES; ++s) mbar_init(smem_u32(&bars[s]), 1);
    mbar_fence_init();
  }
  __syncthreads();

  auto issue = [&](int c) {
    const int s = c % PAIR_STAGES;
    const uint32_t bar = smem_u32(&bars[s]);
    mbar_expect_tx(bar, 2 * RG_CHUNK_BYTES);
    bulk_g2s(smem_u32(smem + (size_t)s * 2 * RG_CHUNK_DBL), gA + (size_t)c * RG_CHUNK_DBL, RG_CHUNK_BYTES, bar);
    bulk_g2s(smem_u32(smem + (size_t)s * 2 * RG_CHUNK_DBL + RG_CHUNK_DBL), gB + (size_t)c * RG_CHUNK_DBL,
             RG_CHUNK_BYTES, bar);
  };
  if (tid == 0) {
    for (int c = 0; c < PAIR_STAGES - 1 && c < nChunks; ++c) issue(c);
  }

  constexpr int NP = FIT ? 9 : 3;
  double acc[NP][2][4];
#pragma unroll
  for (int x = 0; x < NP; ++x)
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[x][n][e] = 0.0;

  // number of k-blocks consumed per MMA step
  constexpr int KB_STEP = (VAR <= 1) ? 1 : (VAR == 2 ? 2 : 4);

  for (int c = 0; c < nChunks; ++c) {
    if (tid == 0 && c + PAIR_STAGES - 1 < nChunks) issue(c + PAIR_STAGES - 1);
    const int s = c % PAIR_STAGES;
    mbar_wait(smem_u32(&bars[s]), (uint32_t)((c / PAIR_STAGES) & 1));
    const double* sA = smem + (size_t)s * 2 * RG_CHUNK_DBL + (wi * 2) * 32 + lane;
    const double* sB = smem + (size_t)s * 2 * RG_CHUNK_DBL + RG_CHUNK_DBL + (wj * 2) * 32 + lane;
#pragma unroll
    for (int kb = 0; kb < KB_PER_CHUNK; kb += KB_STEP) {
      // fragments: fa[p][v1][v0], fb[q][nj][v1]
      double fa[3][KB_STEP][2], fb[3][2][KB_STEP];
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int v1 = 0; v1 < KB_STEP; ++v1) {
          fa[p][v1][0] = sA[((kb + v1) * 3 + p) * 128];
          fa[p][v1][1] = sA[((kb + v1) * 3 + p) * 128 + 32];
          fb[p][0][v1] = sB[((kb + v1) * 3 + p) * 128];
          fb[p][1][v1] = sB[((kb + v1) * 3 + p) * 128 + 32];
        }
#pragma unroll
      for (int p = 0; p < 3; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          if (!FIT && p != q) continue;
          const int x = FIT ? (p * 3 + q) : p;
#pragma unroll
          for (int nj = 0; nj < 2; ++nj) {
            if constexpr (VAR == 0) {
              mma_884(acc[x][nj][0], acc[x][nj][1], fa[p][0][0], fb[q][nj][0]);
              mma_884(acc[x][nj][2], acc[x][nj][3], fa[p][0][1], fb[q][nj][0]);
            } else if constexpr (VAR == 1) {
              mma_1684(acc[x][nj], fa[p][0][0], fa[p][0][1], fb[q][nj][0]);
            } else if constexpr (VAR == 2) {
              mma_1688(acc[x][nj], &fa[p][0][0], &fb[q][nj][0]);
            } else {
              mma_16816(acc[x][nj], &fa[p][0][0], &fb[q][nj][0]);
            }
          }
        }
    }
    __syncthreads();  // everyone is done with stage s before it is refilled
  }

  // ---- epilogue: per-pair solve, straight into cpptraj's matrix layout ----
  // total mass < Constants::SMALL: the reference reports -1 for every pair (src/Frame.cpp:1160-1163, :1300-1303)
  const bool massOk = a.totalMass[0] >= 1e-14;
  const double invM = massOk ? 1.0 / a.totalMass[0] : 0.0;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nj = 0; nj < 2; ++nj)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int i = it * ROWG + wi * 16 + h * 8 + g;
        const int j = jt * ROWG + wj * 16 + nj * 8 + 2 * t + e;
        if (i >= a.nRows || j >= a.nCols) continue;
        if (TRI && j <= i) continue;
        const double ga = a.GA[i], gb = a.GB[j];
        double r;
        if (FIT) {
          double S[9];
#pragma unroll
          for (int x = 0; x < 9; ++x) S[x] = acc[FIT ? x : 0][nj][e + 2 * h];
          r = rmsd_fit_from_cov(S, 0.5 * (ga + gb), invM);
        } else {
          r = rmsd_nofit_from_trace(acc[0][nj][e + 2 * h] + acc[1][nj][e + 2 * h] + acc[2][nj][e + 2 * h], ga,
                                    gb, invM);
        }
        if (!massOk) r = -1.0;
        size_t idx;
        if (TRI)
          idx = tri_row_start((size_t)a.nCols, (size_t)i) + (size_t)(j - i - 1) - a.outBase;
        else
          idx = (size_t)i * a.ldo + (size_t)j;
        a.out[idx] = (float)r;
      }
}

// ----------------------------------------------------------------------------
// onevn_kernel: FB frames per CTA share each reference load.
//   S = sum m (x - c) r^T = sum m x r^T - c (sum m r)^T,   G_t = sum m |x|^2 - M |c|^2
// computed relative to the frame's first selected atom so the subtraction loses
// no digits.  T = float (COORDS) or double (cpptraj Frame).
// ----------------------------------------------------------------------------
constexpr int ONEVN_FB = 4;
constexpr int ONEVN_THREADS = 256;

struct OneVNArgs {
  const void* crd;
  size_t stride;            // elements per frame
  const int* frameIdx;      // nullable: frame n is row frameIdx[n] - srcBase of crd (sieved frame lists)
  long srcBase;
  int nFrames;
  const int* atomIdx;       // nullable
  int nAtoms;
  const double* refw;       // 4 doubles per atom: rx, ry, rz, m   (ref as given)
  const double* refsum;     // [0..2] = sum m r, [3] = M, [4] = sum m |r|^2
  const int* skipIf;        // nullable: hdr of the streaming variant; [2] != 0 => that variant did the work
  int fit;
  double* rmsd;
  double* rot;              // nullable, 9 per frame
  double* trans;            // nullable, 3 per frame
};

/// Rotation U (row-major, x' = U x as applied by Frame::Trans_Rot_Trans, src/Frame.h:572-581)
/// from the covariance S[a][b] = sum m xt_a xr_b and its largest quartic root lam:
/// eigenvector of the key matrix via the adjugate of (K - lam I), most stable column.
__device__ inline void rotation_from_cov(const double* S, double lam, double* U) {
  const double Sxx = S[0], Sxy = S[1], Sxz = S[2], Syx = S[3], Syy = S[4], Syz = S[5], Szx = S[6],
               Szy = S[7], Szz = S[8];
  // Key matrix for rotating the TARGET onto the REFERENCE.
  double K[4][4];
  K[0][0] = Sxx + Syy + Szz - lam;
  K[0][1] = K[1][0] = Syz - Szy;
  K[0][2] = K[2][0] = Szx - Sxz;
  K[0][3] = K[3][0] = Sxy - Syx;
  K[1][1] = Sxx - Syy - Szz - lam;
  K[1][2] = K[2][1] = Sxy + Syx;
  K[1][3] = K[3][1] = Szx + Sxz;
  K[2][2] = -Sxx + Syy - Szz - lam;
  K[2][3] = K[3][2] = Syz + Szy;
  K[3][3] = -Sxx - Syy + Szz - lam;
  // adjugate columns (cofactors); pick the column with the largest norm
  double best[4] = {1.0, 0.0, 0.0, 0.0}, bestn = -1.0;
  for (int col = 0; col < 4; ++col) {
    double v[4];
    for (int row = 0; row < 4; ++row) {
      // cofactor C[col][row] = (-1)^(row+col) * minor(col,row); adj = C^T, K symmetric
      int rI[3], cI[3], n = 0, m = 0;
      for (int x = 0; x < 4; ++x) if (x != col) rI[n++] = x;
      for (int x = 0; x < 4; ++x) if (x != row) cI[m++] = x;
      const double d = K[rI[0]][cI[0]] * (K[rI[1]][cI[1]] * K[rI[2]][cI[2]] - K[rI[1]][cI[2]] * K[rI[2]][cI[1]]) -
                       K[rI[0]][cI[1]] * (K[rI[1]][cI[0]] * K[rI[2]][cI[2]] - K[rI[1]][cI[2]] * K[rI[2]][cI[0]]) +
                       K[rI[0]][cI[2]] * (K[rI[1]][cI[0]] * K[rI[2]][cI[1]] - K[rI[1]][cI[1]] * K[rI[2]][cI[0]]);
      v[row] = ((row + col) & 1) ? -d : d;
    }
    const double nn = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
    if (nn > bestn) { bestn = nn; best[0] = v[0]; best[1] = v[1]; best[2] = v[2]; best[3] = v[3]; }
  }
  double q0 = best[0], q1 = best[1], q2 = best[2], q3 = best[3];
  const double nrm = sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
  if (nrm > 0.0) { q0 /= nrm; q1 /= nrm; q2 /= nrm; q3 /= nrm; } else { q0 = 1.0; q1 = q2 = q3 = 0.0; }
  const double a2 = q0 * q0, x2 = q1 * q1, y2 = q2 * q2, z2 = q3 * q3;
  const double xy = q1 * q2, az = q0 * q3, zx = q3 * q1, ay = q0 * q2, yz = q2 * q3, ax = q0 * q1;
  U[0] = a2 + x2 - y2 - z2; U[1] = 2 * (xy - az);      U[2] = 2 * (zx + ay);
  U[3] = 2 * (xy + az);     U[4] = a2 - x2 + y2 - z2;  U[5] = 2 * (yz - ax);
  U[6] = 2 * (zx - ay);     U[7] = 2 * (yz + ax);      U[8] = a2 - x2 - y2 + z2;
}

template <typename T>
__global__ void __launch_bounds__(ONEVN_THREADS) onevn_kernel(OneVNArgs a) {
  __shared__ double red[ONEVN_THREADS / 32][ONEVN_FB][14];
  if (a.skipIf && a.skipIf[2] != 0) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nGroups = (a.nFrames + ONEVN_FB - 1) / ONEVN_FB;
  for (int grp = blockIdx.x; grp < nGroups; grp += gridDim.x) {
  if (grp != (int)blockIdx.x) __syncthreads();   // (red[] of the previous group has been read)
  const int fbase = grp * ONEVN_FB;
  const T* crd = reinterpret_cast<const T*>(a.crd);
  const T* src[ONEVN_FB];
  double o[ONEVN_FB][3];
  const int at0 = a.atomIdx ? a.atomIdx[0] : 0;
#pragma unroll
  for (int f = 0; f < ONEVN_FB; ++f) {
    const int fr = min(fbase + f, a.nFrames - 1);
    const size_t row = a.frameIdx ? (size_t)((long)a.frameIdx[fr] - a.srcBase) : (size_t)fr;
    src[f] = crd + row * a.stride;
    o[f][0] = (double)src[f][3 * (size_t)at0];
    o[f][1] = (double)src[f][3 * (size_t)at0 + 1];
    o[f][2] = (double)src[f][3 * (size_t)at0 + 2];
  }
  // acc[f][0..8] = sum m x_a r_b (fit) ; [9..11] = sum m x ; [12] = sum m |x|^2
  // nofit: [0] = sum m |r - x|^2 (direct, as the reference does)
  double acc[ONEVN_FB][13];
#pragma unroll
  for (int f = 0; f < ONEVN_FB; ++f)
#pragma unroll
    for (int x = 0; x < 13; ++x) acc[f][x] = 0.0;

  for (int k = tid; k < a.nAtoms; k += ONEVN_THREADS) {
    const int at = a.atomIdx ? a.atomIdx[k] : k;
    const double4 rw = reinterpret_cast<const double4*>(a.refw)[k];
#pragma unroll
    for (int f = 0; f < ONEVN_FB; ++f) {
      const T* p = src[f] + 3 * (size_t)at;
      if (a.fit) {
        const double x = (double)p[0] - o[f][0], y = (double)p[1] - o[f][1], z = (double)p[2] - o[f][2];
        const double mx = rw.w * x, my = rw.w * y, mz = rw.w * z;
        acc[f][0] += mx * rw.x; acc[f][1] += mx * rw.y; acc[f][2] += mx * rw.z;
        acc[f][3] += my * rw.x; acc[f][4] += my * rw.y; acc[f][5] += my * rw.z;
        acc[f][6] += mz * rw.x; acc[f][7] += mz * rw.y; acc[f][8] += mz * rw.z;
        acc[f][9] += mx; acc[f][10] += my; acc[f][11] += mz;
        acc[f][12] += mx * x + my * y + mz * z;
      } else {
        const double dx = rw.x - (double)p[0], dy = rw.y - (double)p[1], dz = rw.z - (double)p[2];
        acc[f][0] += rw.w * (dx * dx + dy * dy + dz * dz);
      }
    }
  }
  const int nred = a.fit ? 13 : 1;
#pragma unroll
  for (int f = 0; f < ONEVN_FB; ++f)
    for (int x = 0; x < nred; ++x) {
      const double v = warp_sum(acc[f][x]);
      if (lane == 0) red[warp][f][x] = v;
    }
  __syncthreads();
  if (tid < ONEVN_FB) {
    const int f = tid, fr = fbase + f;
    if (fr < a.nFrames) {
      double v[13];
      for (int x = 0; x < nred; ++x) {
        double s = 0.0;
        for (int w = 0; w < ONEVN_THREADS / 32; ++w) s += red[w][f][x];
        v[x] = s;
      }
      const double M = a.refsum[3];
      if (M < 1e-14) {   // src/Frame.cpp:1160-1163, :1300-1303
        a.rmsd[fr] = -1.0;
      } else if (!a.fit) {
        a.rmsd[fr] = (v[0] < 0.0) ? 0.0 : sqrt(v[0] / M);
      } else {
        // centre relative to the frame origin o
        const double cx = v[9] / M, cy = v[10] / M, cz = v[11] / M;
        double S[9];
        S[0] = v[0] - cx * a.refsum[0]; S[1] = v[1] - cx * a.refsum[1]; S[2] = v[2] - cx * a.refsum[2];
        S[3] = v[3] - cy * a.refsum[0]; S[4] = v[4] - cy * a.refsum[1]; S[5] = v[5] - cy * a.refsum[2];
        S[6] = v[6] - cz * a.refsum[0]; S[7] = v[7] - cz * a.refsum[1]; S[8] = v[8] - cz * a.refsum[2];
        const double gt = v[12] - M * (cx * cx + cy * cy + cz * cz);
        const double e0 = 0.5 * (gt + a.refsum[4]);
        const Quartic q = quartic_of(S);
        const double lam = largest_root(q, e0, S);
        const double e = e0 - lam;
        a.rmsd[fr] = (e < 0.0) ? 0.0 : sqrt(2.0 * e / M);
        if (a.rot) rotation_from_cov(S, lam, a.rot + 9 * (size_t)fr);
        if (a.trans) {
          // Trans = -(centre of the target) in absolute coordinates (src/Frame.cpp:1164-1170)
          a.trans[3 * (size_t)fr] = -(cx + o[f][0]);
          a.trans[3 * (size_t)fr + 1] = -(cy + o[f][1]);
          a.trans[3 * (size_t)fr + 2] = -(cz + o[f][2]);
        }
      }
    }
  }
  }   // frame groups
}

// ----------------------------------------------------------------------------
// One-vs-many, streaming variant (the HBM-bound path of Action_Rmsd::DoAction, src/Action_Rmsd.cpp:361-417).
// Frames are contiguous rows of COORDS, so the span of a frame that holds the selected atoms is moved by the
// TMA engine (cp.async.bulk, SASS UBLKCP) into a 2-stage shared-memory ring -- 4 frames x 12 KB plus the chunk's
// reference atoms (32 KB) and atom numbers per stage -- one thread issuing the next stage while the 256 threads work
// on the current one: every operand of the inner loop comes from shared memory (ncu: the L2 latency of the
// reference / index loads was the top stall when they were ordinary loads).
// The 256 threads gather the selected atoms of the staged chunk from shared memory (conflict-free: consecutive
// atoms are 12 or 24 bytes apart), accumulate the same 13 FP64 sums per frame as onevn_kernel and write them
// (plus the frame's origin shift) to a 16-double record per frame; onevn_finish_kernel (one thread per frame)
// does the per-frame solve, rotation and translation.  Needs a sorted selection (AtomMask::Selected() is) and a
// 16-byte aligned COORDS base; the host checks the alignment, the device the order (hdr[2]), and onevn_kernel
// remains the general path.
// ----------------------------------------------------------------------------
constexpr int ONEVN_S_STAGES = 2;                                // (5 stages of half the size measured slower: per-step costs dominate)
constexpr int ONEVN_S_CHUNK_BYTES = 12288;                       // per frame and stage
constexpr int ONEVN_S_BUF_BYTES = ONEVN_S_CHUNK_BYTES + 32;      // + alignment slack at both ends
constexpr int ONEVN_S_MAX_APC = ONEVN_S_CHUNK_BYTES / 12;        // selected atoms a chunk can hold (float frames)
constexpr int ONEVN_S_REF_OFF = ONEVN_FB * ONEVN_S_BUF_BYTES;    // reference (rx, ry, rz, m) of the chunk's selected atoms
constexpr int ONEVN_S_IDX_OFF = ONEVN_S_REF_OFF + ONEVN_S_MAX_APC * 32;   // their atom numbers
constexpr int ONEVN_S_STAGE_BYTES = ONEVN_S_IDX_OFF + ONEVN_S_MAX_APC * 4 + 32;
constexpr int ONEVN_S_SMEM_BYTES = ONEVN_S_STAGES * ONEVN_S_STAGE_BYTES + 64;
constexpr int ONEVN_REC = 16;                                    // doubles per frame record: 13 sums + origin

struct OneVNStreamArgs {
  const void* crd;
  size_t stride;            // elements per frame
  const int* frameIdx;      // nullable: frame n is row frameIdx[n] - srcBase of crd
  long srcBase;
  int nFrames;
  const int* atomIdx;       // nullable (identity)
  int nAtoms;
  const double* refw;       // 4 doubles per selected atom: rx, ry, rz, m
  const int* hdr;           // [0] first atom of the span, [1] chunks per frame, [2] 1 = selection sorted (usable)
  const int* kLo;           // [chunks + 1]: first selected-atom index of every chunk
  int fit;
  double* rec;              // ONEVN_REC doubles per frame
};

/// Chunk table of the streaming kernel.  One block.  atomsPerChunk = ONEVN_S_CHUNK_BYTES / (3 * sizeof(T)).
__device__ __forceinline__ void onevn_chunks_body(const int* atomIdx, int nAtoms, int atomsPerChunk, int maxChunks,
                                                  int* hdr, int* kLo) {
  // One pass, no dependent loads: a sorted selection starts at its first and ends at its last entry; atom k opens every
  // chunk between its predecessor's chunk (exclusive) and its own (inclusive): kLo[c] = first k with atom >= a0 + c * APC.
  __shared__ int sSorted;
  if (threadIdx.x == 0) sSorted = 1;
  __syncthreads();
  const int a0 = atomIdx ? atomIdx[0] : 0;
  const int aMax = atomIdx ? atomIdx[nAtoms - 1] : nAtoms - 1;
  const int nCh = aMax >= a0 ? (aMax - a0) / atomsPerChunk + 1 : 0;
  int ok = (aMax >= a0) ? 1 : 0;
  for (int k = threadIdx.x; k < nAtoms; k += blockDim.x) {
    const int a = atomIdx ? atomIdx[k] : k;
    const int prev = k > 0 ? (atomIdx ? atomIdx[k - 1] : k - 1) : a0 - 1;
    if (k > 0 && prev >= a) ok = 0;
    if (a < a0 || a > aMax) { ok = 0; continue; }
    const int cHi = (a - a0) / atomsPerChunk;
    const int cLo = (k > 0 && prev >= a0) ? (prev - a0) / atomsPerChunk + 1 : 0;
    for (int c = cLo; c <= cHi && c <= maxChunks; ++c) kLo[c] = k;
  }
  if (threadIdx.x == 0 && nCh >= 0 && nCh <= maxChunks) kLo[nCh] = nAtoms;
  if (!ok) atomicAnd(&sSorted, 0);
  __syncthreads();
  const bool usable = sSorted && nCh >= 1 && nCh <= maxChunks;
  if (threadIdx.x == 0) { hdr[0] = a0; hdr[1] = usable ? nCh : 0; hdr[2] = usable ? 1 : 0; }
}
__global__ void __launch_bounds__(256) onevn_chunks_kernel(const int* atomIdx, int nAtoms, int atomsPerChunk, int maxChunks,
                                                           int* hdr, int* kLo) {
  onevn_chunks_body(atomIdx, nAtoms, atomsPerChunk, maxChunks, hdr, kLo);
}

constexpr int ONEVN_S_THREADS = 2 * ONEVN_THREADS;   // 16 warps: threads [0,256) take frames 0,1 of the stage, [256,512) frames 2,3
constexpr int ONEVN_S_FT = ONEVN_FB / 2;             // frames per thread (more warps per scheduler hide the FP64 / LDS latencies)

template <typename T>
__global__ void __launch_bounds__(ONEVN_S_THREADS, 1) onevn_stream_kernel(OneVNStreamArgs a) {
  if (a.hdr[2] == 0) return;   // selection not sorted: onevn_kernel (launched next) does the work
  constexpr int APC = ONEVN_S_CHUNK_BYTES / (3 * (int)sizeof(T));   // atoms per chunk
  extern __shared__ __align__(128) unsigned char smem_ov[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_ov + ONEVN_S_STAGES * ONEVN_S_STAGE_BYTES);
  __shared__ double red[ONEVN_S_THREADS / 32][ONEVN_S_FT][13];
  __shared__ double oS[ONEVN_FB][3];   // origin shifts of the current frame group (for the records; `o` stays in registers)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = tid / ONEVN_THREADS, tk = tid % ONEVN_THREADS;   // frame pair of this thread, index within the pair's 256 threads
  const int a0 = a.hdr[0], nCh = a.hdr[1];
  const int nGroups = (a.nFrames + ONEVN_FB - 1) / ONEVN_FB;
  const int myGroups = ((int)blockIdx.x < nGroups) ? (nGroups - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const long total = (long)myGroups * nCh;   // (group, chunk) steps of this CTA
  const unsigned char* base = reinterpret_cast<const unsigned char*>(a.crd);
  const int spanAtoms = a.kLo[nCh] > 0 ? ((a.atomIdx ? a.atomIdx[a.nAtoms - 1] : a.nAtoms - 1) - a0 + 1) : 0;
  if (tid == 0) {
    for (int s = 0; s < ONEVN_S_STAGES; ++s) mbar_init(smem_u32(&full[s]), 1);
    mbar_fence_init();
  }
  __syncthreads();

  // Producer side (thread 0): stage `step` of this CTA's sequence.  Bytes below a 16-byte boundary at the end of a
  // chunk (at most 3 floats) are copied with ordinary loads; the barrier at the end of every step orders them.
  // Whole warp 0: lanes 0..3 copy one frame's chunk each, lane 4 the reference atoms, lane 5 the atom numbers (a 1-D
  // bulk copy costs ~225 cycles of issue whatever its size: six lanes issue side by side instead of one after the other).
  auto issue = [&](long step) {
    const int g = (int)blockIdx.x + (int)(step / nCh) * (int)gridDim.x;
    const int c = (int)(step % nCh);
    const int st = (int)(step % ONEVN_S_STAGES);
    const int nAt = min(APC, spanAtoms - c * APC);
    unsigned char* sbuf = smem_ov + st * ONEVN_S_STAGE_BYTES;
    const int kA = a.kLo[c], nSel = a.kLo[c + 1] - kA;
    uint32_t bytes = 0, len = 0, dstOff = 0;
    const unsigned char* src = nullptr;
    if (lane < ONEVN_FB) {
      const int fr = min(g * ONEVN_FB + lane, a.nFrames - 1);
      const size_t row = a.frameIdx ? (size_t)((long)a.frameIdx[fr] - a.srcBase) : (size_t)fr;
      const size_t off = (row * a.stride + (size_t)3 * ((size_t)a0 + (size_t)c * APC)) * sizeof(T);
      const size_t s0 = off & ~(size_t)15;
      len = (uint32_t)(off - s0) + (uint32_t)nAt * 3u * (uint32_t)sizeof(T);
      src = base + s0; dstOff = (uint32_t)lane * ONEVN_S_BUF_BYTES;
    } else if (lane == ONEVN_FB) {
      len = (uint32_t)nSel * 32u;
      src = reinterpret_cast<const unsigned char*>(a.refw) + (size_t)kA * 32; dstOff = ONEVN_S_REF_OFF;
    } else if (lane == ONEVN_FB + 1 && a.atomIdx && nSel > 0) {
      const size_t off = (size_t)kA * 4, s0 = off & ~(size_t)15;
      len = (uint32_t)(off - s0) + (uint32_t)nSel * 4u;
      src = reinterpret_cast<const unsigned char*>(a.atomIdx) + s0; dstOff = ONEVN_S_IDX_OFF;
    }
    bytes = len & ~15u;
    for (uint32_t t = bytes; t < len; t += 4)   // tail below a 16-byte boundary: whole 4-byte words, ordinary loads
      *reinterpret_cast<uint32_t*>(sbuf + dstOff + t) = *reinterpret_cast<const uint32_t*>(src + t);
    uint32_t tot = bytes;
#pragma unroll
    for (int o2 = 4; o2 > 0; o2 >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o2);   // lanes 0..7 hold the sum of lanes 0..7
    const uint32_t bar = smem_u32(&full[st]);
    if (lane == 0) mbar_expect_tx(bar, tot);
    __syncwarp();
    if (bytes) bulk_g2s(smem_u32(sbuf + dstOff), src, bytes, bar);
    __syncwarp();
  };
  if (warp == 0)
    for (long s = 0; s < ONEVN_S_STAGES - 1 && s < total; ++s) issue(s);
  __syncthreads();   // the producer's ordinary (tail) stores of the first stages are visible to everyone

  double acc[ONEVN_S_FT][13];
  double o[ONEVN_S_FT][3];
  for (long step = 0; step < total; ++step) {
    const int gi = (int)(step / nCh), c = (int)(step % nCh);
    const int g = (int)blockIdx.x + gi * (int)gridDim.x;
    const int st = (int)(step % ONEVN_S_STAGES);
    if (c == 0) {
      const int at0 = a.atomIdx ? a.atomIdx[0] : 0;
#pragma unroll
      for (int f = 0; f < ONEVN_S_FT; ++f) {
        const int fr = min(g * ONEVN_FB + ONEVN_S_FT * half + f, a.nFrames - 1);
        const size_t rowI = a.frameIdx ? (size_t)((long)a.frameIdx[fr] - a.srcBase) : (size_t)fr;
        const T* row = reinterpret_cast<const T*>(a.crd) + rowI * a.stride + (size_t)3 * at0;
        o[f][0] = (double)row[0]; o[f][1] = (double)row[1]; o[f][2] = (double)row[2];
        if (tk == 0) { oS[ONEVN_S_FT * half + f][0] = o[f][0]; oS[ONEVN_S_FT * half + f][1] = o[f][1]; oS[ONEVN_S_FT * half + f][2] = o[f][2]; }   // read after the group's barriers
#pragma unroll
        for (int x = 0; x < 13; ++x) acc[f][x] = 0.0;
      }
    }
    if (warp == 0 && step + ONEVN_S_STAGES - 1 < total) issue(step + ONEVN_S_STAGES - 1);
    const int chunkA0 = a0 + c * APC;
    const int kA = a.kLo[c], nSel = a.kLo[c + 1] - kA;
    mbar_wait(smem_u32(&full[st]), (uint32_t)((step / ONEVN_S_STAGES) & 1));
    const unsigned char* sbuf = smem_ov + st * ONEVN_S_STAGE_BYTES;
    const T* fp[ONEVN_S_FT];
#pragma unroll
    for (int f = 0; f < ONEVN_S_FT; ++f) {
      const int fr = min(g * ONEVN_FB + ONEVN_S_FT * half + f, a.nFrames - 1);
      const size_t row = a.frameIdx ? (size_t)((long)a.frameIdx[fr] - a.srcBase) : (size_t)fr;
      const size_t off = (row * a.stride + (size_t)3 * ((size_t)a0 + (size_t)c * APC)) * sizeof(T);
      fp[f] = reinterpret_cast<const T*>(sbuf + (ONEVN_S_FT * half + f) * ONEVN_S_BUF_BYTES + (off & 15));
    }
    const double4* sref = reinterpret_cast<const double4*>(sbuf + ONEVN_S_REF_OFF);
    const int* sidx = reinterpret_cast<const int*>(sbuf + ONEVN_S_IDX_OFF + (((size_t)kA * 4) & 15));
    for (int kk = tk; kk < nSel; kk += ONEVN_THREADS) {
      const int at = (a.atomIdx ? sidx[kk] : kA + kk) - chunkA0;
      const double4 rw = sref[kk];
#pragma unroll
      for (int f = 0; f < ONEVN_S_FT; ++f) {
        const T* p = fp[f] + 3 * at;
        if (a.fit) {
          const double x = (double)p[0] - o[f][0], y = (double)p[1] - o[f][1], z = (double)p[2] - o[f][2];
          const double mx = rw.w * x, my = rw.w * y, mz = rw.w * z;
          acc[f][0] += mx * rw.x; acc[f][1] += mx * rw.y; acc[f][2] += mx * rw.z;
          acc[f][3] += my * rw.x; acc[f][4] += my * rw.y; acc[f][5] += my * rw.z;
          acc[f][6] += mz * rw.x; acc[f][7] += mz * rw.y; acc[f][8] += mz * rw.z;
          acc[f][9] += mx; acc[f][10] += my; acc[f][11] += mz;
          acc[f][12] += mx * x + my * y + mz * z;
        } else {
          const double dx = rw.x - (double)p[0], dy = rw.y - (double)p[1], dz = rw.z - (double)p[2];
          acc[f][0] += rw.w * (dx * dx + dy * dy + dz * dz);
        }
      }
    }
    if (c == nCh - 1) {   // frame group complete: block reduction, one record per frame
      const int nred = a.fit ? 13 : 1;
#pragma unroll
      for (int f = 0; f < ONEVN_S_FT; ++f)
        for (int x = 0; x < nred; ++x) {
          const double v = warp_sum(acc[f][x]);
          if (lane == 0) red[warp][f][x] = v;
        }
      __syncthreads();
      if (tid < ONEVN_FB * ONEVN_REC) {
        const int f = tid / ONEVN_REC, x = tid % ONEVN_REC, fr = g * ONEVN_FB + f;
        if (fr < a.nFrames) {
          double v = 0.0;
          if (x < nred) {   // the 8 warps of the half that owns frame f
            const int w0 = (f / ONEVN_S_FT) * (ONEVN_THREADS / 32);
            for (int w = 0; w < ONEVN_THREADS / 32; ++w) v += red[w0 + w][f % ONEVN_S_FT][x];
          }
          else if (x >= 13) v = oS[f][x - 13];
          a.rec[(size_t)fr * ONEVN_REC + x] = v;
        }
      }
    }
    __syncthreads();   // every thread is done with this stage (and with `red`): it may be refilled
  }
}

/// One thread per frame: record -> RMSD (+ rotation, translation).  Same arithmetic as the tail of onevn_kernel.
__global__ void __launch_bounds__(128) onevn_finish_kernel(const double* rec, const int* hdr, int nFrames, const double* refsum, int fit,
                                                           double* rmsd, double* rot, double* trans) {
  if (hdr[2] == 0) return;
  const int fr = blockIdx.x * blockDim.x + threadIdx.x;
  if (fr >= nFrames) return;
  const double* v = rec + (size_t)fr * ONEVN_REC;
  const double M = refsum[3];
  if (M < 1e-14) { rmsd[fr] = -1.0; return; }   // src/Frame.cpp:1160-1163, :1300-1303
  if (!fit) { rmsd[fr] = (v[0] < 0.0) ? 0.0 : sqrt(v[0] / M); return; }
  const double cx = v[9] / M, cy = v[10] / M, cz = v[11] / M;
  double S[9];
  S[0] = v[0] - cx * refsum[0]; S[1] = v[1] - cx * refsum[1]; S[2] = v[2] - cx * refsum[2];
  S[3] = v[3] - cy * refsum[0]; S[4] = v[4] - cy * refsum[1]; S[5] = v[5] - cy * refsum[2];
  S[6] = v[6] - cz * refsum[0]; S[7] = v[7] - cz * refsum[1]; S[8] = v[8] - cz * refsum[2];
  const double gt = v[12] - M * (cx * cx + cy * cy + cz * cz);
  const double e0 = 0.5 * (gt + refsum[4]);
  const Quartic q = quartic_of(S);
  const double lam = largest_root(q, e0, S);
  const double e = e0 - lam;
  rmsd[fr] = (e < 0.0) ? 0.0 : sqrt(2.0 * e / M);
  if (rot) rotation_from_cov(S, lam, rot + 9 * (size_t)fr);
  if (trans) {
    trans[3 * (size_t)fr] = -(cx + v[13]);
    trans[3 * (size_t)fr + 1] = -(cy + v[14]);
    trans[3 * (size_t)fr + 2] = -(cz + v[15]);
  }
}

// ----------------------------------------------------------------------------
// One-vs-many, streaming variant 2 ("chunk-major"), fitted RMSD only.
// What bounded variant 1 (ncu, profiles/r1e_onevn_stream_kernel_ncu_details.csv: 0.41 of HBM, issue 47 %, nothing
// saturated): every (4 frames x 1024 atoms) step re-staged the chunk's reference atoms and atom numbers (36 KB next to
// 48 KB of frame data, six bulk copies of ~225 TMA-issue cycles each), ran a 512-thread reduction with two CTA-wide
// barriers, and only one of its two 84 KB stages was ever in flight.
// Here a CTA owns ONE 1024-atom chunk of the selection for its whole life: the chunk's reference (pre-multiplied by
// the masses) and atom numbers are loaded into shared memory once; what streams through a 6-deep mbarrier ring is
// frame data only: one stage = the chunk's span of two frames (2 x 12 KB, two bulk copies issued by a producer warp).
// Six consumer warp pairs take the stages in turn (seven ring slots: a pair that hands a slot back moves on to a slot
// that is already filled instead of waiting for the refill of its own); the two warps of a pair (one per half of the
// chunk's selected atoms) accumulate 13 FP64 sums for both frames
// relative to a local origin (the chunk's first atom of each frame, read from the stage itself), reduce them inside
// the warp and write one 16-double partial record per (frame, part); no CTA-wide barrier in the steady state.
// onevn_finish2_kernel (one thread per frame) shifts the partial sums to a common origin, adds them up and solves.
// Grid: nChunks x floor(SMs / nChunks) CTAs; CTA (c, k) takes frame pairs k, k + P, k + 2P, ...
// ----------------------------------------------------------------------------
constexpr int ONEVN2_CHUNK_BYTES = 12288;                        // per frame and stage: 1024 float atoms / 512 double atoms
constexpr int ONEVN2_MAX_APC = ONEVN2_CHUNK_BYTES / 12;          // 1024
constexpr int ONEVN2_FR = 2;                                     // frames per stage
constexpr int ONEVN2_BUF_BYTES = ONEVN2_CHUNK_BYTES + 32;        // + alignment slack at both ends
#ifndef B200_ONEVN2_STAGES
#define B200_ONEVN2_STAGES 7
#endif
#ifndef B200_ONEVN2_PAIRS
#define B200_ONEVN2_PAIRS 6   // consumer warp pairs; fewer than ring slots: a pair that frees a slot moves on to another, already filled one
#endif
#ifndef B200_ONEVN2_UNROLL
#define B200_ONEVN2_UNROLL 2   // (+3 % over 1; 7 stages and an integer-pipe float->double conversion measured no better / worse)
#endif
#ifndef B200_ONEVN2_ICVT
#define B200_ONEVN2_ICVT 0
#endif
constexpr int ONEVN2_STAGES = B200_ONEVN2_STAGES;
constexpr int ONEVN2_PAIRS = B200_ONEVN2_PAIRS;
constexpr int ONEVN2_UNROLL = B200_ONEVN2_UNROLL;
constexpr int ONEVN2_STAGE_BYTES = ONEVN2_FR * ONEVN2_BUF_BYTES;
constexpr int ONEVN2_REF_BYTES = ONEVN2_MAX_APC * 32;            // (m rx, m ry, m rz, m) of the chunk's selected atoms
constexpr int ONEVN2_IDX_BYTES = ONEVN2_MAX_APC * 4 + 32;
constexpr int ONEVN2_SMEM_BYTES = ONEVN2_REF_BYTES + ONEVN2_IDX_BYTES + ONEVN2_STAGES * ONEVN2_STAGE_BYTES + 128 +
                                  ONEVN2_PAIRS * 2 * ONEVN2_FR * 16 * 8;   // + the pair buffers
constexpr int ONEVN2_CONSUMERS = 2 * ONEVN2_PAIRS;              // warps: two per consumer pair
constexpr int ONEVN2_THREADS = 32 * (ONEVN2_CONSUMERS + 1);     // + the producer warp

struct OneVN2Args {
  const void* crd;
  size_t stride;            // elements per frame
  const int* frameIdx;      // nullable: frame n is row frameIdx[n] - srcBase of crd
  long srcBase;
  int nFrames;
  const int* atomIdx;       // nullable (identity)
  int nAtoms;
  const double* refmw;      // 4 doubles per selected atom: m rx, m ry, m rz, m
  const int* hdr;           // [0] first atom of the span, [1] chunks per frame, [2] 1 = selection sorted (usable)
  const int* kLo;           // [chunks + 1]: first selected-atom index of every chunk
  double* rec;              // 16 doubles per (frame, part): 13 sums + the local origin; part = 2 * chunk + half
};

/// float -> double on the integer pipe (exponent rebias, mantissa shift): the conversion pipe moves 16 lanes/clk/SM, a
/// third of what the streaming kernel needs at HBM speed.  Exact for zero and normal numbers; denormal floats
/// (|x| < 1.2e-38, never a coordinate) are flushed to zero.
__device__ __forceinline__ double onevn2_cvt(float v) {
#if B200_ONEVN2_ICVT
  const unsigned int u = __float_as_uint(v);
  const unsigned int ex = (u >> 23) & 0xffu;
  const unsigned int hi = (u & 0x80000000u) | (((ex + 896u) << 20) | ((u & 0x007fffffu) >> 3));
  return ex == 0u ? __hiloint2double((int)(u & 0x80000000u), 0) : __hiloint2double((int)hi, (int)(u << 29));
#else
  return (double)v;
#endif
}
__device__ __forceinline__ double onevn2_cvt(double v) { return v; }

/// Sums 13 per-lane values over the warp; lane x (and x + 16) receives the total of v[x & 15] (0 for x & 15 >= 13).
/// Halving exchange: at the step with lane mask 8 a lane keeps the half of the (zero-padded) 16 values that its bit 3
/// selects and adds the partner's copy of that half, then 8 -> 4 -> 2 -> 1 values and one last exchange across the half
/// warps: 16 shuffle+add pairs instead of the 65 of thirteen butterfly reductions.
__device__ __forceinline__ double warp_transpose_sum13(const double (&v)[13], int lane) {
  double a[8], b[4], c[2];
  const bool u8 = (lane & 8) != 0, u4 = (lane & 4) != 0, u2 = (lane & 2) != 0, u1 = (lane & 1) != 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const double hi = (i + 8 < 13) ? v[i + 8] : 0.0;
    const double keep = u8 ? hi : v[i], send = u8 ? v[i] : hi;
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double keep = u4 ? a[i + 4] : a[i], send = u4 ? a[i] : a[i + 4];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double keep = u2 ? b[i + 2] : b[i], send = u2 ? b[i] : b[i + 2];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  const double keep = u1 ? c[1] : c[0], send = u1 ? c[0] : c[1];
  double r = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  r += __shfl_xor_sync(0xffffffffu, r, 16);
  return r;
}

template <typename T>
__global__ void __launch_bounds__(ONEVN2_THREADS, 1) onevn_stream2_kernel(OneVN2Args a) {
  if (a.hdr[2] == 0) return;   // selection not sorted: onevn_kernel (launched later) does the work
  constexpr int APC = ONEVN2_CHUNK_BYTES / (3 * (int)sizeof(T));
  extern __shared__ __align__(128) unsigned char smem_o2[];
  // (m rx, m ry) and (m rz, m) of the chunk's selected atoms as two double2 arrays: a warp's 16-byte loads are then
  // 16 bytes apart (conflict-free); as one double4 array they were 32 bytes apart, a 2-way bank conflict on every load
  // (ncu, profiles/r2_onevn_stream2_kernel_ncu_raw.csv: 12.6 M conflict wavefronts of 37 M shared-load wavefronts)
  double2* srefA = reinterpret_cast<double2*>(smem_o2);
  double2* srefB = srefA + ONEVN2_MAX_APC;
  unsigned char* sidxRaw = smem_o2 + ONEVN2_REF_BYTES;
  unsigned char* ring = sidxRaw + ONEVN2_IDX_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + ONEVN2_STAGES * ONEVN2_STAGE_BYTES);
  uint64_t* empty = full + ONEVN2_STAGES;
  double* pairBuf = reinterpret_cast<double*>(empty + ONEVN2_STAGES);   // [pair][parity][frame][16]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int a0 = a.hdr[0], nCh = a.hdr[1];
  const int P = (int)gridDim.x / nCh;                    // CTAs per chunk (the host sizes the grid as nCh * P)
  if ((int)blockIdx.x >= nCh * P) return;
  const int c = (int)blockIdx.x % nCh, k = (int)blockIdx.x / nCh;
  const int nPairs = (a.nFrames + ONEVN2_FR - 1) / ONEVN2_FR;
  const int myPairs = k < nPairs ? (nPairs - 1 - k) / P + 1 : 0;
  const int kA = a.kLo[c], nSel = a.kLo[c + 1] - kA;
  const int chunkA0 = a0 + c * APC;
  const int lastAtom = a.atomIdx ? a.atomIdx[a.nAtoms - 1] : a.nAtoms - 1;
  const int nAt = min(APC, lastAtom - chunkA0 + 1);      // atoms of the span covered by this chunk
  if (tid == 0) {
    for (int s = 0; s < ONEVN2_STAGES; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 2); }
    mbar_fence_init();
  }
  // the chunk's reference atoms and atom numbers: resident for the life of the CTA
  for (int i = tid; i < nSel; i += ONEVN2_THREADS) {
    const double4 r4 = reinterpret_cast<const double4*>(a.refmw)[kA + i];
    srefA[i] = make_double2(r4.x, r4.y); srefB[i] = make_double2(r4.z, r4.w);
    if (a.atomIdx) reinterpret_cast<int*>(sidxRaw)[i] = a.atomIdx[kA + i] - chunkA0;
  }
  __syncthreads();
  if (nSel <= 0 || myPairs <= 0) return;
  const int* sidx = reinterpret_cast<const int*>(sidxRaw);
  const unsigned char* base = reinterpret_cast<const unsigned char*>(a.crd);

  if (warp == ONEVN2_CONSUMERS) {
    // ===================== producer: lanes 0 and 1 copy one frame's chunk each =====================
    for (int n = 0; n < myPairs; ++n) {
      const int st = n % ONEVN2_STAGES;
      mbar_wait(smem_u32(&empty[st]), (uint32_t)(((n / ONEVN2_STAGES) & 1) ^ 1));
      unsigned char* sbuf = ring + st * ONEVN2_STAGE_BYTES;
      uint32_t bytes = 0, len = 0;
      const unsigned char* src = nullptr;
      if (lane < ONEVN2_FR) {
        const int fr = min((k + n * P) * ONEVN2_FR + lane, a.nFrames - 1);
        const size_t row = a.frameIdx ? (size_t)((long)a.frameIdx[fr] - a.srcBase) : (size_t)fr;
        const size_t off = (row * a.stride + (size_t)3 * (size_t)chunkA0) * sizeof(T);
        const size_t s0 = off & ~(size_t)15;
        len = (uint32_t)(off - s0) + (uint32_t)nAt * 3u * (uint32_t)sizeof(T);
        src = base + s0;
        bytes = len & ~15u;
        unsigned char* dst = sbuf + lane * ONEVN2_BUF_BYTES;
        for (uint32_t t = bytes; t < len; t += 4)   // tail below a 16-byte boundary: whole 4-byte words, ordinary loads
          *reinterpret_cast<uint32_t*>(dst + t) = *reinterpret_cast<const uint32_t*>(src + t);
      }
      uint32_t tot = bytes + __shfl_xor_sync(0xffffffffu, bytes, 1);
      const uint32_t bar = smem_u32(&full[st]);
      __syncwarp();                                  // (the tail stores precede the arrive that publishes the stage)
      if (lane == 0) mbar_expect_tx(bar, tot);
      __syncwarp();
      if (lane < ONEVN2_FR && bytes) bulk_g2s(smem_u32(sbuf + lane * ONEVN2_BUF_BYTES), src, bytes, bar);
      __syncwarp();
    }
    return;
  }
  // ===================== consumers: warp w belongs to pair w / 2 and works on half w % 2 of the chunk's selected atoms =====================
  const int h = warp & 1, myPair = warp >> 1;
  const int half0 = (nSel + 1) / 2;
  const int kBeg = h ? half0 : 0, kEnd = h ? nSel : half0;
  // one record per (frame, chunk): the warp of the upper half hands its sums to the warp of the lower half through
  // pairBuf (both use the chunk's first selected atom as origin, so the sums simply add); a 64-thread named barrier per
  // frame pair, buffers alternating with the iteration parity
  const int part = c;
  const int nParts = nCh;
  int iter = 0;
  for (int n = myPair; n < myPairs; n += ONEVN2_PAIRS, ++iter) {
    const int st = n % ONEVN2_STAGES;
    mbar_wait(smem_u32(&full[st]), (uint32_t)((n / ONEVN2_STAGES) & 1));
    const unsigned char* sbuf = ring + st * ONEVN2_STAGE_BYTES;
    const int f0 = (k + n * P) * ONEVN2_FR;
    const T* fp[ONEVN2_FR];
    double o[ONEVN2_FR][3];
#pragma unroll
    for (int f = 0; f < ONEVN2_FR; ++f) {
      const int fr = min(f0 + f, a.nFrames - 1);
      const size_t row = a.frameIdx ? (size_t)((long)a.frameIdx[fr] - a.srcBase) : (size_t)fr;
      const size_t off = (row * a.stride + (size_t)3 * (size_t)chunkA0) * sizeof(T);
      fp[f] = reinterpret_cast<const T*>(sbuf + f * ONEVN2_BUF_BYTES + (off & 15));
      // local origin: the first selected atom of the chunk (keeps the sums small, costs no global read)
      const int at0 = a.atomIdx ? sidx[0] : (kA - chunkA0);
      o[f][0] = onevn2_cvt(fp[f][3 * at0]); o[f][1] = onevn2_cvt(fp[f][3 * at0 + 1]); o[f][2] = onevn2_cvt(fp[f][3 * at0 + 2]);
    }
    double acc[ONEVN2_FR][13];
#pragma unroll
    for (int f = 0; f < ONEVN2_FR; ++f)
#pragma unroll
      for (int x = 0; x < 13; ++x) acc[f][x] = 0.0;
#pragma unroll ONEVN2_UNROLL
    for (int kk = kBeg + lane; kk < kEnd; kk += 32) {
      const int at = a.atomIdx ? sidx[kk] : (kA + kk - chunkA0);
      const double2 rwa = srefA[kk], rwb = srefB[kk];
      double4 rw; rw.x = rwa.x; rw.y = rwa.y; rw.z = rwb.x; rw.w = rwb.y;
#pragma unroll
      for (int f = 0; f < ONEVN2_FR; ++f) {
        const T* p = fp[f] + 3 * at;
        const double x = onevn2_cvt(p[0]) - o[f][0], y = onevn2_cvt(p[1]) - o[f][1], z = onevn2_cvt(p[2]) - o[f][2];
        acc[f][0] = fma(x, rw.x, acc[f][0]); acc[f][1] = fma(x, rw.y, acc[f][1]); acc[f][2] = fma(x, rw.z, acc[f][2]);
        acc[f][3] = fma(y, rw.x, acc[f][3]); acc[f][4] = fma(y, rw.y, acc[f][4]); acc[f][5] = fma(y, rw.z, acc[f][5]);
        acc[f][6] = fma(z, rw.x, acc[f][6]); acc[f][7] = fma(z, rw.y, acc[f][7]); acc[f][8] = fma(z, rw.z, acc[f][8]);
        acc[f][9] = fma(x, rw.w, acc[f][9]); acc[f][10] = fma(y, rw.w, acc[f][10]); acc[f][11] = fma(z, rw.w, acc[f][11]);
        acc[f][12] = fma(rw.w, fma(x, x, fma(y, y, z * z)), acc[f][12]);
      }
    }
    // the stage's bytes have been consumed: hand it back before the reduction
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&empty[st]));
    double mine[ONEVN2_FR];   // lane x keeps sum x (x < 13), lanes 13..15 the origin
#pragma unroll
    for (int f = 0; f < ONEVN2_FR; ++f) mine[f] = warp_transpose_sum13(acc[f], lane);
    double* pb = pairBuf + ((myPair * 2 + (iter & 1)) * ONEVN2_FR) * 16;
    if (h == 1 && lane < 13) {
#pragma unroll
      for (int f = 0; f < ONEVN2_FR; ++f) pb[f * 16 + lane] = mine[f];
    }
    asm volatile("bar.sync %0, 64;" ::"r"(1 + myPair) : "memory");
    if (h == 0) {
#pragma unroll
      for (int f = 0; f < ONEVN2_FR; ++f) {
        const int fr = f0 + f;
        double v = mine[f];
        if (lane < 13) v += pb[f * 16 + lane];
        else if (lane < 16) v = o[f][lane - 13];
        if (fr < a.nFrames && lane < 16) a.rec[((size_t)fr * nParts + part) * 16 + lane] = v;
      }
    }
  }
}

/// Per-part constants of the reference for onevn_finish2_kernel: partSum[p] = (sum m rx, sum m ry, sum m rz, sum m)
/// over the selected atoms of part p = chunk p.  One block.
__device__ __forceinline__ void onevn_parts_body(const double* refmw, const int* hdr, const int* kLo, double* partSum) {
  if (hdr[2] == 0) return;
  const int nParts = hdr[1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int p = warp; p < nParts; p += nw) {
    const int kb = kLo[p], ke = kLo[p + 1];
    double s0 = 0, s1 = 0, s2 = 0, sm = 0;
    for (int k = kb + lane; k < ke; k += 32) {
      const double4 r = reinterpret_cast<const double4*>(refmw)[k];
      s0 += r.x; s1 += r.y; s2 += r.z; sm += r.w;
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); sm = warp_sum(sm);
    if (lane == 0) { partSum[4 * p] = s0; partSum[4 * p + 1] = s1; partSum[4 * p + 2] = s2; partSum[4 * p + 3] = sm; }
  }
}

__global__ void __launch_bounds__(256) onevn_parts_kernel(const double* refmw, const int* hdr, const int* kLo, double* partSum) {
  onevn_parts_body(refmw, hdr, kLo, partSum);
}

/// One thread per frame: the partial records of its parts (one per chunk), shifted to the origin of the first non-empty
/// part and added up, then the same solve as onevn_finish_kernel.  (Measured alternatives: half a warp per frame with
/// coalesced record loads and shuffles -- twice as slow: the solve's registers cap the occupancy and the part loop is a
/// chain of dependent loads; 256 frames per block in two phases -- five times as slow: too few blocks.)
__global__ void __launch_bounds__(128) onevn_finish2_kernel(const double* rec, const int* hdr, const int* kLo, const double* partSum,
                                                            int nFrames, const double* refsum, double* rmsd, double* rot, double* trans) {
  if (hdr[2] == 0) return;
  const int fr = blockIdx.x * blockDim.x + threadIdx.x;
  if (fr >= nFrames) return;
  const int nParts = hdr[1];
  const double M = refsum[3];
  if (M < 1e-14) { rmsd[fr] = -1.0; return; }   // src/Frame.cpp:1160-1163
  double v[13];
#pragma unroll
  for (int x = 0; x < 13; ++x) v[x] = 0.0;
  double ox = 0.0, oy = 0.0, oz = 0.0;
  bool have = false;
  for (int p = 0; p < nParts; ++p) {
    if (kLo[p + 1] - kLo[p] <= 0) continue;   // chunks without selected atoms wrote no record
    const double mp = partSum[4 * p + 3];
    const double2* r2 = reinterpret_cast<const double2*>(rec + ((size_t)fr * nParts + p) * 16);
    double r[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { const double2 t = r2[i]; r[2 * i] = t.x; r[2 * i + 1] = t.y; }
    if (!have) { ox = r[13]; oy = r[14]; oz = r[15]; have = true; }
    const double dx = r[13] - ox, dy = r[14] - oy, dz = r[15] - oz;   // x - o* = (x - o_p) + d
    const double sx = partSum[4 * p], sy = partSum[4 * p + 1], sz = partSum[4 * p + 2];
    v[0] += r[0] + dx * sx; v[1] += r[1] + dx * sy; v[2] += r[2] + dx * sz;
    v[3] += r[3] + dy * sx; v[4] += r[4] + dy * sy; v[5] += r[5] + dy * sz;
    v[6] += r[6] + dz * sx; v[7] += r[7] + dz * sy; v[8] += r[8] + dz * sz;
    v[9] += r[9] + mp * dx; v[10] += r[10] + mp * dy; v[11] += r[11] + mp * dz;
    v[12] += r[12] + 2.0 * (dx * r[9] + dy * r[10] + dz * r[11]) + mp * (dx * dx + dy * dy + dz * dz);
  }
  const double cx = v[9] / M, cy = v[10] / M, cz = v[11] / M;
  double S[9];
  S[0] = v[0] - cx * refsum[0]; S[1] = v[1] - cx * refsum[1]; S[2] = v[2] - cx * refsum[2];
  S[3] = v[3] - cy * refsum[0]; S[4] = v[4] - cy * refsum[1]; S[5] = v[5] - cy * refsum[2];
  S[6] = v[6] - cz * refsum[0]; S[7] = v[7] - cz * refsum[1]; S[8] = v[8] - cz * refsum[2];
  const double gt = v[12] - M * (cx * cx + cy * cy + cz * cz);
  const double e0 = 0.5 * (gt + refsum[4]);
  const Quartic q = quartic_of(S);
  const double lam = largest_root(q, e0, S);
  const double e = e0 - lam;
  rmsd[fr] = (e < 0.0) ? 0.0 : sqrt(2.0 * e / M);
  if (rot) rotation_from_cov(S, lam, rot + 9 * (size_t)fr);
  if (trans) {
    trans[3 * (size_t)fr] = -(cx + ox);
    trans[3 * (size_t)fr + 1] = -(cy + oy);
    trans[3 * (size_t)fr + 2] = -(cz + oz);
  }
}

/// Frame-to-centroid distances: dist[k * nFrames + f] (one column per centroid, as the one-vs-many passes wrote them)
/// -> nearest centroid per frame (first minimum wins, as List::AddFramesByCentroid's `dist < mindist`,
/// src/Cluster/List.cpp:183-189) and, optionally, the frame-major table distOut[f * K + k].
__global__ void __launch_bounds__(256) centroid_argmin_kernel(const double* dist, int nFrames, int K, double* distOut, int* closest,
                                                              double* closestDist) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nFrames) return;
  double best = dist[f];
  int bk = 0;
  if (distOut) distOut[(size_t)f * K] = best;
  for (int k = 1; k < K; ++k) {
    const double d = dist[(size_t)k * nFrames + f];
    if (distOut) distOut[(size_t)f * K + k] = d;
    if (d < best) { best = d; bk = k; }
  }
  if (closest) closest[f] = bk;
  if (closestDist) closestDist[f] = best;
}

/// Same for the frames x centroids table the pair engine wrote (float, frame-major): dist[f * K + k].
__global__ void __launch_bounds__(256) centroid_argmin_rows_kernel(const float* dist, int nFrames, int K, double* distOut, int* closest,
                                                                   double* closestDist) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nFrames) return;
  const float* row = dist + (size_t)f * K;
  float best = row[0];
  int bk = 0;
  if (distOut) distOut[(size_t)f * K] = (double)best;
  for (int k = 1; k < K; ++k) {
    const float d = row[k];
    if (distOut) distOut[(size_t)f * K + k] = (double)d;
    if (d < best) { best = d; bk = k; }
  }
  if (closest) closest[f] = bk;
  if (closestDist) closestDist[f] = (double)best;
}

/// refw[k] = (rx, ry, rz, m); refsum = (sum m r, M, sum m|r|^2).  One block.
__device__ __forceinline__ void onevn_setup_body(const double* ref, const double* mass, int n, double* refw, double* refsum,
                                                 double* refmw) {
  __shared__ double part[32][5];
  double s0 = 0, s1 = 0, s2 = 0, sm = 0, sg = 0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const double m = mass ? mass[k] : 1.0;
    const double x = ref[3 * k], y = ref[3 * k + 1], z = ref[3 * k + 2];
    refw[4 * k] = x; refw[4 * k + 1] = y; refw[4 * k + 2] = z; refw[4 * k + 3] = m;
    if (refmw) { refmw[4 * k] = m * x; refmw[4 * k + 1] = m * y; refmw[4 * k + 2] = m * z; refmw[4 * k + 3] = m; }   // (streaming variant 2)
    s0 += m * x; s1 += m * y; s2 += m * z; sm += m; sg += m * (x * x + y * y + z * z);
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); sm = warp_sum(sm); sg = warp_sum(sg);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if ((threadIdx.x & 31) == 0) { part[w][0] = s0; part[w][1] = s1; part[w][2] = s2; part[w][3] = sm; part[w][4] = sg; }
  __syncthreads();
  if (threadIdx.x < 5) {
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += part[i][threadIdx.x];   // fixed order: deterministic
    refsum[threadIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) onevn_setup_kernel(const double* ref, const double* mass, int n, double* refw, double* refsum,
                                                          double* refmw = nullptr) {
  onevn_setup_body(ref, mass, n, refw, refsum, refmw);
}
/// Everything a one-vs-many pass needs before the streaming kernel, in ONE launch of one 1024-thread block (three
/// single-block launches of 12-16 us each before): the reference set-up (skipped when ref == nullptr: a streaming handle
/// did it when the reference was loaded), the chunk table, and -- variant 2, partSum != nullptr -- the per-part sums.
__global__ void __launch_bounds__(1024) onevn_prep_kernel(const double* ref, const double* mass, int n, double* refw, double* refsum,
                                                          double* refmw, const int* atomIdx, int atomsPerChunk, int maxChunks,
                                                          int* hdr, int* kLo, double* partSum) {
  if (ref) onevn_setup_body(ref, mass, n, refw, refsum, refmw);
  onevn_chunks_body(atomIdx, n, atomsPerChunk, maxChunks, hdr, kLo);
  __syncthreads();   // (refmw, hdr, kLo were written by this block: visible to it after the barrier)
  if (partSum) onevn_parts_body(refmw, hdr, kLo, partSum);
}

// ----------------------------------------------------------------------------
// centroid_build_kernel: Metric_RMS::CalculateCentroid (src/Cluster/Metric_RMS.cpp:86-113) for many clusters at once.
// The reference fits every frame of a cluster, in order, to the running SUM of the frames before it and adds the
// rotated frame -- a scan: frame j cannot start before frame j-1 has been added.  One CTA per cluster walks its frame
// list; within a frame the atoms are spread over the threads: pass A accumulates the covariance with the running sum
// (17 FP64 sums, block reduction), one thread solves for the rotation (same quartic / adjugate code as the one-vs-many
// path), pass B rotates the centred frame and adds it.  The running sum lives in the output array (L2-resident).
// ----------------------------------------------------------------------------
constexpr int CENT_THREADS = 512;
struct CentroidArgs {
  const float* crd; size_t stride; long srcBase;   // frame f is row f - srcBase
  const int* frames;      // concatenated frame lists
  const int* offsets;     // [K + 1]
  const int* atomIdx; int nAtoms;
  const double* mass;     // nullable
  int fit;
  double* out;            // K x 3 nAtoms: the centroid (SMEM == false: also the running sum)
};
/// SMEM: the running sum of the cluster (3 nAtoms doubles) lives in dynamic shared memory -- every frame reads and rewrites it,
/// and a frame cannot start before the previous one has been added: its latency is the kernel's time (a 50,000-frame cluster of
/// 2,000 atoms: 39 -> ~10 us per frame) -- else (more than ~9,000 atoms) in the output array (L2).
template <bool SMEM>
__global__ void __launch_bounds__(CENT_THREADS) centroid_build_kernel(CentroidArgs a) {
  extern __shared__ __align__(16) unsigned char smem_cent[];
  __shared__ double red[CENT_THREADS / 32][17];
  __shared__ double tot[17];
  __shared__ double bc[16];   // broadcast: U (9), centre (3), [12] total mass
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k = blockIdx.x;
  const int j0 = a.offsets[k], m = a.offsets[k + 1] - j0;
  if (m <= 0) return;
  double* outc = a.out + (size_t)k * 3 * a.nAtoms;
  double* cent = SMEM ? reinterpret_cast<double*>(smem_cent) : outc;
  // v[0..n) summed over the block: the totals land in tot[] (valid after the barrier at the end)
  auto block_sum = [&](const double* v, int n) {
    for (int x = 0; x < n; ++x) { const double s = warp_sum(v[x]); if (lane == 0) red[warp][x] = s; }
    __syncthreads();
    if (warp == 0 && lane < n) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < CENT_THREADS / 32; ++w) s += red[w][lane];
      tot[lane] = s;
    }
  };
  // next frame of the list: its lines are asked for while this one is being solved (the frames of a cluster are scattered)
  auto prefetch_frame = [&](int j) {
    if (j >= m) return;
    const float* nxt = a.crd + (size_t)((long)a.frames[j0 + j] - a.srcBase) * a.stride;
    const size_t at0 = (size_t)(a.atomIdx ? a.atomIdx[0] : 0), at1 = (size_t)(a.atomIdx ? a.atomIdx[a.nAtoms - 1] : a.nAtoms - 1);
    const char* lo = reinterpret_cast<const char*>(nxt + 3 * at0);
    const char* hi = reinterpret_cast<const char*>(nxt + 3 * at1 + 3);
    for (const char* q = lo + (size_t)tid * 128; q < hi; q += (size_t)CENT_THREADS * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
  };
  // ---- first frame: the start of the sum, centred when fitting (:94-97)
  {
    const float* src = a.crd + (size_t)((long)a.frames[j0] - a.srcBase) * a.stride;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = tid; i < a.nAtoms; i += CENT_THREADS) {
      const size_t at = (size_t)(a.atomIdx ? a.atomIdx[i] : i);
      const double w = a.mass ? a.mass[i] : 1.0;
      const double x = (double)src[3 * at], y = (double)src[3 * at + 1], z = (double)src[3 * at + 2];
      cent[3 * i] = x; cent[3 * i + 1] = y; cent[3 * i + 2] = z;
      v[0] += w * x; v[1] += w * y; v[2] += w * z; v[3] += w;
    }
    block_sum(v, 4);
    prefetch_frame(1);
    __syncthreads();
    const double M0 = tot[3];
    if (tid == 0) bc[12] = M0;
    if (a.fit && M0 != 0.0) {
      const double cx = tot[0] / M0, cy = tot[1] / M0, cz = tot[2] / M0;
      for (int i = tid; i < a.nAtoms; i += CENT_THREADS) { cent[3 * i] -= cx; cent[3 * i + 1] -= cy; cent[3 * i + 2] -= cz; }
    }
    __syncthreads();
  }
  const double M = bc[12];
  for (int j = 1; j < m; ++j) {
    const float* src = a.crd + (size_t)((long)a.frames[j0 + j] - a.srcBase) * a.stride;
    if (a.fit) {
      const size_t at0 = (size_t)(a.atomIdx ? a.atomIdx[0] : 0);
      const double ox = (double)src[3 * at0], oy = (double)src[3 * at0 + 1], oz = (double)src[3 * at0 + 2];
      // pass A: v[0..8] = sum m x' r^T, [9..11] = sum m x', [12] = sum m |x'|^2, [13..15] = sum m r, [16] = sum m |r|^2
      double v[17];
#pragma unroll
      for (int x = 0; x < 17; ++x) v[x] = 0.0;
      for (int i = tid; i < a.nAtoms; i += CENT_THREADS) {
        const size_t at = (size_t)(a.atomIdx ? a.atomIdx[i] : i);
        const double w = a.mass ? a.mass[i] : 1.0;
        const double x = (double)src[3 * at] - ox, y = (double)src[3 * at + 1] - oy, z = (double)src[3 * at + 2] - oz;
        const double rx = cent[3 * i], ry = cent[3 * i + 1], rz = cent[3 * i + 2];
        const double mx = w * x, my = w * y, mz = w * z;
        v[0] += mx * rx; v[1] += mx * ry; v[2] += mx * rz;
        v[3] += my * rx; v[4] += my * ry; v[5] += my * rz;
        v[6] += mz * rx; v[7] += mz * ry; v[8] += mz * rz;
        v[9] += mx; v[10] += my; v[11] += mz;
        v[12] += mx * x + my * y + mz * z;
        v[13] += w * rx; v[14] += w * ry; v[15] += w * rz;
        v[16] += w * (rx * rx + ry * ry + rz * rz);
      }
      block_sum(v, 17);
      prefetch_frame(j + 1);
      if (warp == 0) {
        __syncwarp();
        if (lane == 0) {
          const double cx = tot[9] / M, cy = tot[10] / M, cz = tot[11] / M;   // centre of the frame relative to o
          double S[9];
          S[0] = tot[0] - cx * tot[13]; S[1] = tot[1] - cx * tot[14]; S[2] = tot[2] - cx * tot[15];
          S[3] = tot[3] - cy * tot[13]; S[4] = tot[4] - cy * tot[14]; S[5] = tot[5] - cy * tot[15];
          S[6] = tot[6] - cz * tot[13]; S[7] = tot[7] - cz * tot[14]; S[8] = tot[8] - cz * tot[15];
          const double gt = tot[12] - M * (cx * cx + cy * cy + cz * cz);
          const double e0 = 0.5 * (gt + tot[16]);
          const double lam = largest_root(quartic_of(S), e0, S);
          rotation_from_cov(S, lam, bc);
          bc[9] = cx + ox; bc[10] = cy + oy; bc[11] = cz + oz;
        }
      }
      __syncthreads();
    } else {
      prefetch_frame(j + 1);
    }
    // pass B: rotate the centred frame (Frame::Rotate, src/Frame.h:508-517) and add it (:102-104)
    for (int i = tid; i < a.nAtoms; i += CENT_THREADS) {
      const size_t at = (size_t)(a.atomIdx ? a.atomIdx[i] : i);
      double x = (double)src[3 * at], y = (double)src[3 * at + 1], z = (double)src[3 * at + 2];
      if (a.fit) {
        x -= bc[9]; y -= bc[10]; z -= bc[11];
        const double X = x * bc[0] + y * bc[1] + z * bc[2], Y = x * bc[3] + y * bc[4] + z * bc[5],
                     Z = x * bc[6] + y * bc[7] + z * bc[8];
        x = X; y = Y; z = Z;
      }
      cent[3 * i] += x; cent[3 * i + 1] += y; cent[3 * i + 2] += z;
    }
    __syncthreads();
  }
  for (int i = tid; i < 3 * a.nAtoms; i += CENT_THREADS) outc[i] = cent[i] / (double)m;   // (:108)
}

// ----------------------------------------------------------------------------
// FP64 MMA issue-peak probe (register-only): roofline denominator for pair_kernel.
// ----------------------------------------------------------------------------
template <int VAR>
__global__ void __launch_bounds__(128) fp64_mma_peak_kernel(double* sink, int iters) {
  double c[12][4];
#pragma unroll
  for (int x = 0; x < 12; ++x)
#pragma unroll
    for (int e = 0; e < 4; ++e) c[x][e] = 0.0;
  double a[8], b[4];
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = 1.0 + threadIdx.x * 1e-9 + e;
#pragma unroll
  for (int e = 0; e < 4; ++e) b[e] = 0.5 + e;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int x = 0; x < 12; ++x) {
      if constexpr (VAR == 0) { mma_884(c[x][0], c[x][1], a[0], b[0]); mma_884(c[x][2], c[x][3], a[1], b[0]); }
      else if constexpr (VAR == 1) mma_1684(c[x], a[0], a[1], b[0]);
      else if constexpr (VAR == 2) mma_1688(c[x], a, b);
      else if constexpr (VAR == 3) mma_16816(c[x], a, b);
      else {  // VAR 4: plain DFMA, 12 x 4 independent chains
#pragma unroll
        for (int e = 0; e < 4; ++e) c[x][e] = fma(a[e], b[e], c[x][e]);
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int x = 0; x < 12; ++x)
#pragma unroll
    for (int e = 0; e < 4; ++e) s += c[x][e];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace b200
