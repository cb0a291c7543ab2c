// pair_i8.cuh -- tcgen05 (5th-gen tensor core) path of the all-pairs best-fit RMSD.
//
// Idea ("exact integer covariance"): best-fit RMSD is 1-Lipschitz in the RMS coordinate
// perturbation, so the centred, sqrt(mass)-scaled coordinates are rounded ONCE per frame to a
// fixed-point grid of spacing 2^-qs (24-bit signed integers; qs chosen from the largest centred
// coordinate so the worst-case RMSD change stays < 8.5e-5 A) and from then on everything is exact:
//   * every integer is split into three balanced signed base-256 digits (int8),
//   * the nine covariance entries of a frame pair are 81 int8 x int8 dot products over the atoms,
//     computed by tcgen05.mma kind::i8 with int32 accumulators in TMEM (exact for < 131072 atoms),
//   * the epilogue recombines the digits in exact int64 arithmetic (S is exact below 2^63, converted to
//     FP64 once; G carries < 2^-52 relative error), then solves the key-matrix quartic per pair.
// The cancellation E0 - lambda that defeats fp32/TF32 accumulation never
// sees a rounding error larger than FP64's.  Replaces, like pair_kernel, the covariance loop of
// Frame::RMSD_CenteredRef (src/Frame.cpp:1184-1208) + the eigen-solve (src/Frame.cpp:1215-1268).
//
// Operand layout in HBM = ready-made shared-memory images for UMMA (no swizzle, K-major):
//   "row group" g = 14 frames = 126 operand rows (+2 zero rows): row = 9*(f%14) + 3*plane + digit
//   block (g, c) = 128 rows x 64 atoms (bytes) = 8 KB, stored as
//        [row/8 (16)] [k16 = (k%64)/16 (4)] [row%8 (8)] [k%16 (16 B)]
//   i.e. 8x16-byte core matrices of 128 contiguous bytes, LBO = 128 B between the core matrices of
//   one K step, SBO = 512 B between 8-row groups.  Blocks are ordered [g][c], so an A operand
//   stage (one row group, 64 atoms) is one contiguous 8 KB bulk copy and a B operand stage (two
//   consecutive row groups = 28 frames = N 256) is two of them.
// One MMA tile = 14 x 28 frame pairs per CTA = D[128 x 256] int32 = 256 TMEM columns (a CTA pair computes
// 28 x 28 with tcgen05.mma.cta_group::2); two accumulator buffers fill the 512 columns, so the epilogue of
// tile n overlaps the MMAs of tile n+1.
#pragma once
#include "rmsd_kernels.cuh"

namespace b200 {

constexpr int I8_FR_PER_RG = 14;             // frames per 128-row operand group
constexpr int I8_ROWS_PER_FR = 9;            // 3 planes x 3 digits
constexpr int I8_KC = 64;                    // atoms (bytes per row) per pipeline stage
constexpr int I8_BLK_BYTES = 128 * I8_KC;    // 8192
/// Pipeline geometry per MMA CTA group.  CG = 1: one CTA computes a 14 x 28 tile (M128 N256), a stage
/// holds 64 atoms of A (1 block) and B (2 blocks).  CG = 2: a CTA pair computes a 28 x 28 tile
/// (tcgen05.mma.cta_group::2, M256 N256); each CTA stages its own A row group and ONE of the two B row
/// groups, so shared-memory fill, operand reads and L2 -> SM traffic per SM drop by a third and a stage holds
/// 128 atoms (4 MMAs) in the same footprint.
#ifndef B200_I8_TPW
#define B200_I8_TPW 1           // tiles per FP64 window (1 or 2); 2 keeps a tile's covariance in shared memory: fewer ring slots
#endif
template <int CG> struct I8Geom;
template <> struct I8Geom<1> {
  static constexpr int BPS = 1;      // 64-atom blocks per stage
  static constexpr int STAGES = B200_I8_TPW == 2 ? 3 : 5;
  static constexpr int BBLK = 2;     // B blocks per 64 atoms held by this CTA
};
#ifndef B200_I8_CG2_BPS
#define B200_I8_CG2_BPS 2
#define B200_I8_CG2_STAGES 4
#endif
template <> struct I8Geom<2> {
  static constexpr int BPS = B200_I8_CG2_BPS;
  static constexpr int STAGES = B200_I8_CG2_STAGES;
  static constexpr int BBLK = 1;
};
template <int CG> __host__ __device__ constexpr int i8_stage_bytes() {
  return I8Geom<CG>::BPS * (1 + I8Geom<CG>::BBLK) * I8_BLK_BYTES;
}
constexpr int I8_TILE_I = I8_FR_PER_RG;      // 14
constexpr int I8_TILE_J = 2 * I8_FR_PER_RG;  // 28
constexpr long long I8_QMAX = 8355711;       // 127*(1+256+65536): largest |q| with balanced digits

__host__ __device__ inline size_t i8_image_bytes(int nRowGroups, int nC) {
  return (size_t)nRowGroups * (size_t)nC * I8_BLK_BYTES;
}

// ----------------------------------------------------------------------------
// Packing: two passes over the raw COORDS (the second one hits L2).
// ----------------------------------------------------------------------------
struct I8StatsArgs {
  const void* crd;          // float (COORDS) or double (centroid frames) rows, see the kernel's template argument
  size_t stride; const int* frameIdx; long srcBase;
  int nFrames; int f0;      // frames [f0, nFrames) of the set are processed by this launch
  const int* atomIdx; int nAtoms;
  const double* centerMass; const double* covMass;
  double* centers;          // 3 per frame
  unsigned int* maxAbsBits; // float bits of max |(x-c)*sqrt(m)|, atomicMax
};

/// One warp per frame: centre (src/Frame.cpp:1043-1055 / :1141-1166) and the extent.
template <typename T>
__global__ void __launch_bounds__(256) i8_stats_kernel(I8StatsArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = a.f0 + blockIdx.x * 8 + warp;
  if (f >= a.nFrames) return;
  const long row = (a.frameIdx ? (long)a.frameIdx[f] : (long)f) - a.srcBase;
  const T* src = reinterpret_cast<const T*>(a.crd) + (size_t)row * a.stride;
  double sx = 0.0, sy = 0.0, sz = 0.0, sm = 0.0;
  for (int k = lane; k < a.nAtoms; k += 32) {
    const int at = a.atomIdx ? a.atomIdx[k] : k;
    const double m = a.centerMass ? a.centerMass[k] : 1.0;
    sx += (double)src[3 * (size_t)at] * m; sy += (double)src[3 * (size_t)at + 1] * m;
    sz += (double)src[3 * (size_t)at + 2] * m; sm += m;
  }
  sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz); sm = warp_sum(sm);
  double cx = 0.0, cy = 0.0, cz = 0.0;
  if (sm != 0.0) { cx = sx / sm; cy = sy / sm; cz = sz / sm; }
  float mx = 0.f;
  for (int k = lane; k < a.nAtoms; k += 32) {
    const int at = a.atomIdx ? a.atomIdx[k] : k;
    const double w = a.covMass ? sqrt(a.covMass[k]) : 1.0;
    const double x = ((double)src[3 * (size_t)at] - cx) * w, y = ((double)src[3 * (size_t)at + 1] - cy) * w,
                 z = ((double)src[3 * (size_t)at + 2] - cz) * w;
    mx = fmaxf(mx, (float)fmax(fabs(x), fmax(fabs(y), fabs(z))));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) {
    a.centers[3 * (size_t)f] = cx; a.centers[3 * (size_t)f + 1] = cy; a.centers[3 * (size_t)f + 2] = cz;
    atomicMax(a.maxAbsBits, __float_as_uint(mx * 1.000001f));
  }
}

struct I8QuantArgs {
  const void* crd; size_t stride; const int* frameIdx; long srcBase;
  int nFrames; int f0;
  const int* atomIdx; int nAtoms; int nC;
  const double* covMass; const double* centers;
  double scale;     // 2^qs
  double invScale2; // 2^-2qs
  uint8_t* image;   // zero-initialised
  double* G;        // sum q^2 * 2^-2qs  (= sum m|x-c|^2 of the rounded coordinates)
};

__device__ __forceinline__ void i8_digits(long long q, int& d0, int& d1, int& d2) {
  d0 = (int)(signed char)(q & 0xff);
  const long long q1 = (q - d0) >> 8;
  d1 = (int)(signed char)(q1 & 0xff);
  d2 = (int)((q1 - d1) >> 8);
}

/// One warp per frame; a lane handles 4 consecutive atoms per step and writes nine 32-bit words
/// (plane x digit), each the 4 K-adjacent bytes of one operand row.
template <typename T>
__global__ void __launch_bounds__(256) i8_quant_kernel(I8QuantArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = a.f0 + blockIdx.x * 8 + warp;
  if (f >= a.nFrames) return;
  const long row = (a.frameIdx ? (long)a.frameIdx[f] : (long)f) - a.srcBase;
  const T* src = reinterpret_cast<const T*>(a.crd) + (size_t)row * a.stride;
  const double cx = a.centers[3 * (size_t)f], cy = a.centers[3 * (size_t)f + 1], cz = a.centers[3 * (size_t)f + 2];
  const int g = f / I8_FR_PER_RG, r0 = I8_ROWS_PER_FR * (f % I8_FR_PER_RG);
  uint8_t* gbase = a.image + (size_t)g * a.nC * I8_BLK_BYTES;
  long long gsum = 0;
  for (int k0 = 4 * lane; k0 < a.nAtoms; k0 += 128) {
    uint32_t word[9];
#pragma unroll
    for (int x = 0; x < 9; ++x) word[x] = 0u;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const int k = k0 + kk;
      if (k < a.nAtoms) {
        const int at = a.atomIdx ? a.atomIdx[k] : k;
        const double w = (a.covMass ? sqrt(a.covMass[k]) : 1.0) * a.scale;
        const double v[3] = {((double)src[3 * (size_t)at] - cx) * w, ((double)src[3 * (size_t)at + 1] - cy) * w,
                             ((double)src[3 * (size_t)at + 2] - cz) * w};
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          long long q = __double2ll_rn(v[p]);
          q = q > I8_QMAX ? I8_QMAX : (q < -I8_QMAX ? -I8_QMAX : q);
          gsum += q * q;
          int d0, d1, d2;
          i8_digits(q, d0, d1, d2);
          word[3 * p + 0] |= (uint32_t)(d0 & 0xff) << (8 * kk);
          word[3 * p + 1] |= (uint32_t)(d1 & 0xff) << (8 * kk);
          word[3 * p + 2] |= (uint32_t)(d2 & 0xff) << (8 * kk);
        }
      }
    }
    const int c = k0 / I8_KC, kb = k0 % I8_KC;
    uint8_t* blk = gbase + (size_t)c * I8_BLK_BYTES + (kb >> 4) * 128 + (kb & 15);
#pragma unroll
    for (int x = 0; x < 9; ++x) {
      const int r = r0 + x;
      *reinterpret_cast<uint32_t*>(blk + (r >> 3) * 512 + (r & 7) * 16) = word[x];
    }
  }
  // 3 N q^2 <= 3 N 2^46: the exact int64 sum over the warp holds up to 43,690 atoms; beyond, the lanes' partial sums
  // (exact) are added in FP64 (relative error 2^-53, the precision G is used at anyway)
  if (a.nAtoms <= 40000) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
    if (lane == 0) a.G[f] = (double)gsum * a.invScale2;
  } else {
    const double gd = warp_sum((double)gsum);
    if (lane == 0) a.G[f] = gd * a.invScale2;
  }
}

// ----------------------------------------------------------------------------
// tcgen05 / TMEM helpers
// ----------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t dstSmem, uint32_t ncols) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dstSmem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dstSmem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
/// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, issued by one thread.  CG == 2: issued by the
/// leader CTA of a pair; M = 256 (128 rows from each CTA's A), each CTA supplies half of B's N columns.
template <int CG>
__device__ __forceinline__ void umma_i8(uint32_t dTmem, uint64_t aDesc, uint64_t bDesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (CG == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dTmem), "l"(aDesc), "l"(bDesc), "r"(idesc),
        "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dTmem), "l"(aDesc), "l"(bDesc), "r"(idesc),
        "r"(accumulate)
        : "memory");
}
/// mbarrier arrive once all previously issued MMAs of this thread have completed.  CG == 2: the arrive is
/// multicast to the barrier at the same shared-memory offset in both CTAs of the pair.
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((unsigned short)3)
                 : "memory");
}
/// One lane of the (converged) warp: the single-thread roles run their loops warp-uniformly, so that addresses,
/// descriptors and counters stay in uniform registers, and only issue under this predicate.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(pred));
  return pred != 0;
}
// ---- thread-block-cluster helpers (CTA pair)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
/// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t cluster_map(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
/// Arrive on an mbarrier of another CTA of the cluster.  Default semantics (release at CTA scope), as CUTLASS's
/// ClusterBarrier::arrive does: the explicit .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR in front of
/// every arrive (and .acquire.cluster waits to a CCTL.IVALL after every wait), which made the operand ring of the
/// CTA-pair kernel latency-bound.  What the barriers order here is async-proxy traffic (TMA writes, tensor-core
/// reads, TMEM) which is fenced by complete_tx / tcgen05.commit / tcgen05.fence, not by generic-proxy scopes.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t clusterAddr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(clusterAddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
/// K-major, no-swizzle UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor, version 1).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smemAddr, uint32_t lboBytes, uint32_t sboBytes) {
  return (uint64_t)((smemAddr >> 4) & 0x3fff) | ((uint64_t)((lboBytes >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sboBytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}
/// kind::i8 instruction descriptor: D = S32, A = B = signed int8, both K-major, M x N.
__host__ __device__ constexpr uint32_t umma_idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, int* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------
// pair_i8_kernel
//
// Persistent, warp-specialised, one CTA per SM.  Roles (20 warps; warp w runs on SM sub-partition w % 4):
//   warps 0..3    drain: one per TMEM sub-partition.  Per exchange group (7 column frames x 9 accumulator
//                 columns, fetched with three tcgen05.ld) fold the three B digits of every (row, column
//                 frame, plane) in-thread -- exact int64, integer pipe only -- and hand 3 int64 per operand
//                 row and column frame to the exchange buffer
//   warps 4..17   solve: one thread per frame pair (2 row frames x 14 column frames per warp): read the 27
//                 partial sums of the pair, fold the three A digits (exact int64), convert once to FP64,
//                 quartic coefficients in FP64 inside the FP64 window, FP32 Newton approach to the root under
//                 the next MMAs, FP64 polish in the next window, float store straight into cpptraj's
//                 Matrix<float> layout
//   warp 18       TMA producer: operand stages through a full/empty mbarrier ring (whole warp loops so that
//                 addresses stay in uniform registers; one elected lane issues)
//   warp 19       MMA issuer: tcgen05.mma kind::i8 K32 into one of two TMEM accumulators, same elect scheme
//                 (CG 2: the leader CTA issues for the pair; the peer's warp 19 relays "landed")
// FP64 window: on this part FP64 instructions share the tensor pipe with tcgen05.mma and run at ~10 % of their
// rate while MMAs are in flight (profiles/r1c_microbench_fp64_under_mma.txt).  The solve warps therefore do
// all their FP64 work in a window that opens when the MMAs of tile n+1 have completed (accFull) and closes
// when every solve warp of the CTA (pair) has arrived on fpDone; the MMA warp issues tile n+2 only then.
// The exchange buffer is a ring of four column groups (7 column frames each) with its own full/empty
// mbarriers, so the drain of tile n+1 runs under the solve of tile n, both under the MMAs of tile n+2: no
// CTA-wide barrier anywhere in the steady state.
// Tiles are enumerated column tile by column tile (valid row tiles only); the list is cut into
// chunks of `chunkLen` consecutive tiles, chunk c going to CTA group c % groups: every group gets
// the same number of tiles (+-1), groups running at the same time work on neighbouring columns
// (their A rows stay in L2).
// ----------------------------------------------------------------------------
constexpr int I8_TPW = B200_I8_TPW;
static_assert(I8_TPW == 1 || I8_TPW == 2, "tiles per FP64 window");
#ifndef B200_I8_DRAIN_WARPS
#define B200_I8_DRAIN_WARPS 4   // measured: 4 (640 threads, 96 registers, no spills) beats 8 (768 threads, 80 registers)
#endif
constexpr int I8_DRAIN_WARPS = B200_I8_DRAIN_WARPS;   // 4 or 8: one or two per TMEM sub-partition (8: one per column half of the tile)
static_assert(I8_DRAIN_WARPS == 4 || I8_DRAIN_WARPS == 8, "drain warps");
constexpr int I8_GROUPS_PER_DRAIN_WARP = 16 / I8_DRAIN_WARPS;
constexpr int I8_SOLVE_WARPS = 14;
constexpr int I8_WARP_PRODUCER = I8_DRAIN_WARPS + I8_SOLVE_WARPS;   // 18
constexpr int I8_WARP_MMA = I8_WARP_PRODUCER + 1;                   // 19
constexpr int I8_THREADS = 32 * (I8_WARP_MMA + 1);                  // 640
constexpr int I8_XJ_DBL = 128 * 3 + 11;      // entries per column frame: 3 per operand row + pad; odd => conflict-free LDS / STS
constexpr int I8_XBUF_BYTES = I8_TILE_J * I8_XJ_DBL * 8;  // 88480
constexpr int I8_MAX_STAGES = 6;
constexpr int I8_NBARS = 2 * I8_MAX_STAGES + 4 + 8 + 2 + 2;
#ifndef B200_I8_EARLY_WINDOW
#define B200_I8_EARLY_WINDOW 0   // 1: the FP64 window opens when the last MMA of the next tile has been ISSUED (not completed)
#endif
// I8_TPW == 2: the covariance (9 doubles) and 2 E0 of the tile that waits for the next FP64 window, one slot per solve
// thread, [entry][thread] (conflict-free); in registers it did not fit the 96-register budget (spills, slower)
constexpr int I8_HELD_BYTES = I8_TPW == 2 ? 10 * I8_SOLVE_WARPS * 32 * 8 : 0;
template <int CG> __host__ __device__ constexpr int i8_smem_bytes() {
  return I8Geom<CG>::STAGES * i8_stage_bytes<CG>() + I8_XBUF_BYTES + 32 + 256 + I8_HELD_BYTES;
}
static_assert(I8_NBARS * 8 + 8 <= 256, "barrier block");
static_assert(I8_XBUF_BYTES % 32 == 0, "barrier alignment");
static_assert(i8_smem_bytes<1>() <= 232448 && i8_smem_bytes<2>() <= 232448, "shared memory budget");

struct PairI8Args {
  const uint8_t* PA;   // operand images of the row frames (i)
  const uint8_t* PB;   // operand images of the column frames (j); even number of row groups allocated
  const double* GA;
  const double* GB;
  int nC;              // 64-atom chunks
  int nRows, nCols;    // valid i / j frames
  int rowLo, rowHi;    // rows written by this launch: [rowLo, rowHi)
  int it0, it1;        // i super-tiles (CG row groups = 14*CG frames) of this launch: [it0, it1)
  int jt0, jt1;        // j tiles (28 frames) of this launch: [jt0, jt1)
  int chunkLen;        // consecutive tiles of the list per chunk (>= 1)
  const double* totalMass;
  double invScale2;    // 2^-2qs: integer covariance -> A^2
  float* out;          // TRI: out[triIndex - outBase]; FULL: out[i*ldo + j]
  size_t outBase;
  size_t ldo;
  // ---- DBG instantiation only
  double* dbgS;        // nullable: 9 doubles per (i,j) at (i*nCols + j)*9, integer units
  long long* dbgClk;   // nullable: per-CTA cycle counters [16] (timing experiments)
  int dbgMode;         // 0 normal; timing experiments: 1 no per-pair solve, 2 drain only frees TMEM,
                       // 3 = 2 + no operand loads, 4 = 2 + every CTA streams the same A rows, 5 = 2 + no MMAs,
                       // 6 = solve warps only recycle the exchange buffer,
                       // 7 = no FP64 window, the solve warps only export the integer covariance to dbgRing (what a tensor SM
                       //     would still do if the per-pair solve ran on other SMs)
  long long* dbgRing;  // mode 7: 4 slots x 10 entries x 448 threads per CTA
};

/// Row super-tiles [it0, hi) of column tile jt hold at least one wanted pair.
template <bool TRI, int CG>
__host__ __device__ __forceinline__ int i8_col_tiles(int it0, int it1, int jt) {
  // TRI: a tile is wanted iff its largest j exceeds its smallest i: 28 jt + 27 > 14 CG it
  //      CG = 1: it <= 2 jt + 1;  CG = 2: it <= jt
  const int lim = CG == 1 ? 2 * jt + 2 : jt + 1;
  const int hi = TRI ? (it1 < lim ? it1 : lim) : it1;
  return hi > it0 ? hi - it0 : 0;
}
template <bool TRI, int CG>
__host__ __device__ inline long i8_count_tiles(int it0, int it1, int jt0, int jt1) {
  long n = 0;
  for (int jt = jt0; jt < jt1; ++jt) n += i8_col_tiles<TRI, CG>(it0, it1, jt);
  return n;
}
/// Walks this CTA group's chunks of the tile list; every role keeps its own copy.
template <bool TRI, int CG>
struct I8TileIter {
  int jt, base;         // column being walked and the list position of its first tile
  int pos, chunkEnd;    // next list position of this group, end of its current chunk
  int skip;             // list positions between the end of one chunk and the start of the group's next one
  __device__ __forceinline__ void init(const PairI8Args& a) {
    jt = a.jt0; base = 0;
    pos = ((int)blockIdx.x / CG) * a.chunkLen; chunkEnd = pos + a.chunkLen;
    skip = ((int)gridDim.x / CG - 1) * a.chunkLen;
  }
  __device__ __forceinline__ bool next(const PairI8Args& a, int& it, int& jtOut) {
    if (pos == chunkEnd) { pos += skip; chunkEnd = pos + a.chunkLen; }
    while (jt < a.jt1) {
      const int c = i8_col_tiles<TRI, CG>(a.it0, a.it1, jt);
      if (pos < base + c) { it = a.it0 + (pos - base); jtOut = jt; ++pos; return true; }
      base += c; ++jt;
    }
    return false;
  }
};

// Per-pair solve of the tcgen05 epilogue.  On this part FP64 instructions share the tensor pipe with
// tcgen05.mma and crawl (~10 % of their rate) while MMAs are in flight (tools/microbench/mma_fp64_mix.cu),
// so the solve is cut in two: i8_coeffs() -- everything that must be FP64, 47 instructions -- runs in a short
// window between two tiles' MMAs (the MMA warp waits for it), i8_root() -- FP32 Newton, integer exponent
// scaling and 8 FP64 instructions for the final correction -- runs under the next tile's MMAs.
//
// Key-matrix quartic P(l) = l^4 + c2 l^2 + c1 l + c0 (src/qcprot.cpp:167-246 states the same identity); the
// wanted root is lambda_max <= E0.  With l = E0 - y:
//     Q(y) = y^4 - 4 E0 y^3 + (6 E0^2 + c2) y^2 - (4 E0^3 + 2 c2 E0 + c1) y + P(E0),
// whose smallest non-negative root y = E0 - lambda_max is the quantity the RMSD needs.  All coefficients are
// computed UNSCALED in FP64 (values up to ~2^240: no division, no normalisation multiplies).
struct I8Quartic { double q0, q1, q2, e0; };
__device__ __forceinline__ I8Quartic i8_coeffs(const double* S, double e0) {
  const double m00 = fma(S[6], S[6], fma(S[3], S[3], S[0] * S[0]));
  const double m11 = fma(S[7], S[7], fma(S[4], S[4], S[1] * S[1]));
  const double m22 = fma(S[8], S[8], fma(S[5], S[5], S[2] * S[2]));
  const double m01 = fma(S[6], S[7], fma(S[3], S[4], S[0] * S[1]));
  const double m02 = fma(S[6], S[8], fma(S[3], S[5], S[0] * S[2]));
  const double m12 = fma(S[7], S[8], fma(S[4], S[5], S[1] * S[2]));
  const double p1 = (m00 + m11) + m22;
  const double off = fma(m01, m01, fma(m02, m02, m12 * m12));
  const double trM2 = fma(2.0, off, fma(m00, m00, fma(m11, m11, m22 * m22)));
  const double det = fma(S[0], fma(S[4], S[8], -S[5] * S[7]), fma(-S[1], fma(S[3], S[8], -S[5] * S[6]),
                                                                  S[2] * fma(S[3], S[7], -S[4] * S[6])));
  const double c2 = -2.0 * p1, c1 = -8.0 * det, c0 = fma(2.0, trM2, -p1 * p1);
  const double e2 = e0 * e0;
  I8Quartic r;
  r.q0 = fma(fma(e2 + c2, e0, c1), e0, c0);
  r.q1 = -fma(fma(4.0, e2, 2.0 * c2), e0, c1);
  r.q2 = fma(6.0, e2, c2);
  r.e0 = e0;
  return r;
}
/// v * 2^e through the exponent field (integer pipe; zero, denormal and underflowing results give 0).
__device__ __forceinline__ double i8_scale2(double v, int e) {
  const int hi = __double2hiint(v);
  const int ex = ((hi >> 20) & 0x7ff);
  const int nx = ex + e;
  return (ex == 0 || nx <= 0) ? 0.0 : __hiloint2double((hi & 0x800fffff) | (nx << 20), __double2loint(v));
}
/// Smallest non-negative root of the quartic.  Exponent scaling to y' = y / 2^k, 2^k <= E0 < 2^(k+1), through the
/// exponent field (integer pipe), then the monotone Newton approach from y' = 0 on the FP32 pipe: nothing here touches
/// the FP64 pipe, so it runs at full speed under the MMAs.  The cancellation happened in the FP64 coefficients; near
/// the root Q(y') is evaluated in FP32 with an absolute error of ~2e-7 * sum |terms|, i.e. the root carries
/// ~2e-7 * sum|terms| / |Q'(y')|.  For RMSDs small against the radius of gyration (y' << gap to the next root) that
/// is 2e-7 RELATIVE, a few 1e-7 A.  An FP64 polish in the next FP64 window used to follow for EVERY pair: it changed
/// cfg2's RMSDs by < 1e-6 A and cost 250 cycles of every window (profiles/r2_pair_i8_experiments.md).  Now only pairs
/// whose estimated error exceeds 5e-6 A (RMSD comparable to the size of the molecule, or a small gap) are polished,
/// in FP64, here -- under the MMAs, where FP64 crawls, but off the tensor pipe's critical path and only when needed.
/// *r2 = (E0 - lambda_max) * 2/M.  Returns false when the root is ill-conditioned beyond what the polish repairs
/// (|Q'| tiny relative to lambda^3, or E0 not a positive integer-scale value): the caller takes the guarded FP64 path.
__device__ __forceinline__ bool i8_root(const I8Quartic& c, double unit, bool wanted, float* r2) {
  const int k = ((__double2hiint(c.e0) >> 20) & 0x7ff) - 1023;
  const double q0 = i8_scale2(c.q0, -4 * k), q1 = i8_scale2(c.q1, -3 * k), q2 = i8_scale2(c.q2, -2 * k);
  const double x0 = i8_scale2(c.e0, -k);   // E0 / 2^k in [1, 2)
  const double us = i8_scale2(unit, k);
  const float f0 = (float)q0, f1 = (float)q1, f2 = (float)q2, xf = (float)x0, unitf = (float)us;
  const float f3 = -4.f * xf;
  const float g3 = 3.f * f3, g2 = 2.f * f2;
  float y = 0.f, dq = f1, step;
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const float qy = fmaf(fmaf(fmaf(y + f3, y, f2), y, f1), y, f0);
    dq = fmaf(fmaf(fmaf(4.f, y, g3), y, g2), y, f1);
    y -= __fdividef(qy, dq);
  }
#pragma unroll 1
  for (int it = 0; it < 24; ++it) {   // until the whole warp has converged (gaps up to y' ~ 0.5: RMSD ~ radius of gyration)
    const float qy = fmaf(fmaf(fmaf(y + f3, y, f2), y, f1), y, f0);
    dq = fmaf(fmaf(fmaf(4.f, y, g3), y, g2), y, f1);
    step = __fdividef(qy, dq);
    y -= step;
    if (!__any_sync(0xffffffffu, fabsf(step) > 2e-7f * fabsf(y))) break;
  }
  dq = fmaf(fmaf(fmaf(4.f, y, g3), y, g2), y, f1);
  float r2f = y * unitf;
  // estimated RMSD error of the FP32 root: rmsd * dy / (2 y), dy = 2e-7 * sum|terms| / |Q'|
  const float terms = fmaf(fmaf(fmaf(y + fabsf(f3), y, fabsf(f2)), y, fabsf(f1)), y, fabsf(f0));
  // (`wanted`: lanes without a pair -- tile padding -- hold a quadruple root at E0 and must not crawl through FP64)
  if (wanted && 1e-7f * terms * unitf > 5e-6f * fabsf(dq) * sqrtf(fmaxf(r2f, 0.f))) {
    double yd = (double)y;
    const double q3 = -4.0 * x0;
    double d = fma(fma(fma(yd + q3, yd, q2), yd, q1), yd, q0) * (double)__frcp_rn(dq);
    yd -= d;
#pragma unroll 1
    for (int it = 0; it < 30 && fabs(d) > 1e-9 * fabs(yd) + 1e-14; ++it) {
      const double Qn = fma(fma(fma(yd + q3, yd, q2), yd, q1), yd, q0);
      const double dQ = fma(fma(fma(4.0, yd, 3.0 * q3), yd, 2.0 * q2), yd, q1);
      d = Qn / dQ;
      yd -= d;
    }
    r2f = (float)(yd * us);
  }
  *r2 = r2f;
  return (fabsf(dq) >= 7e-3f * xf * xf * xf) && (k > 0) && (y < 3.f);
}
/// Guarded FP64 path (Newton on the unscaled quartic, SVD for double roots); rarely taken.  The covariance
/// travels BY VALUE: a pointer parameter would force the caller's S into local memory on every pair.
__device__ __noinline__ double i8_relative_gap_slow(double s0, double s1, double s2, double s3, double s4, double s5,
                                                    double s6, double s7, double s8, double e0) {
  const double S[9] = {s0, s1, s2, s3, s4, s5, s6, s7, s8};
  return (e0 > 0.0) ? (e0 - largest_root(quartic_of(S), e0, S)) / e0 : 0.0;
}

template <bool TRI, int CG, bool DBG>
__global__ void __launch_bounds__(I8_THREADS, 1) pair_i8_kernel(PairI8Args a) {
  using GEO = I8Geom<CG>;
  constexpr int STAGES = GEO::STAGES, BPS = GEO::BPS, BBLK = GEO::BBLK;
  constexpr int STAGE_BYTES = i8_stage_bytes<CG>();
  extern __shared__ __align__(1024) unsigned char smem_i8[];
  unsigned char* stages = smem_i8;
  unsigned char* xbufRaw = smem_i8 + STAGES * STAGE_BYTES;
  long long* xbuf = reinterpret_cast<long long*>(xbufRaw);   // exchange buffer: exact int64 covariance entries
  uint64_t* bars = reinterpret_cast<uint64_t*>(xbufRaw + I8_XBUF_BYTES);
  uint64_t* fullBar = bars;                            // [STAGES]  operands landed (CG 2: in both CTAs, seen by the leader)
  uint64_t* emptyBar = bars + I8_MAX_STAGES;           // [STAGES]  MMAs reading the stage done
  uint64_t* accFull = bars + 2 * I8_MAX_STAGES;        // [2]  accumulator complete
  uint64_t* accEmpty = bars + 2 * I8_MAX_STAGES + 2;   // [2]  accumulator drained (CG 2: by both CTAs; leader's copy is used)
  uint64_t* xFull = bars + 2 * I8_MAX_STAGES + 4;      // [4]  exchange group written
  uint64_t* xEmpty = bars + 2 * I8_MAX_STAGES + 8;     // [4]  exchange group read
  uint64_t* fpDone = bars + 2 * I8_MAX_STAGES + 12;    // [1]  FP64 window of a tile closed (CG 2: by both CTAs; leader's copy)
  uint64_t* fpLocal = bars + 2 * I8_MAX_STAGES + 13;   // [1]  same, this CTA's solve warps only: the drain warps start on the next tile
  uint64_t* winOpen = bars + 2 * I8_MAX_STAGES + 14;   // [2]  the last MMA of a tile has been issued (B200_I8_EARLY_WINDOW)
  uint32_t* tmemBaseSlot = reinterpret_cast<uint32_t*>(bars + I8_NBARS);
  double* heldBuf = reinterpret_cast<double*>(xbufRaw + I8_XBUF_BYTES + 32 + 256);   // (I8_TPW == 2 only)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = CG == 1 ? 0u : cluster_ctarank();   // 0 = leader (issues the MMAs)
  const int dbgMode = DBG ? a.dbgMode : 0;
  long long* const dbgClk = DBG ? a.dbgClk : nullptr;
  // Ring depth: short selections (fewer stages per tile than ring slots) use a shallower ring.
  const int stagesPerTile = (a.nC + BPS - 1) / BPS;
  const int depth = stagesPerTile < STAGES ? stagesPerTile : STAGES;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      // CG 2, leader: its own producer's expect_tx arrive + the peer's relayed "my half has landed"
      mbar_init(smem_u32(&fullBar[s]), (CG == 2 && rank == 0) ? 2 : 1);
      mbar_init(smem_u32(&emptyBar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&accFull[b]), 1); mbar_init(smem_u32(&accEmpty[b]), I8_DRAIN_WARPS * CG); }
    for (int g = 0; g < 4; ++g) { mbar_init(smem_u32(&xFull[g]), 4); mbar_init(smem_u32(&xEmpty[g]), I8_SOLVE_WARPS / 2); }
    mbar_init(smem_u32(&winOpen[0]), 1); mbar_init(smem_u32(&winOpen[1]), 1);
    mbar_init(smem_u32(fpDone), I8_SOLVE_WARPS * CG);
    mbar_init(smem_u32(fpLocal), I8_SOLVE_WARPS);
    mbar_fence_init();
  }
  if (warp == I8_WARP_MMA) tmem_alloc<CG>(smem_u32(tmemBaseSlot), 512);
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();   // the peer's barriers are initialised before anything remote touches them
  tc_fence_after();
  const uint32_t tmemBase = *tmemBaseSlot;
  I8TileIter<TRI, CG> tiles;
  tiles.init(a);
  int it, jt;

  if (warp == I8_WARP_PRODUCER) {
    // ===================== TMA producer (one thread per CTA) =====================
    int stage = 0; uint32_t phase = 0;
    const uint32_t stage0 = smem_u32(stages);
    long long cwEmpty = 0;
    while (tiles.next(a, it, jt)) {
      // this CTA's A row group, and the B row group(s) it stages: both (CG 1) or the rank-th (CG 2)
      const int itA = (DBG && dbgMode == 4) ? a.it0 : it;   // timing experiment 4: every CTA streams the same A rows
      const uint8_t* gA = a.PA + (size_t)(CG * itA + (int)rank) * a.nC * I8_BLK_BYTES;
      const uint8_t* gB0 = a.PB + (size_t)(2 * jt + (CG == 2 ? (int)rank : 0)) * a.nC * I8_BLK_BYTES;
      const uint8_t* gB1 = gB0 + (size_t)a.nC * I8_BLK_BYTES;   // CG 1 only
      for (int c = 0; c < a.nC; c += BPS) {
        const int nb = (a.nC - c < BPS) ? a.nC - c : BPS;
        const long long c0 = (DBG && dbgClk) ? clock64() : 0;
        mbar_wait(smem_u32(&emptyBar[stage]), phase ^ 1u);
        if (DBG && dbgClk) cwEmpty += clock64() - c0;
        const uint32_t bar = smem_u32(&fullBar[stage]);
        const uint32_t dst = stage0 + (uint32_t)stage * STAGE_BYTES;
        if (!elect_one()) {
          // the other lanes only keep the loop converged
        } else if (DBG && dbgMode == 3) {
          mbar_arrive(bar);
        } else {
          mbar_expect_tx(bar, (uint32_t)(nb * (1 + BBLK) * I8_BLK_BYTES));
          // consecutive 64-atom blocks of one row group are contiguous in the image: one bulk copy per operand
          bulk_g2s(dst, gA + (size_t)c * I8_BLK_BYTES, (uint32_t)(nb * I8_BLK_BYTES), bar);
          if constexpr (CG == 2) {
            bulk_g2s(dst + BPS * I8_BLK_BYTES, gB0 + (size_t)c * I8_BLK_BYTES, (uint32_t)(nb * I8_BLK_BYTES), bar);
          } else {
#pragma unroll
            for (int cc = 0; cc < BPS; ++cc)
              if (cc < nb) {
                bulk_g2s(dst + (BPS + 2 * cc) * I8_BLK_BYTES, gB0 + (size_t)(c + cc) * I8_BLK_BYTES, I8_BLK_BYTES, bar);
                bulk_g2s(dst + (BPS + 2 * cc + 1) * I8_BLK_BYTES, gB1 + (size_t)(c + cc) * I8_BLK_BYTES, I8_BLK_BYTES, bar);
              }
          }
        }
        __syncwarp();
        if (++stage == depth) { stage = 0; phase ^= 1u; }
      }
    }
    if (DBG && dbgClk && lane == 0) dbgClk[16 * blockIdx.x + 0] += cwEmpty;
  } else if (warp == I8_WARP_MMA && rank == 0) {
    // ===================== MMA issuer (one thread; CG 2: of the leader CTA) =====================
    constexpr uint32_t idesc = umma_idesc_i8(128 * CG, 256);
    // descriptor = constant high part (LBO 128 B, SBO 512 B, version 1) | (smem address >> 4)
    constexpr uint64_t descHi = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46);
    const uint32_t stage0 = smem_u32(stages);
    int stage = 0; uint32_t phase = 0; int n = 0;
    long long cwAcc = 0, cwFull = 0, cwFpd = 0;
    const long long cStart = (DBG && dbgClk) ? clock64() : 0;
    while (tiles.next(a, it, jt)) {
      const int b = n & 1;
      long long c0 = (DBG && dbgClk) ? clock64() : 0;
      if constexpr (CG == 2) mbar_wait_cluster(smem_u32(&accEmpty[b]), (uint32_t)(((n >> 1) & 1) ^ 1));
      else mbar_wait(smem_u32(&accEmpty[b]), (uint32_t)(((n >> 1) & 1) ^ 1));
      // FP64 window: the solve warps finished the FP64 part of tile n-2 (they started it when tile n-1's MMAs completed)
      const long long cA = (DBG && dbgClk) ? clock64() : 0;
      if (!(DBG && ((dbgMode >= 2 && dbgMode <= 5) || dbgMode == 7))) {
        if (I8_TPW == 1) { if (n >= 2) mbar_wait(smem_u32(fpDone), (uint32_t)(n & 1)); }
        else if (n >= 2 && (n & 1) == 0) mbar_wait(smem_u32(fpDone), (uint32_t)(((n >> 1) - 1) & 1));   // window k closes before tile 2k+2
      }
      if (DBG && dbgClk) { const long long c2 = clock64(); cwAcc += c2 - c0; cwFpd += c2 - cA; }
      tc_fence_after();
      const uint32_t dTmem = tmemBase + (uint32_t)(b * 256);
      for (int c = 0; c < a.nC; c += BPS) {
        const int nb = (a.nC - c < BPS) ? a.nC - c : BPS;
        c0 = (DBG && dbgClk) ? clock64() : 0;
        if constexpr (CG == 2) mbar_wait_cluster(smem_u32(&fullBar[stage]), phase);
        else mbar_wait(smem_u32(&fullBar[stage]), phase);
        if (DBG && dbgClk) cwFull += clock64() - c0;
        tc_fence_after();
        const uint32_t sA = stage0 + (uint32_t)stage * STAGE_BYTES;
        const uint32_t sB = sA + BPS * I8_BLK_BYTES;
        if (elect_one()) {
        if (!(DBG && dbgMode == 5))   // timing experiment 5: operand pipeline without the MMAs
#pragma unroll
        for (int cc = 0; cc < BPS; ++cc) {
          if (cc < nb) {
            const uint64_t dA = descHi | (uint64_t)(((sA + cc * I8_BLK_BYTES) >> 4) & 0x3fff);
            const uint64_t dB = descHi | (uint64_t)(((sB + cc * BBLK * I8_BLK_BYTES) >> 4) & 0x3fff);
#pragma unroll
            for (int k = 0; k < I8_KC / 32; ++k)   // one K=32 step = two 16-byte core matrices = 256 B further on
              umma_i8<CG>(dTmem, dA + (uint64_t)(k * 16), dB + (uint64_t)(k * 16), idesc, (uint32_t)((c | cc | k) != 0));
          }
        }
        umma_commit<CG>(smem_u32(&emptyBar[stage]));
        if (c + BPS >= a.nC) {
          umma_commit<CG>(smem_u32(&accFull[b]));   // last stage of the tile: accumulator complete
          if (B200_I8_EARLY_WINDOW) {   // tell the solve warps of both CTAs now: their wake-up overlaps the MMAs' completion
            mbar_arrive(smem_u32(&winOpen[b]));
            if constexpr (CG == 2) mbar_arrive_cluster(cluster_map(smem_u32(&winOpen[b]), 1));
          }
        }
        }
        __syncwarp();
        if (++stage == depth) { stage = 0; phase ^= 1u; }
      }
      ++n;
    }
    if (DBG && dbgClk && lane == 0) {
      dbgClk[16 * blockIdx.x + 1] += cwAcc; dbgClk[16 * blockIdx.x + 2] += cwFull;
      dbgClk[16 * blockIdx.x + 3] += clock64() - cStart; dbgClk[16 * blockIdx.x + 4] += n;
      dbgClk[16 * blockIdx.x + 14] += cwFpd;
    }
  } else if (warp == I8_WARP_MMA) {
    // ===================== CG 2, peer CTA: relay "my operand half has landed" to the leader =====================
    if constexpr (CG == 2) {
      int stage = 0; uint32_t phase = 0;
      while (tiles.next(a, it, jt)) {
        for (int c = 0; c < a.nC; c += BPS) {
          mbar_wait(smem_u32(&fullBar[stage]), phase);
          if (elect_one()) mbar_arrive_cluster(cluster_map(smem_u32(&fullBar[stage]), 0));
          __syncwarp();
          if (++stage == depth) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp < I8_DRAIN_WARPS) {
    // ===================== drain warps =====================
    const int sp = warp & 3;          // TMEM sub-partition: lanes 32*sp .. 32*sp+31
    const int hh = warp >> 2;         // 8 drain warps: column half of the tile (exchange groups 2*hh, 2*hh+1); 4: all groups
    const int r = 32 * sp + lane;     // operand row of this thread: 9*i + 3*p + digit = 3*triple + digit
    long long* xdst = xbuf + 3 * r;
    int n = 0;
    long long cwAcc = 0, cwX = 0, cwLd = 0, cwFold = 0;
    const long long cStart = (DBG && dbgClk) ? clock64() : 0;
    // One column frame: 9 accumulators (q-major, B digit minor) -> 3 exact int64 sums over the B digits.  INTEGER
    // ONLY: on this part FP64 instructions share the tensor pipe with tcgen05.mma and crawl (~10 % of their rate)
    // while MMAs are in flight (tools/microbench/mma_fp64_mix.cu), so the drain must not touch the FP64 pipe; and
    // no cross-lane work either: the A digits (three operand rows = three lanes) are folded by the solve threads.
    auto fold_frame = [&](const int* v, long long* dst) {
#pragma unroll
      for (int q = 0; q < 3; ++q)
        dst[q] = (long long)v[3 * q] + (long long)v[3 * q + 1] * 256 + (long long)v[3 * q + 2] * 65536;
    };
    while (tiles.next(a, it, jt)) {
      const int b = n & 1;
      long long c0 = (DBG && dbgClk) ? clock64() : 0;
      mbar_wait(smem_u32(&accFull[b]), (uint32_t)((n >> 1) & 1));
      // The FP64 window of tile n-1 opens at this very moment (the solve warps wait for the same barrier).  The drain
      // is integer work that would take every other issue slot of the four schedulers while the window's FP64
      // instructions want them: start it when this CTA's solve warps are through (there is slack: the drain of a tile
      // takes ~3600 cycles, the MMAs of the next one ~4100).
      if (!(DBG && ((dbgMode >= 2 && dbgMode <= 5) || dbgMode == 7))) {
        if (I8_TPW == 1) { if (n >= 1) mbar_wait(smem_u32(fpLocal), (uint32_t)((n - 1) & 1)); }
        else if (n & 1) mbar_wait(smem_u32(fpLocal), (uint32_t)(((n - 1) >> 1) & 1));   // odd tiles complete as a window opens
      }
      if (DBG && dbgClk) cwAcc += clock64() - c0;
      tc_fence_after();
      if (DBG && dbgMode >= 2 && dbgMode <= 5) {   // timing experiment: MMA + operand pipeline only
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster(cluster_map(smem_u32(&accEmpty[b]), 0));
          else mbar_arrive(smem_u32(&accEmpty[b]));
        }
        ++n;
        continue;
      }
#pragma unroll 1
      for (int gg = 0; gg < I8_GROUPS_PER_DRAIN_WARP; ++gg) {
        const int g = I8_GROUPS_PER_DRAIN_WARP * hh + gg;
        // column frames 7g .. 7g+6: accumulator columns 9*jl (frames 0..13) or 128 + 9*(jl-14) (frames 14..27)
        const uint32_t tcol = tmemBase + (uint32_t)(b * 256 + (g >> 1) * 128 + (g & 1) * 63) + ((uint32_t)(32 * sp) << 16);
        long long* dst = xdst + (size_t)(7 * g) * I8_XJ_DBL;
        int v[32];
        long long c1 = (DBG && dbgClk) ? clock64() : 0;
        if (DBG && dbgMode == 8) {   // timing experiment 8: no TMEM loads
#pragma unroll
          for (int x = 0; x < 32; ++x) v[x] = x + lane;
        } else
        tmem_ld32(tcol, v);                     // frames 0..2 of the group (27 columns)
        c0 = (DBG && dbgClk) ? clock64() : 0;
        mbar_wait(smem_u32(&xEmpty[g]), (uint32_t)((n & 1) ^ 1));   // the solve warps are done with this group of tile n-1
        if (DBG && dbgClk) { const long long c2 = clock64(); cwX += c2 - c0; c1 += c2 - c0; }
        tmem_ld_wait();
        if (DBG && dbgClk) { const long long c2 = clock64(); cwLd += c2 - c1; c1 = c2; }
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) fold_frame(v + 9 * jj, dst + jj * I8_XJ_DBL);
        if (DBG && dbgClk) { const long long c2 = clock64(); cwFold += c2 - c1; c1 = c2; }
        if (!(DBG && dbgMode == 8)) tmem_ld32(tcol + 27, v);                // frames 3..5
        tmem_ld_wait();
        if (DBG && dbgClk) { const long long c2 = clock64(); cwLd += c2 - c1; c1 = c2; }
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) fold_frame(v + 9 * jj, dst + (3 + jj) * I8_XJ_DBL);
        if (DBG && dbgClk) { const long long c2 = clock64(); cwFold += c2 - c1; c1 = c2; }
        if (!(DBG && dbgMode == 8)) tmem_ld16(tcol + 47, v);                // columns 47..62: frame 6 is the last 9 (stays inside the accumulator)
        tmem_ld_wait();
        if (DBG && dbgClk) { const long long c2 = clock64(); cwLd += c2 - c1; c1 = c2; }
        if (gg == I8_GROUPS_PER_DRAIN_WARP - 1) {   // this warp is done reading accumulator buffer b
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_cluster(cluster_map(smem_u32(&accEmpty[b]), 0));
            else mbar_arrive(smem_u32(&accEmpty[b]));
          }
        }
        fold_frame(v + 7, dst + 6 * I8_XJ_DBL);
        if (DBG && dbgClk) { const long long c2 = clock64(); cwFold += c2 - c1; }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&xFull[g]));
      }
      ++n;
    }
    if (DBG && dbgClk && lane == 0 && warp == 0) {
      dbgClk[16 * blockIdx.x + 5] += cwAcc; dbgClk[16 * blockIdx.x + 6] += cwX;
      dbgClk[16 * blockIdx.x + 7] += clock64() - cStart;
      dbgClk[16 * blockIdx.x + 10] += cwLd; dbgClk[16 * blockIdx.x + 11] += cwFold;
    }
  } else {
    // ===================== solve warps: one thread per frame pair =====================
    const int u = warp - I8_DRAIN_WARPS;          // 0..13
    const int h = u & 1;                          // column half of the tile: frames 14h .. 14h+13 = exchange groups 2h, 2h+1
    const int il = 2 * (u >> 1) + (lane >> 4);    // row frame of the tile
    const int jlRaw = lane & 15;
    const bool laneOn = jlRaw < I8_TILE_I;
    const int jl = I8_TILE_I * h + (laneOn ? jlRaw : I8_TILE_I - 1);
    const double outScale = 2.0 * a.invScale2 / a.totalMass[0];   // rmsd^2 = (E0 - lambda) * outScale (integer units)
    const double halfToInt = 0.5 / a.invScale2;                    // G (A^2) -> integer units, halved (exact power of two)
    const long long* src = xbuf + (size_t)jl * I8_XJ_DBL + 27 * il;   // [row 9*il + 3*p + digit][q]
    const bool solveOff = DBG && dbgMode >= 2 && dbgMode <= 5;
    const uint32_t fpDoneAddr = CG == 2 ? cluster_map(smem_u32(fpDone), 0) : smem_u32(fpDone);
    int n = 0;
    long long cwX = 0, cwWin = 0, cwFp = 0, cwRoot = 0;
    const long long cStart = (DBG && dbgClk) ? clock64() : 0;
    const bool plain = !(DBG && (dbgMode == 1 || dbgMode == 6));   // timing experiments 1 / 6: handshakes without the per-pair solve
    // Output slot of pair (i, j); 0 for pairs this thread does not own (never stored).
    auto out_index = [&](int i, int j, bool valid) -> size_t {
      if (TRI) return valid ? tri_row_start((size_t)a.nCols, (size_t)i) + (size_t)(j - i - 1) - a.outBase : 0;
      return (size_t)i * a.ldo + (size_t)j;
    };
    // FP32 root + store of one pair (under the MMAs; no FP64 instruction unless the root is ill-conditioned).
    // clamp: src/Frame.cpp:1264-1268; the result is stored as float (Matrix<float>), so the square root is taken in float.
    auto finish_pair = [&](const I8Quartic& cq, const double* S, bool valid, size_t idx) {
      float r2;
      const bool ok = i8_root(cq, outScale, valid, &r2);
      if (valid && !ok)
        r2 = (float)(i8_relative_gap_slow(S[0], S[1], S[2], S[3], S[4], S[5], S[6], S[7], S[8], cq.e0) * (cq.e0 * outScale));
      if (valid) a.out[idx] = (r2 > 0.f) ? sqrtf(r2) : 0.f;
    };
    // the same for the held tile: its covariance is re-read from shared memory only on the (rare) guarded path
    auto finish_pair_held = [&](const I8Quartic& cq, bool valid, size_t idx) {
      float r2;
      const bool ok = i8_root(cq, outScale, valid, &r2);
      if (valid && !ok) {
        const double* hp = heldBuf + (u * 32 + lane);
        constexpr int HS2 = I8_SOLVE_WARPS * 32;
        r2 = (float)(i8_relative_gap_slow(hp[0], hp[HS2], hp[2 * HS2], hp[3 * HS2], hp[4 * HS2], hp[5 * HS2], hp[6 * HS2], hp[7 * HS2],
                                          hp[8 * HS2], cq.e0) * (cq.e0 * outScale));
      }
      if (valid) a.out[idx] = (r2 > 0.f) ? sqrtf(r2) : 0.f;
    };
    // A tile held back for the next window (I8_TPW == 2): its covariance and 2 E0 wait in shared memory, its output slot
    // in registers.
    double* held = heldBuf + (u * 32 + lane);
    constexpr int HS = I8_SOLVE_WARPS * 32;   // stride between the entries of one thread
    bool heldLive = false, heldValid = false;
    size_t heldIdx = 0;
    int itN, jtN;
    bool have = tiles.next(a, it, jt);
    while (have) {
      const bool hasNext = tiles.next(a, itN, jtN);   // one tile of look-ahead: is there an MMA to wait for?
      if (solveOff) { ++n; it = itN; jt = jtN; have = hasNext; continue; }   // (the MMA warp does not wait for fpDone in these modes)
      const int i = I8_TILE_I * (CG * it + (int)rank) + il, j = I8_TILE_J * jt + jl;
      const bool valid = laneOn && i < a.nRows && j < a.nCols && i >= a.rowLo && i < a.rowHi && (!TRI || j > i);
      double ga = 0.5, gb = 0.5;
      if (valid) { ga = __ldg(a.GA + i); gb = __ldg(a.GB + j); }
      // ---- 1. the exact integer covariance of the pair (under the MMAs of tile n+1)
      long long c0 = (DBG && dbgClk) ? clock64() : 0;
      mbar_wait(smem_u32(&xFull[2 * h]), (uint32_t)(n & 1));
      mbar_wait(smem_u32(&xFull[2 * h + 1]), (uint32_t)(n & 1));
      if (DBG && dbgClk) cwX += clock64() - c0;
      double S[9];
      if (!(DBG && dbgMode == 6)) {
        long long SI[9];
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
          for (int q = 0; q < 3; ++q)   // fold the A digits: exact int64 (|S| < 2^63 for < 131072 atoms)
            SI[3 * p + q] = src[9 * p + q] + (src[9 * p + 3 + q] << 8) + (src[9 * p + 6 + q] << 16);
        if (DBG && dbgMode == 7) {
          long long* ring = a.dbgRing + ((size_t)(blockIdx.x * 4 + (n & 3)) * 10) * 448 + (u * 32 + lane);
#pragma unroll
          for (int x = 0; x < 9; ++x) __stcg(ring + x * 448, SI[x]);
          __stcg(ring + 9 * 448, valid ? (((long long)i << 32) | (long long)j) : -1ll);
        }
#pragma unroll
        for (int x = 0; x < 9; ++x) S[x] = __ll2double_rn(SI[x]);   // I2F.F64.S64: conversion pipe, not the FP64 pipe
      } else {
#pragma unroll
        for (int x = 0; x < 9; ++x) S[x] = 0.0;
      }
      // All nine are materialised here (the loads have completed, the conversions are done; no FP64-pipe instruction
      // under the MMAs): hand the exchange groups back to the drain warps.
      asm volatile("" ::"d"(S[0]), "d"(S[1]), "d"(S[2]), "d"(S[3]), "d"(S[4]), "d"(S[5]), "d"(S[6]), "d"(S[7]), "d"(S[8]) : "memory");
      __syncwarp();
      if (lane == 0) { mbar_arrive(smem_u32(&xEmpty[2 * h])); mbar_arrive(smem_u32(&xEmpty[2 * h + 1])); }
      if (DBG && dbgMode == 7) { ++n; it = itN; jt = jtN; have = hasNext; continue; }
      if (DBG && a.dbgS && valid) {
#pragma unroll
        for (int x = 0; x < 9; ++x) a.dbgS[((size_t)i * a.nCols + j) * 9 + x] = S[x];
      }
      const size_t idx = out_index(i, j, valid);
      // ---- 2. FP64 window, once per I8_TPW tiles.  It opens when the MMAs of tile n+1 have completed and closes when
      //         every solve warp of the CTA (pair) has arrived on fpDone: the MMA warp issues tile n+2 only then.
      //         Each window costs ~600 cycles of idle tensor pipe beyond its FP64 work (pipe drain, handshake, refill):
      //         with two tiles per window -- the odd tile's covariance waits in registers -- half as many are paid.
      const bool windowTile = I8_TPW == 1 || (n & 1) == 0 || !hasNext;
      if (!windowTile) {
#pragma unroll
        for (int x = 0; x < 9; ++x) held[x * HS] = S[x];
        held[9 * HS] = ga + gb;   // (one FP64 add under the MMAs per held tile)
        heldLive = true; heldValid = valid; heldIdx = idx;
      } else {
        c0 = (DBG && dbgClk) ? clock64() : 0;
        if (hasNext) mbar_wait(smem_u32(B200_I8_EARLY_WINDOW ? &winOpen[(n + 1) & 1] : &accFull[(n + 1) & 1]), (uint32_t)(((n + 1) >> 1) & 1));
        const long long cF = (DBG && dbgClk) ? clock64() : 0;
        I8Quartic cq, cqh;
        cq.q0 = cq.q1 = cq.q2 = 0.0; cq.e0 = 1.0; cqh = cq;
        if (plain) {
          // everything that must be FP64 (E0 too: outside the window its instructions crawl on the critical path)
          if (I8_TPW == 2 && heldLive) {
            double Sh[9];
#pragma unroll
            for (int x = 0; x < 9; ++x) Sh[x] = held[x * HS];
            cqh = i8_coeffs(Sh, held[9 * HS] * halfToInt);
          }
          cq = i8_coeffs(S, (ga + gb) * halfToInt);
          asm volatile("" ::"d"(cq.q0), "d"(cq.q1), "d"(cq.q2), "d"(cqh.q0), "d"(cqh.q1), "d"(cqh.q2) : "memory");   // computed before the arrive below
        }
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster(fpDoneAddr); else mbar_arrive(fpDoneAddr);
          mbar_arrive(smem_u32(fpLocal));
        }
        if (DBG && dbgClk) { const long long c2 = clock64(); cwWin += c2 - c0; cwFp += c2 - cF; }
        // ---- 3. roots and stores (under the MMAs of the next tiles)
        const long long cR = (DBG && dbgClk) ? clock64() : 0;
        if (plain) {
          if (I8_TPW == 2 && heldLive) finish_pair_held(cqh, heldValid, heldIdx);
          finish_pair(cq, S, valid, idx);
        } else if (!(DBG && dbgMode == 6)) {
          if (valid) a.out[idx] = (float)(S[0] + S[4] + S[8]);
        }
        heldLive = false;
        if (DBG && dbgClk) cwRoot += clock64() - cR;
      }
      ++n;
      it = itN; jt = jtN; have = hasNext;
    }
    if (DBG && dbgClk && lane == 0 && u == 0) {
      dbgClk[16 * blockIdx.x + 8] += cwX; dbgClk[16 * blockIdx.x + 9] += clock64() - cStart;
      dbgClk[16 * blockIdx.x + 12] += cwWin; dbgClk[16 * blockIdx.x + 13] += cwFp; dbgClk[16 * blockIdx.x + 15] += cwRoot;
    }
  }
  // ---- teardown ----
  __syncwarp();   // single-lane roles: reconverge before the (warp-aligned) block and cluster barriers
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();   // the pair's MMAs, multicast commits and remote arrives are all done
  tc_fence_after();
  if (warp == I8_WARP_MMA) tmem_dealloc<CG>(tmemBase, 512);
}


// ----------------------------------------------------------------------------
// tcgen05 kind::i8 issue-peak probe: one CTA per SM, operands fixed in shared memory (no loads),
// one thread issues M128 x N256 x K32 MMAs back to back.  Roofline denominator of pair_i8_kernel.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) i8_mma_peak_kernel(int iters, int* sink, int variant) {
  // variant 0: consecutive MMAs alternate between the two accumulators (independent);
  //         1: all MMAs accumulate into one accumulator (the dependent chain of a real K loop);
  //         2: as 1, plus a tcgen05.commit to a (never waited) mbarrier after every second MMA, as the pair kernel does
  extern __shared__ __align__(1024) unsigned char smem_pk[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmemSlot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (3 * I8_BLK_BYTES) / 4; i += 128) reinterpret_cast<uint32_t*>(smem_pk)[i] = 0x01010101u * (uint32_t)(i & 3);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<1>(smem_u32(&tmemSlot), 512);
  // make the generic-proxy smem writes visible to the async (tensor core) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmemBase = tmemSlot;
  if (tid == 0) {
    constexpr uint32_t idesc = umma_idesc_i8(128, 256);
    const uint32_t sA = smem_u32(smem_pk), sB = sA + I8_BLK_BYTES;
    for (int it = 0; it < iters; ++it) {
      const uint32_t buf = variant == 0 ? (uint32_t)((it & 1) * 256) : 0u;
      umma_i8<1>(tmemBase + buf, umma_desc(sA + (it & 1) * 256, 128, 512),
              umma_desc(sB + (it & 1) * 256, 128, 512), idesc, (uint32_t)((variant == 0) ? (it > 1) : ((it & 31) != 0)));
      if (variant == 2 && (it & 1)) umma_commit<1>(smem_u32(&bar2));
    }
    umma_commit<1>(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
  }
  __syncthreads();
  if (warp == 0) {
    int v[4];
    tmem_ld4(tmemBase, v);
    tmem_ld_wait();
    if (sink && v[0] == 0x7fffffff) sink[blockIdx.x] = v[1];
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tmemBase, 512);
}


// ----------------------------------------------------------------------------
// tcgen05 latency probe: one thread issues nMma MMAs (M128 N256 K32, one accumulator), commits to an
// mbarrier and waits for it; returns the average cycles from first issue to wake-up.  nMma = 0 measures
// the commit -> mbarrier -> try_wait round trip alone.  Sizes the operand ring of pair_i8_kernel.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) i8_mma_latency_kernel(int nMma, int reps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_pk[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmemSlot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (3 * I8_BLK_BYTES) / 4; i += 128) reinterpret_cast<uint32_t*>(smem_pk)[i] = 0x01010101u * (uint32_t)(i & 3);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<1>(smem_u32(&tmemSlot), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmemBase = tmemSlot;
  if (tid == 0) {
    constexpr uint32_t idesc = umma_idesc_i8(128, 256);
    const uint32_t sA = smem_u32(smem_pk), sB = sA + I8_BLK_BYTES;
    long long tot = 0;
    uint32_t phase = 0;
    for (int r = 0; r < reps + 2; ++r) {
      const long long t0 = clock64();
      for (int k = 0; k < nMma; ++k)
        umma_i8<1>(tmemBase, umma_desc(sA + (k & 1) * 256, 128, 512), umma_desc(sB + (k & 1) * 256, 128, 512), idesc, (uint32_t)(k != 0));
      umma_commit<1>(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), phase);
      phase ^= 1u;
      tc_fence_after();
      const long long t1 = clock64();
      if (r >= 2) tot += t1 - t0;
    }
    out[blockIdx.x] = tot / reps;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_before(); tmem_dealloc<1>(tmemBase, 512); }
}

}  // namespace b200
