// pair_i8.cuh -- tcgen05 (5th-gen tensor core) path of the all-pairs best-fit RMSD.
//
// Idea ("exact integer covariance"): best-fit RMSD is 1-Lipschitz in the RMS coordinate
// perturbation, so the centred, sqrt(mass)-scaled coordinates are rounded ONCE per frame to a
// fixed-point grid of spacing 2^-qs (24-bit signed integers; qs chosen from the largest centred
// coordinate so the worst-case RMSD change stays < 5.3e-5 A) and from then on everything is exact:
//   * every integer is split into three balanced signed base-256 digits (int8),
//   * the nine covariance entries of a frame pair are 81 int8 x int8 dot products over the atoms,
//     computed by tcgen05.mma kind::i8 with int32 accumulators in TMEM (exact for < 131072 atoms),
//   * the epilogue recombines the digits in integer/FP64 arithmetic (S and G carry < 2^-52
//     relative error), and runs the same FP64 per-pair solve as the FP64 DMMA path.
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
// One MMA tile = 14 x 28 frame pairs = D[128 x 256] int32 = 256 TMEM columns; two accumulator
// buffers fill the 512 columns, so the epilogue of tile n overlaps the MMAs of tile n+1.
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
/// groups, so shared-memory fill and operand-read traffic per SM drop by a third and a stage can hold
/// 128 atoms (4 MMAs) in the same footprint.
template <int CG> struct I8Geom;
template <> struct I8Geom<1> {
  static constexpr int BPS = 1;      // 64-atom blocks per stage
  static constexpr int STAGES = 5;
  static constexpr int BBLK = 2;     // B blocks per 64 atoms held by this CTA
};
template <> struct I8Geom<2> {
  static constexpr int BPS = 2;
  static constexpr int STAGES = 4;
  static constexpr int BBLK = 1;
};
template <int CG> __host__ __device__ constexpr int i8_stage_bytes() { return I8Geom<CG>::BPS * (1 + I8Geom<CG>::BBLK) * I8_BLK_BYTES; }
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
  const float* crd; size_t stride; const int* frameIdx; long srcBase;
  int nFrames; int f0;
  const int* atomIdx; int nAtoms;
  const double* centerMass; const double* covMass;
  double* centers;          // 3 per frame
  unsigned int* maxAbsBits; // float bits of max |(x-c)*sqrt(m)|, atomicMax
};

/// One warp per frame: centre (src/Frame.cpp:1043-1055 / :1141-1166) and the extent.
__global__ void __launch_bounds__(256) i8_stats_kernel(I8StatsArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = a.f0 + blockIdx.x * 8 + warp;
  if (f >= a.nFrames) return;
  const long row = (a.frameIdx ? (long)a.frameIdx[f] : (long)f) - a.srcBase;
  const float* src = a.crd + (size_t)row * a.stride;
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
  const float* crd; size_t stride; const int* frameIdx; long srcBase;
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
__global__ void __launch_bounds__(256) i8_quant_kernel(I8QuantArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = a.f0 + blockIdx.x * 8 + warp;
  if (f >= a.nFrames) return;
  const long row = (a.frameIdx ? (long)a.frameIdx[f] : (long)f) - a.srcBase;
  const float* src = a.crd + (size_t)row * a.stride;
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
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
  if (lane == 0) a.G[f] = (double)gsum * a.invScale2;
}

// ----------------------------------------------------------------------------
// tcgen05 / TMEM helpers
// ----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
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
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t clusterAddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(clusterAddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITC_LOOP:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAITC_DONE;\n"
      "bra.uni WAITC_LOOP;\n"
      "WAITC_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
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
// Persistent, warp-specialised, one CTA per SM.  Roles (20 warps):
//   warp 18       TMA producer: operand stages (A block + 2 B blocks per 64 atoms) through a
//                 5-deep full/empty mbarrier ring
//   warp 19       MMA issuer: tcgen05.mma kind::i8 M128 N256 K32 into one of two TMEM accumulators
//   warps 0..3    drain: one per TMEM sub-partition; tcgen05.ld 64 accumulator columns (7 column
//                 frames x 9), fold the three B digits in-thread, apply the A digit weight of
//                 the own operand row, hand 21 doubles per thread to the exchange buffer
//   warps 4..17   solve: one thread per frame pair (2 row frames x 14 column frames per warp):
//                 gather the 27 partial sums of the pair, fold the A digits, per-pair solve,
//                 store the float straight into cpptraj's Matrix<float> layout
// The exchange buffer is a ring of four column groups (7 column frames each) with its own
// full/empty mbarriers, so the drain of tile n+1 runs under the solve of tile n, both under the
// MMAs of tile n+2: no CTA-wide barrier anywhere in the steady state.
// Tiles are enumerated column tile by column tile (valid row tiles only), tile t of that list
// going to CTA t % gridDim.x: every CTA gets the same number of tiles (+-1) and CTAs running
// at the same time share their B operand in L2.
// ----------------------------------------------------------------------------
constexpr int I8_DRAIN_WARPS = 4;
constexpr int I8_SOLVE_WARPS = 14;
constexpr int I8_WARP_PRODUCER = I8_DRAIN_WARPS + I8_SOLVE_WARPS;   // 18
constexpr int I8_WARP_MMA = I8_WARP_PRODUCER + 1;                   // 19
constexpr int I8_THREADS = 32 * (I8_WARP_MMA + 1);                  // 640
constexpr int I8_XROW_BYTES = 24;            // 3 doubles per (operand row, column frame)
constexpr int I8_XJ_STRIDE = 128 * I8_XROW_BYTES + 88;  // 3160 B: (stride/4) % 32 == 22 -> conflict-free LDS.64 / STS.64
constexpr int I8_XBUF_BYTES = I8_TILE_J * I8_XJ_STRIDE;  // 88480
constexpr int I8_MAX_STAGES = 5;
constexpr int I8_NBARS = 2 * I8_MAX_STAGES + 4 + 8;
template <int CG> __host__ __device__ constexpr int i8_smem_bytes() { return I8Geom<CG>::STAGES * i8_stage_bytes<CG>() + I8_XBUF_BYTES + 256; }   // + barriers
static_assert(I8_NBARS * 8 + 8 <= 256, "barrier block");
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
  const double* totalMass;
  double invScale2;    // 2^-2qs: integer covariance -> A^2
  float* out;          // TRI: out[triIndex - outBase]; FULL: out[i*ldo + j]
  size_t outBase;
  size_t ldo;
  double* dbgS;        // nullable: 9 doubles per (i,j) at (i*nCols + j)*9, integer units
  long long* dbgClk;   // nullable: per-CTA cycle counters [16] (timing experiments)
  int dbgMode;         // 0 normal; timing experiments: 1 no per-pair solve, 2 drain only frees TMEM,
                       // 3 = 2 + no operand loads, 6 = solve warps only recycle the exchange buffer
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
/// Walks this CTA group's share of the tile list; every role keeps its own copy (two registers).
template <bool TRI, int CG>
struct I8TileIter {
  int jt, base, t;
  __device__ __forceinline__ void init(const PairI8Args& a) { jt = a.jt0; base = 0; t = (int)blockIdx.x / CG; }
  __device__ __forceinline__ bool next(const PairI8Args& a, int& it, int& jtOut) {
    while (jt < a.jt1) {
      const int c = i8_col_tiles<TRI, CG>(a.it0, a.it1, jt);
      if (t < base + c) { it = a.it0 + (t - base); jtOut = jt; t += (int)gridDim.x / CG; return true; }
      base += c; ++jt;
    }
    return false;
  }
};

/// (E0 - lambda_max)/E0 for the tcgen05 epilogue.  FP64 dependent-issue latency is long on this part, so
/// the FP64 chain is kept as short as the cancellation allows.  With everything divided by E0 the key-matrix
/// quartic is P(x) = x^4 + c2 x^2 + c1 x + c0 with the wanted root x = lambda/E0 in (0,1]; x = 1 - y gives
///     Q(y) = y^4 - 4 y^3 + (6 + c2) y^2 - (4 + 2 c2 + c1) y + (1 + c2 + c1 + c0),
/// whose smallest non-negative root y is the quantity the RMSD needs.  The coefficients (q0 is a
/// cancellation down to ~y) and ONE final Newton correction are FP64; the monotone approach from y = 0 runs
/// in FP32, which resolves y to ~1e-7 RELATIVE whatever its magnitude.  Returns false when the root is
/// ill-conditioned or the correction is not small: the caller then takes the guarded FP64 path.
__device__ __forceinline__ bool i8_relative_gap(const double* S, double sInv, double& gap) {
  double T[9];
#pragma unroll
  for (int x = 0; x < 9; ++x) T[x] = S[x] * sInv;
  const double m00 = fma(T[6], T[6], fma(T[3], T[3], T[0] * T[0]));
  const double m11 = fma(T[7], T[7], fma(T[4], T[4], T[1] * T[1]));
  const double m22 = fma(T[8], T[8], fma(T[5], T[5], T[2] * T[2]));
  const double m01 = fma(T[6], T[7], fma(T[3], T[4], T[0] * T[1]));
  const double m02 = fma(T[6], T[8], fma(T[3], T[5], T[0] * T[2]));
  const double m12 = fma(T[7], T[8], fma(T[4], T[5], T[1] * T[2]));
  const double p1 = (m00 + m11) + m22;
  const double off = fma(m01, m01, fma(m02, m02, m12 * m12));
  const double trM2 = fma(2.0, off, fma(m00, m00, fma(m11, m11, m22 * m22)));
  const double det = fma(T[0], fma(T[4], T[8], -T[5] * T[7]), fma(-T[1], fma(T[3], T[8], -T[5] * T[6]),
                                                                  T[2] * fma(T[3], T[7], -T[4] * T[6])));
  const double c2 = -2.0 * p1, c1 = -8.0 * det, c0 = fma(2.0, trM2, -p1 * p1);
  const double q0 = ((1.0 + c2) + c1) + c0;
  const double q1 = -((4.0 + 2.0 * c2) + c1);
  const double q2 = 6.0 + c2;
  const float f0 = (float)q0, f1 = (float)q1, f2 = (float)q2;
  float y = 0.f, dq = f1, step;
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const float qy = fmaf(fmaf(fmaf(y - 4.f, y, f2), y, f1), y, f0);
    dq = fmaf(fmaf(fmaf(4.f, y, -12.f), y, 2.f * f2), y, f1);
    y -= __fdividef(qy, dq);
  }
#pragma unroll 1
  for (int it = 0; it < 24; ++it) {   // until the whole warp has converged (gaps up to y ~ 0.5: RMSD ~ radius of gyration)
    const float qy = fmaf(fmaf(fmaf(y - 4.f, y, f2), y, f1), y, f0);
    dq = fmaf(fmaf(fmaf(4.f, y, -12.f), y, 2.f * f2), y, f1);
    step = __fdividef(qy, dq);
    y -= step;
    if (!__any_sync(0xffffffffu, fabsf(step) > 4e-7f * fabsf(y))) break;
  }
  double yd = (double)y;
  const double Q = fma(fma(fma(yd - 4.0, yd, q2), yd, q1), yd, q0);
  dq = fmaf(fmaf(fmaf(4.f, y, -12.f), y, 2.f * f2), y, f1);
  const double d = Q * (double)__frcp_rn(dq);
  yd -= d;
  gap = yd;
  // conditioning (|P'| relative to lambda^3 = 1 here) and size of the correction (FP32 left ~1e-7 relative)
  return (fabsf(dq) >= 7e-3f) && (fabs(d) <= 1e-5 * fabs(yd) + 1e-14) && (yd < 1.5);
}
/// Guarded FP64 path (Newton on the unscaled quartic, SVD for double roots); rarely taken.
__device__ __noinline__ double i8_relative_gap_slow(const double* S, double e0) {
  return (e0 > 0.0) ? (e0 - largest_root(quartic_of(S), e0, S)) / e0 : 0.0;
}

template <bool TRI, int CG>
__global__ void __launch_bounds__(I8_THREADS, 1) pair_i8_kernel(PairI8Args a) {
  using GEO = I8Geom<CG>;
  constexpr int STAGES = GEO::STAGES, BPS = GEO::BPS, BBLK = GEO::BBLK;
  constexpr int STAGE_BYTES = i8_stage_bytes<CG>();
  extern __shared__ __align__(1024) unsigned char smem_i8[];
  unsigned char* stages = smem_i8;
  unsigned char* xbuf = smem_i8 + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xbuf + I8_XBUF_BYTES);
  uint64_t* fullBar = bars;                            // [STAGES]  operands landed (CG 2: in both CTAs, seen by the leader)
  uint64_t* emptyBar = bars + I8_MAX_STAGES;           // [STAGES]  MMAs reading the stage done
  uint64_t* accFull = bars + 2 * I8_MAX_STAGES;        // [2]  accumulator complete
  uint64_t* accEmpty = bars + 2 * I8_MAX_STAGES + 2;   // [2]  accumulator drained (CG 2: by both CTAs; leader's copy is used)
  uint64_t* xFull = bars + 2 * I8_MAX_STAGES + 4;      // [4]  exchange group written
  uint64_t* xEmpty = bars + 2 * I8_MAX_STAGES + 8;     // [4]  exchange group read
  uint32_t* tmemBaseSlot = reinterpret_cast<uint32_t*>(bars + I8_NBARS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = CG == 1 ? 0u : cluster_ctarank();   // 0 = leader (issues the MMAs)

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      // CG 2, leader: its own producer's expect_tx arrive + the peer's relayed "my half has landed"
      mbar_init(smem_u32(&fullBar[s]), (CG == 2 && rank == 0) ? 2 : 1);
      mbar_init(smem_u32(&emptyBar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&accFull[b]), 1); mbar_init(smem_u32(&accEmpty[b]), I8_DRAIN_WARPS * CG); }
    for (int g = 0; g < 4; ++g) { mbar_init(smem_u32(&xFull[g]), I8_DRAIN_WARPS); mbar_init(smem_u32(&xEmpty[g]), I8_SOLVE_WARPS / 2); }
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
    if (lane == 0)
    while (tiles.next(a, it, jt)) {
      // this CTA's A row group, and the B row group(s) it stages: both (CG 1) or the rank-th (CG 2)
      const uint8_t* gA = a.PA + (size_t)(CG * it + (int)rank) * a.nC * I8_BLK_BYTES;
      const uint8_t* gB0 = a.PB + (size_t)(2 * jt + (CG == 2 ? (int)rank : 0)) * a.nC * I8_BLK_BYTES;
      const uint8_t* gB1 = gB0 + (size_t)a.nC * I8_BLK_BYTES;   // CG 1 only
      for (int c = 0; c < a.nC; c += BPS) {
        const int nb = (a.nC - c < BPS) ? a.nC - c : BPS;
        const long long c0 = a.dbgClk ? clock64() : 0;
        mbar_wait(smem_u32(&emptyBar[stage]), phase ^ 1u);
        if (a.dbgClk) cwEmpty += clock64() - c0;
        const uint32_t bar = smem_u32(&fullBar[stage]);
        const uint32_t dst = stage0 + (uint32_t)stage * STAGE_BYTES;
        if (a.dbgMode == 3) {
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
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
    if (a.dbgClk && lane == 0) a.dbgClk[16 * blockIdx.x + 0] += cwEmpty;
  } else if (warp == I8_WARP_MMA && rank == 0) {
    // ===================== MMA issuer (one thread; CG 2: of the leader CTA) =====================
    constexpr uint32_t idesc = umma_idesc_i8(128 * CG, 256);
    // descriptor = constant high part (LBO 128 B, SBO 512 B, version 1) | (smem address >> 4)
    constexpr uint64_t descHi = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46);
    const uint32_t stage0 = smem_u32(stages);
    int stage = 0; uint32_t phase = 0; int n = 0;
    long long cwAcc = 0, cwFull = 0;
    const long long cStart = a.dbgClk ? clock64() : 0;
    if (lane == 0)
    while (tiles.next(a, it, jt)) {
      const int b = n & 1;
      long long c0 = a.dbgClk ? clock64() : 0;
      if constexpr (CG == 2) mbar_wait_cluster(smem_u32(&accEmpty[b]), (uint32_t)(((n >> 1) & 1) ^ 1));
      else mbar_wait(smem_u32(&accEmpty[b]), (uint32_t)(((n >> 1) & 1) ^ 1));
      if (a.dbgClk) cwAcc += clock64() - c0;
      tc_fence_after();
      const uint32_t dTmem = tmemBase + (uint32_t)(b * 256);
      for (int c = 0; c < a.nC; c += BPS) {
        const int nb = (a.nC - c < BPS) ? a.nC - c : BPS;
        c0 = a.dbgClk ? clock64() : 0;
        if constexpr (CG == 2) mbar_wait_cluster(smem_u32(&fullBar[stage]), phase);
        else mbar_wait(smem_u32(&fullBar[stage]), phase);
        if (a.dbgClk) cwFull += clock64() - c0;
        tc_fence_after();
        const uint32_t sA = stage0 + (uint32_t)stage * STAGE_BYTES;
        const uint32_t sB = sA + BPS * I8_BLK_BYTES;
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
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit<CG>(smem_u32(&accFull[b]));
      ++n;
    }
    if (a.dbgClk && lane == 0) {
      a.dbgClk[16 * blockIdx.x + 1] += cwAcc; a.dbgClk[16 * blockIdx.x + 2] += cwFull;
      a.dbgClk[16 * blockIdx.x + 3] += clock64() - cStart; a.dbgClk[16 * blockIdx.x + 4] += n;
    }
  } else if (warp == I8_WARP_MMA) {
    // ===================== CG 2, peer CTA: relay "my operand half has landed" to the leader =====================
    if constexpr (CG == 2) {
      int stage = 0; uint32_t phase = 0;
      if (lane == 0)
      while (tiles.next(a, it, jt)) {
        for (int c = 0; c < a.nC; c += BPS) {
          mbar_wait(smem_u32(&fullBar[stage]), phase);
          mbar_arrive_cluster(cluster_map(smem_u32(&fullBar[stage]), 0));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp < I8_DRAIN_WARPS) {
    // ===================== drain warps =====================
    const int sp = warp;              // TMEM sub-partition: lanes 32*sp .. 32*sp+31
    const int r = 32 * sp + lane;     // operand row of this thread: 9*i + 3*p + digit
    const double wdig = (r % 3 == 0) ? 1.0 : ((r % 3 == 1) ? 256.0 : 65536.0);
    const double wmagic = -6755399441055744.0 * wdig;   // -(1.5 * 2^52) * weight
    unsigned char* xrow = xbuf + (size_t)r * I8_XROW_BYTES;
    int n = 0;
    long long cwAcc = 0, cwX = 0;
    const long long cStart = a.dbgClk ? clock64() : 0;
    while (tiles.next(a, it, jt)) {
      const int b = n & 1;
      long long c0 = a.dbgClk ? clock64() : 0;
      mbar_wait(smem_u32(&accFull[b]), (uint32_t)((n >> 1) & 1));
      if (a.dbgClk) cwAcc += clock64() - c0;
      tc_fence_after();
      if (a.dbgMode == 2 || a.dbgMode == 3) {   // timing experiment: MMA + operand pipeline only
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
      for (int g = 0; g < 4; ++g) {
        // column frames 7g .. 7g+6: accumulator columns 9*jl (frames 0..13) or 128 + 9*(jl-14) (frames 14..27)
        const uint32_t tcol = tmemBase + (uint32_t)(b * 256 + (g >> 1) * 128 + (g & 1) * 63) + ((uint32_t)(32 * sp) << 16);
        int v[64];
        tmem_ld32(tcol, v);
        tmem_ld32(tcol + 32, v + 32);
        c0 = a.dbgClk ? clock64() : 0;
        mbar_wait(smem_u32(&xEmpty[g]), (uint32_t)((n & 1) ^ 1));   // the solve warps are done with this group of tile n-1
        if (a.dbgClk) cwX += clock64() - c0;
        tmem_ld_wait();
        if (g == 3) {   // this warp is done reading accumulator buffer b
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) mbar_arrive_cluster(cluster_map(smem_u32(&accEmpty[b]), 0));
            else mbar_arrive(smem_u32(&accEmpty[b]));
          }
        }
        double* dst = reinterpret_cast<double*>(xrow + (size_t)(7 * g) * I8_XJ_STRIDE);
#pragma unroll
        for (int jj = 0; jj < 7; ++jj) {
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            // exact int64 -> FP64 without the (slow, XU-pipe) I2F.F64.S64: for |V| < 2^51 the bit pattern
            // 0x4338000000000000 + V is the double 1.5*2^52 + V; one FMA removes the offset and applies the
            // digit weight of this operand row (both exact).
            const long long sv = 0x4338000000000000LL + (long long)v[9 * jj + 3 * q] +
                                 (long long)v[9 * jj + 3 * q + 1] * 256 + (long long)v[9 * jj + 3 * q + 2] * 65536;
            dst[(size_t)jj * (I8_XJ_STRIDE / 8) + q] = fma(__longlong_as_double(sv), wdig, wmagic);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&xFull[g]));
      }
      ++n;
    }
    if (a.dbgClk && lane == 0 && warp == 0) {
      a.dbgClk[16 * blockIdx.x + 5] += cwAcc; a.dbgClk[16 * blockIdx.x + 6] += cwX;
      a.dbgClk[16 * blockIdx.x + 7] += clock64() - cStart;
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
    const double toInt = 1.0 / a.invScale2;                        // G (A^2) -> integer units (exact power of two)
    const double* src = reinterpret_cast<const double*>(xbuf + (size_t)jl * I8_XJ_STRIDE + (size_t)(9 * il) * I8_XROW_BYTES);
    const bool solveOff = (a.dbgMode == 2 || a.dbgMode == 3);
    int n = 0;
    long long cwX = 0;
    const long long cStart = a.dbgClk ? clock64() : 0;
    while (tiles.next(a, it, jt)) {
      if (solveOff) continue;
      const int i = I8_TILE_I * (CG * it + (int)rank) + il, j = I8_TILE_J * jt + jl;
      const bool valid = laneOn && i < a.nRows && j < a.nCols && i >= a.rowLo && i < a.rowHi && (!TRI || j > i);
      // everything stays in integer units (exact); one scale at the very end
      double e0 = 1.0;
      if (valid) e0 = 0.5 * (__ldg(a.GA + i) + __ldg(a.GB + j)) * toInt;
      const double sInv = 1.0 / e0;
      long long c0 = a.dbgClk ? clock64() : 0;
      mbar_wait(smem_u32(&xFull[2 * h]), (uint32_t)(n & 1));
      mbar_wait(smem_u32(&xFull[2 * h + 1]), (uint32_t)(n & 1));
      if (a.dbgClk) cwX += clock64() - c0;
      double S[9];
      if (a.dbgMode != 6) {
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
          for (int q = 0; q < 3; ++q) S[3 * p + q] = (src[(3 * p) * 3 + q] + src[(3 * p + 1) * 3 + q]) + src[(3 * p + 2) * 3 + q];
      } else {
#pragma unroll
        for (int x = 0; x < 9; ++x) S[x] = 0.0;
      }
      // the loads above have completed (their values are consumed): hand the groups back to the drain warps
      {
        double keep = S[0];
#pragma unroll
        for (int x = 1; x < 9; ++x) keep += S[x];
        asm volatile("" ::"d"(keep) : "memory");
      }
      __syncwarp();
      if (lane == 0) { mbar_arrive(smem_u32(&xEmpty[2 * h])); mbar_arrive(smem_u32(&xEmpty[2 * h + 1])); }
      ++n;
      if (a.dbgMode == 6) continue;
      if (a.dbgS && valid) {
#pragma unroll
        for (int x = 0; x < 9; ++x) a.dbgS[((size_t)i * a.nCols + j) * 9 + x] = S[x];
      }
      float rms;
      if (a.dbgMode == 1) {
        rms = (float)(S[0] + S[4] + S[8]);
      } else {
        double gap;   // (E0 - lambda_max) / E0
        const bool ok = i8_relative_gap(S, sInv, gap);
        if (valid && !ok) gap = i8_relative_gap_slow(S, e0);
        // clamp: src/Frame.cpp:1264-1268; the result is stored as float (Matrix<float>), so the root is taken in
        // float: relative error 1.2e-7, i.e. < 4e-7 A for RMSDs of a few A
        const double r2 = gap * (e0 * outScale);
        rms = (r2 > 0.0) ? sqrtf((float)r2) : 0.f;
      }
      if (valid) {
        size_t idx;
        if (TRI)
          idx = tri_row_start((size_t)a.nCols, (size_t)i) + (size_t)(j - i - 1) - a.outBase;
        else
          idx = (size_t)i * a.ldo + (size_t)j;
        a.out[idx] = rms;
      }
    }
    if (a.dbgClk && lane == 0 && u == 0) {
      a.dbgClk[16 * blockIdx.x + 8] += cwX; a.dbgClk[16 * blockIdx.x + 9] += clock64() - cStart;
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

}  // namespace b200
