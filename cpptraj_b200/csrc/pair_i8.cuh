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
constexpr int I8_STAGES = 5;
constexpr int I8_STAGE_BYTES = 3 * I8_BLK_BYTES;  // A (1 block) + B (2 blocks) = 24 KB
constexpr int I8_TILE_I = I8_FR_PER_RG;      // 14
constexpr int I8_TILE_J = 2 * I8_FR_PER_RG;  // 28
constexpr int I8_EPI_WARPS = 16;            // 4 per TMEM sub-partition: 7 j frames (63 accumulator columns) each
constexpr int I8_EPI_THREADS = I8_EPI_WARPS * 32;
constexpr int I8_THREADS = I8_EPI_THREADS + 64;   // + TMA producer warp + MMA issuer warp
constexpr int I8_XROW_BYTES = 24;            // 3 doubles per (operand row, j frame)
constexpr int I8_XJ_STRIDE = 128 * I8_XROW_BYTES + 88;  // 3160 B: (stride/4) % 32 == 22 -> conflict-free LDS.64
constexpr int I8_XBUF_BYTES = I8_TILE_J * I8_XJ_STRIDE; // 88480
constexpr int I8_SMEM_BYTES = I8_STAGES * I8_STAGE_BYTES + I8_XBUF_BYTES + 256 + 512;   // + barriers + per-tile G
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
__device__ __forceinline__ void tmem_alloc(uint32_t dstSmem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dstSmem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
/// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, issued by one thread.
__device__ __forceinline__ void umma_i8(uint32_t dTmem, uint64_t aDesc, uint64_t bDesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dTmem), "l"(aDesc), "l"(bDesc), "r"(idesc),
      "r"(accumulate)
      : "memory");
}
/// mbarrier arrive once all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(I8_EPI_THREADS) : "memory"); }

// ----------------------------------------------------------------------------
// pair_i8_kernel
// ----------------------------------------------------------------------------
struct PairI8Args {
  const uint8_t* PA;   // operand images of the row frames (i)
  const uint8_t* PB;   // operand images of the column frames (j); even number of row groups allocated
  const double* GA;
  const double* GB;
  int nC;              // 64-atom chunks
  int nRows, nCols;    // valid i / j frames
  int rowLo, rowHi;    // rows written by this launch: [rowLo, rowHi)
  int it0, nIt;        // i tiles (14 frames) of this launch
  int jt0, nJt;        // j tiles (28 frames) of this launch
  const double* totalMass;
  double invScale2;    // 2^-2qs: integer covariance -> A^2
  float* out;          // TRI: out[triIndex - outBase]; FULL: out[i*ldo + j]
  size_t outBase;
  size_t ldo;
  double* dbgS;        // nullable: 9 doubles per (i,j) at (i*nCols + j)*9, integer units
  long long* dbgClk;   // nullable: per-CTA cycle counters [16] (timing experiments)
  int dbgMode;         // 0 normal; timing experiments: 1 no per-pair solve, 2 epilogue only frees TMEM, 3 = 2 + no operand loads
};

template <bool TRI>
__device__ __forceinline__ bool i8_tile(const PairI8Args& a, int t, int& it, int& jt) {
  it = a.it0 + t % a.nIt;
  jt = a.jt0 + t / a.nIt;
  if (I8_TILE_I * it >= a.nRows || I8_TILE_J * jt >= a.nCols) return false;
  if (I8_TILE_I * it >= a.rowHi || I8_TILE_I * it + I8_TILE_I <= a.rowLo) return false;
  if (TRI && I8_TILE_J * jt + I8_TILE_J - 1 <= I8_TILE_I * it) return false;  // every j <= every i
  return true;
}

template <bool TRI>
__global__ void __launch_bounds__(I8_THREADS, 1) pair_i8_kernel(PairI8Args a) {
  extern __shared__ __align__(1024) unsigned char smem_i8[];
  unsigned char* stages = smem_i8;
  unsigned char* xbuf = smem_i8 + I8_STAGES * I8_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xbuf + I8_XBUF_BYTES);
  uint64_t* fullBar = bars;                    // [I8_STAGES]
  uint64_t* emptyBar = bars + I8_STAGES;       // [I8_STAGES]
  uint64_t* accFull = bars + 2 * I8_STAGES;    // [2]
  uint64_t* accEmpty = bars + 2 * I8_STAGES + 2;  // [2]
  uint32_t* tmemBaseSlot = reinterpret_cast<uint32_t*>(bars + 2 * I8_STAGES + 4);
  double* gbuf = reinterpret_cast<double*>(xbuf + I8_XBUF_BYTES + 256);   // [14 GA | 28 GB] of the current tile, integer units

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nTiles = a.nIt * a.nJt;

  if (tid == 0) {
    for (int s = 0; s < I8_STAGES; ++s) { mbar_init(smem_u32(&fullBar[s]), 1); mbar_init(smem_u32(&emptyBar[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&accFull[b]), 1); mbar_init(smem_u32(&accEmpty[b]), I8_EPI_WARPS); }
    mbar_fence_init();
  }
  if (warp == I8_EPI_WARPS + 1) tmem_alloc(smem_u32(tmemBaseSlot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmemBase = *tmemBaseSlot;

  if (warp == I8_EPI_WARPS) {
    // ===================== TMA producer (one thread) =====================
    int stage = 0; uint32_t phase = 0;
    const uint32_t stage0 = smem_u32(stages);
    long long cwEmpty = 0;
    if (lane == 0)
    for (int t = blockIdx.x; t < nTiles; t += gridDim.x) {
      int it, jt;
      if (!i8_tile<TRI>(a, t, it, jt)) continue;
      const uint8_t* gA = a.PA + (size_t)it * a.nC * I8_BLK_BYTES;
      const uint8_t* gB0 = a.PB + (size_t)(2 * jt) * a.nC * I8_BLK_BYTES;
      const uint8_t* gB1 = gB0 + (size_t)a.nC * I8_BLK_BYTES;
      for (int c = 0; c < a.nC; ++c) {
        const long long c0 = a.dbgClk ? clock64() : 0;
        mbar_wait(smem_u32(&emptyBar[stage]), phase ^ 1u);
        if (a.dbgClk) cwEmpty += clock64() - c0;
        {
          const uint32_t bar = smem_u32(&fullBar[stage]);
          const uint32_t dst = stage0 + (uint32_t)stage * I8_STAGE_BYTES;
          if (a.dbgMode == 3) {
            mbar_arrive(bar);
          } else {
            mbar_expect_tx(bar, I8_STAGE_BYTES);
            bulk_g2s(dst, gA + (size_t)c * I8_BLK_BYTES, I8_BLK_BYTES, bar);
            bulk_g2s(dst + I8_BLK_BYTES, gB0 + (size_t)c * I8_BLK_BYTES, I8_BLK_BYTES, bar);
            bulk_g2s(dst + 2 * I8_BLK_BYTES, gB1 + (size_t)c * I8_BLK_BYTES, I8_BLK_BYTES, bar);
          }
        }
        if (++stage == I8_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
    if (a.dbgClk && lane == 0) a.dbgClk[16 * blockIdx.x + 0] += cwEmpty;
  } else if (warp == I8_EPI_WARPS + 1) {
    // ===================== MMA issuer (one thread) =====================
    constexpr uint32_t idesc = umma_idesc_i8(128, 256);
    // descriptor = constant high part (LBO 128 B, SBO 512 B, version 1) | (smem address >> 4)
    constexpr uint64_t descHi = ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46);
    const uint32_t stage0 = smem_u32(stages);
    int stage = 0; uint32_t phase = 0; int n = 0;
    long long cwAcc = 0, cwFull = 0, cTot = 0;
    const long long cStart = a.dbgClk ? clock64() : 0;
    if (lane == 0)
    for (int t = blockIdx.x; t < nTiles; t += gridDim.x) {
      int it, jt;
      if (!i8_tile<TRI>(a, t, it, jt)) continue;
      const int b = n & 1;
      long long c0 = a.dbgClk ? clock64() : 0;
      mbar_wait(smem_u32(&accEmpty[b]), (uint32_t)(((n >> 1) & 1) ^ 1));
      if (a.dbgClk) cwAcc += clock64() - c0;
      tc_fence_after();
      const uint32_t dTmem = tmemBase + (uint32_t)(b * 256);
      for (int c = 0; c < a.nC; ++c) {
        c0 = a.dbgClk ? clock64() : 0;
        mbar_wait(smem_u32(&fullBar[stage]), phase);
        if (a.dbgClk) cwFull += clock64() - c0;
        tc_fence_after();
        const uint32_t sA = stage0 + (uint32_t)stage * I8_STAGE_BYTES;
        const uint64_t dA = descHi | (uint64_t)((sA >> 4) & 0x3fff);
        const uint64_t dB = descHi | (uint64_t)(((sA + I8_BLK_BYTES) >> 4) & 0x3fff);
#pragma unroll
        for (int k = 0; k < I8_KC / 32; ++k)   // one K=32 step = two 16-byte core matrices = 256 B further on
          if (a.dbgMode != 5) umma_i8(dTmem, dA + (uint64_t)(k * 16), dB + (uint64_t)(k * 16), idesc, (uint32_t)((c | k) != 0));
        umma_commit(smem_u32(&emptyBar[stage]));
        if (++stage == I8_STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(smem_u32(&accFull[b]));
      ++n;
    }
    if (a.dbgClk && lane == 0) {
      cTot = clock64() - cStart;
      a.dbgClk[16 * blockIdx.x + 1] += cwAcc; a.dbgClk[16 * blockIdx.x + 2] += cwFull;
      a.dbgClk[16 * blockIdx.x + 3] += cTot; a.dbgClk[16 * blockIdx.x + 4] += n;
    }
  } else {
    // ===================== epilogue warps =====================
    // Latency-bound work (TMEM loads, int64 -> FP64 conversions, Newton iterations): 16 warps keep
    // four independent instruction streams per scheduler in flight.
    const int sp = warp & 3;          // TMEM sub-partition: lanes 32*sp .. 32*sp+31
    const int cg = warp >> 2;         // column group: j frames 7*cg .. 7*cg+6 of the tile
    const int r = 32 * sp + lane;     // operand row of this thread: 9*i + 3*p + digit
    const double wdig = (r % 3 == 0) ? 1.0 : ((r % 3 == 1) ? 256.0 : 65536.0);
    const double wmagic = -6755399441055744.0 * wdig;   // -(1.5 * 2^52) * weight
    const double outScale = 2.0 * a.invScale2 / a.totalMass[0];   // rmsd^2 = (E0 - lambda) * outScale (integer units)
    const double toInt = 1.0 / a.invScale2;                        // G (A^2) -> integer units (exact power of two)
    // first accumulator column of frame 7*cg: frames 0..13 start at 9*jl, frames 14..27 at 128 + 9*(jl-14)
    const int col0 = (cg >> 1) * 128 + (cg & 1) * 63;
    int n = 0;
    long long cwFullAcc = 0, cRow = 0, cBar1 = 0, cPair = 0, cBar2 = 0;
    for (int t = blockIdx.x; t < nTiles; t += gridDim.x) {
      int it, jt;
      if (!i8_tile<TRI>(a, t, it, jt)) continue;
      const int b = n & 1;
      long long c0 = a.dbgClk ? clock64() : 0, c1;
      mbar_wait(smem_u32(&accFull[b]), (uint32_t)((n >> 1) & 1));
      if (a.dbgClk) { c1 = clock64(); cwFullAcc += c1 - c0; c0 = c1; }
      tc_fence_after();
      if (a.dbgMode >= 2 && a.dbgMode <= 3) {   // timing experiment: MMA + operand pipeline only
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&accEmpty[b]));
        ++n;
        continue;
      }
      // ---- row phase: digits of B folded in-thread, result to the exchange buffer ----
      {
        double gval = 0.0;
        if (tid < I8_TILE_I) {
          const int i = I8_TILE_I * it + tid;
          if (i < a.nRows) gval = a.GA[i] * toInt;
        } else if (tid < I8_TILE_I + I8_TILE_J) {
          const int j = I8_TILE_J * jt + tid - I8_TILE_I;
          if (j < a.nCols) gval = a.GB[j] * toInt;
        }
        int v[64];
        const uint32_t tcol = tmemBase + (uint32_t)(b * 256 + col0) + ((uint32_t)(32 * sp) << 16);
        tmem_ld32(tcol, v);
        tmem_ld32(tcol + 32, v + 32);
        tmem_ld_wait();
        // this warp is done reading accumulator buffer b
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&accEmpty[b]));
        double* dst = reinterpret_cast<double*>(xbuf + (size_t)(7 * cg) * I8_XJ_STRIDE + (size_t)r * I8_XROW_BYTES);
        if (a.dbgMode != 7)
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
        if (tid < I8_TILE_I + I8_TILE_J) gbuf[tid] = gval;
      }
      if (a.dbgClk) { c1 = clock64(); cRow += c1 - c0; c0 = c1; }
      epi_bar_sync();   // exchange buffer complete
      if (a.dbgClk) { c1 = clock64(); cBar1 += c1 - c0; c0 = c1; }
      // ---- pair phase: one thread per frame pair ----
      for (int e = tid; e < I8_TILE_I * I8_TILE_J && a.dbgMode != 6; e += I8_EPI_THREADS) {
        const int il = e / I8_TILE_J, jl = e % I8_TILE_J;
        const int i = I8_TILE_I * it + il, j = I8_TILE_J * jt + jl;
        if (i >= a.nRows || j >= a.nCols || i < a.rowLo || i >= a.rowHi) continue;
        if (TRI && j <= i) continue;
        const double* src = reinterpret_cast<const double*>(xbuf + (size_t)jl * I8_XJ_STRIDE +
                                                            (size_t)(9 * il) * I8_XROW_BYTES);
        double S[9];
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
          for (int q = 0; q < 3; ++q) S[3 * p + q] = src[(3 * p) * 3 + q] + src[(3 * p + 1) * 3 + q] + src[(3 * p + 2) * 3 + q];
        if (a.dbgS) {
#pragma unroll
          for (int x = 0; x < 9; ++x) a.dbgS[((size_t)i * a.nCols + j) * 9 + x] = S[x];
        }
        // everything stays in integer units (exact); one scale at the very end
        const double e0 = 0.5 * (gbuf[il] + gbuf[I8_TILE_I + jl]);
        float rms;
        if (a.dbgMode == 1) {
          rms = (float)(S[0] + S[4] + S[8]);
        } else {
          double gap;   // (E0 - lambda_max) / E0
          if (!relative_gap_fast(S, e0, gap)) gap = (e0 > 0.0) ? (e0 - largest_root(quartic_of(S), e0, S)) / e0 : 0.0;
          // clamp: src/Frame.cpp:1264-1268; the result is stored as float (Matrix<float>), so the root is taken in
          // float: relative error 1.2e-7, i.e. < 4e-7 A for RMSDs of a few A
          const double r2 = gap * e0 * outScale;
          rms = (r2 > 0.0) ? sqrtf((float)r2) : 0.f;
        }
        size_t idx;
        if (TRI)
          idx = tri_row_start((size_t)a.nCols, (size_t)i) + (size_t)(j - i - 1) - a.outBase;
        else
          idx = (size_t)i * a.ldo + (size_t)j;
        a.out[idx] = rms;
      }
      if (a.dbgClk) { c1 = clock64(); cPair += c1 - c0; c0 = c1; }
      epi_bar_sync();   // exchange buffer free again
      if (a.dbgClk) { c1 = clock64(); cBar2 += c1 - c0; c0 = c1; }
      ++n;
    }
    if (a.dbgClk && (tid == 0 || tid == 64)) {
      long long* o = a.dbgClk + 16 * blockIdx.x + (tid == 0 ? 5 : 10);
      o[0] += cwFullAcc; o[1] += cRow; o[2] += cBar1; o[3] += cPair; o[4] += cBar2;
    }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == I8_EPI_WARPS + 1) tmem_dealloc(tmemBase, 512);
}


// ----------------------------------------------------------------------------
// tcgen05 kind::i8 issue-peak probe: one CTA per SM, operands fixed in shared memory (no loads),
// one thread issues M128 x N256 x K32 MMAs back to back.  Roofline denominator of pair_i8_kernel.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) i8_mma_peak_kernel(int iters, int* sink) {
  extern __shared__ __align__(1024) unsigned char smem_pk[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmemSlot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (3 * I8_BLK_BYTES) / 4; i += 128) reinterpret_cast<uint32_t*>(smem_pk)[i] = 0x01010101u * (uint32_t)(i & 3);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmemSlot), 512);
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
      umma_i8(tmemBase + (uint32_t)((it & 1) * 256), umma_desc(sA + (it & 1) * 256, 128, 512),
              umma_desc(sB + (it & 1) * 256, 128, 512), idesc, (uint32_t)(it > 1));
    }
    umma_commit(smem_u32(&bar));
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
  if (warp == 0) tmem_dealloc(tmemBase, 512);
}

}  // namespace b200
