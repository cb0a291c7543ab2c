// hieragglo.cuh -- hierarchical agglomerative clustering on a device-resident pairwise-distance triangle.
//
// Replaces the merge loop of Algorithm_HierAgglo::DoClustering / MergeClosest
// (src/Cluster/Algorithm_HierAgglo.cpp:97-245) together with the Cluster::DynamicMatrix bookkeeping
// (src/Cluster/DynamicMatrix.h:43-126, DynamicMatrix.cpp:7-33).  The reference recomputes the linkage of the merged
// cluster from frame pairs (O(|C1|*N) cache reads per merge, :248-350); here the cluster-distance triangle is updated
// in place by the equivalent recurrences
//     single   : d(C1+C2, k) = min(d(C1,k), d(C2,k))          (exact: a minimum of the same floats)
//     complete : d(C1+C2, k) = max(d(C1,k), d(C2,k))          (exact)
//     average  : S(C1+C2, k) = S(C1,k) + S(C2,k),  d = float(S / double(n1*nk))   (S = sum of the float distances in
//                double; identical to the reference's frame-pair sum while the partial sums are exactly representable,
//                i.e. up to ~2^29/range frame pairs per cluster pair, equal to ~1e-16 relative beyond that)
// so a merge costs O(N).  Everything that decides WHICH pair merges is replicated exactly: FindMin takes the lowest
// column among equal minima, a cluster's closest index changes only on a strictly smaller distance, is re-scanned
// (lowest index among equal minima) when its closest distance grew or its closest cluster was merged away, and the
// history-dependent tie case of the merged cluster's own closest index is replayed sequentially.  A re-scan (O(N)) is
// skipped when its outcome is certain: every cluster keeps a lower bound lb2 of its distances to all clusters other
// than its closest one (exact after a scan, only ever lowered in between); a new distance to the merged cluster that
// is strictly below lb2 makes the merged cluster the closest whatever the scan would have found.
//
// Layout: the cluster distances live in a full symmetric n x n matrix (expanded once from the cache triangle), so the
// two rows a merge combines and every row that has to be re-scanned are contiguous; an update writes the row and
// (scattered) the column of the merged cluster.
// Execution: ONE thread-block cluster (1..16 CTAs of 1024 threads) runs all merges in a single launch; the phases of a
// merge are separated by cluster barriers (hardware barrier + acquire/release at cluster scope) instead of kernel
// launches or grid-wide atomics: 3 barriers per merge in the common case.  Per merge each thread handles a strided slice
// of the clusters k: two gathers from the triangle (rows C1 and C2), one store.
#pragma once
#include <cooperative_groups.h>
#include <cstdint>

namespace b200 {
namespace cg = cooperative_groups;

typedef unsigned long long ha_u64;

struct HaCtl {
  ha_u64 minKey[2];   // FindMin: (distance, column), by merge parity
  ha_u64 rowKey[2];   // minimum of the merged cluster's new row: (distance, lowest k)
  int tieCount[2];    // number of k attaining that minimum
  int nA[2], nB[2];   // list lengths: rows to re-scan before / after the update
  int nCalls, nMerges, stopped;
  long long clk[8];   // leader thread: cycles per phase (findmin, newrow, ignore-rescan, update, post-rescan), summed
  long long cnt[4];   // rows re-scanned before / after the update, merges whose row minimum was tied
};

struct HaArgs {
  float* D;            // cluster distances, symmetric n x n (row-major; the diagonal is unused)
  double* S;           // average linkage: sums of frame-pair distances (same layout), else null
  int n, linkage, target;
  double eps;
  int* closest; float* cmin;        // per cluster: closest cluster and D(i, closest[i])
  unsigned char* ign; int* nfr;     // merged-away flag, frame count
  float* vnew; double* snew; float* oold;   // the merged cluster's new row, its sums, its old row
  int* listA; int* listB;
  ha_u64* rkey;        // per cluster: (distance, index) minimum of a pending re-scan, ~0 when none
  unsigned int* rlb;   // ... and the second smallest distance of that re-scan (ordered bits), 0xffffffff when none
  float* lb2;          // per cluster: a lower bound of its distances to every cluster other than closest[]
  HaCtl* ctl;
  int* mergeInto; int* mergeFrom; float* findMin;
};

__device__ __forceinline__ unsigned int ha_ord(float v) {
  unsigned int b = __float_as_uint(v + 0.0f);   // (-0 -> +0)
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ha_unord(unsigned int u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ ha_u64 ha_key(float v, int idx) { return ((ha_u64)ha_ord(v) << 32) | (unsigned int)idx; }
__device__ __forceinline__ ha_u64 ha_warp_min(ha_u64 k) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    ha_u64 t = __shfl_xor_sync(0xffffffffu, k, o);
    k = t < k ? t : k;
  }
  return k;
}
/// Minimum over the CTA; result in every thread.  sm: 33 entries.
__device__ __forceinline__ ha_u64 ha_block_min(ha_u64 k, ha_u64* sm) {
  k = ha_warp_min(k);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = k;
  __syncthreads();
  if (threadIdx.x < 32) {
    ha_u64 t = (threadIdx.x < (blockDim.x >> 5)) ? sm[threadIdx.x] : ~0ull;
    t = ha_warp_min(t);
    if (threadIdx.x == 0) sm[32] = t;
  }
  __syncthreads();
  return sm[32];
}

constexpr int HA_U = 4;   // clusters per thread and batch: all loads of a batch are issued before the first use

/// Smallest (distance, index) key and second smallest distance (ordered bits; 0xffffffff = none) of a set of entries.
struct HaMin2 { ha_u64 best; unsigned int sec; };
__device__ __forceinline__ void ha_m2_add(HaMin2& m, ha_u64 key) {
  const ha_u64 hi = key < m.best ? m.best : key;
  m.best = key < m.best ? key : m.best;
  m.sec = min(m.sec, (unsigned int)(hi >> 32));
}
__device__ __forceinline__ HaMin2 ha_m2_merge(const HaMin2& x, const HaMin2& y) {
  HaMin2 r;
  const ha_u64 hi = x.best < y.best ? y.best : x.best;
  r.best = x.best < y.best ? x.best : y.best;
  r.sec = min(min(x.sec, y.sec), (unsigned int)(hi >> 32));
  return r;
}
__device__ __forceinline__ HaMin2 ha_m2_warp(HaMin2 m) {
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    HaMin2 t;
    t.best = __shfl_xor_sync(0xffffffffu, m.best, o);
    t.sec = __shfl_xor_sync(0xffffffffu, m.sec, o);
    m = ha_m2_merge(m, t);
  }
  return m;
}
__device__ __forceinline__ float ha_lb_of(unsigned int sec) { return sec == 0xffffffffu ? __int_as_float(0x7f800000) : ha_unord(sec); }

#ifndef B200_HA_RESCAN_U
#define B200_HA_RESCAN_U 8    // re-scans of whole rows: loads in flight per thread (a row of 50,000 clusters is 250 KB: latency, not bandwidth; 12 and 16 spill)
#endif
constexpr int HA_UR = B200_HA_RESCAN_U;

/// Minimum and second minimum of this thread's share of one row (clusters still present; the row is contiguous): elements
/// first, first + step, ...; U loads of the row and of the merged-away flags in flight.
template <int U>
__device__ __forceinline__ HaMin2 ha_row_part_min(const HaArgs& a, int row, int first, int step) {
  const int n = a.n;
  const float* R = a.D + (size_t)row * n;
  HaMin2 m; m.best = ~0ull; m.sec = 0xffffffffu;
  for (int j0 = first; j0 < n; j0 += step * U) {
    float v[U]; unsigned char ig[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = j0 + step * u;
      const bool in = j < n && j != row;
      ig[u] = in ? a.ign[j] : (unsigned char)1;
      v[u] = in ? R[j] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (!ig[u]) ha_m2_add(m, ha_key(v[u], j0 + step * u));
  }
  return m;
}
/// One row's minimum and second minimum by one warp.
template <int U = HA_U>
__device__ __forceinline__ HaMin2 ha_warp_row_min(const HaArgs& a, int row, int lane) {
  return ha_m2_warp(ha_row_part_min<U>(a, row, lane, 32));
}

/// DynamicMatrix::updateClosestIdx (DynamicMatrix.h:43-62) for every row of a list (lowest index among the minima over
/// the clusters still present); the results go to rkey[row] / rlb[row], to be read after the next cluster barrier.  A
/// short list: the whole team scans each row (one round of loads per row).  A long one (a big cluster was merged away
/// and was the closest of many): one warp per row, all warps of the team side by side.
__device__ __forceinline__ void ha_team_rescan(const HaArgs& a, const int* list, int nList, int tid, int nThr, ha_u64* sm) {
  const int n = a.n;
  const int nCta = nThr / (int)blockDim.x, cta = tid / (int)blockDim.x;
  unsigned int* smSec = reinterpret_cast<unsigned int*>(sm + 34);
  if (nList > 25 * nCta) {
    // very long lists: one warp per row, all warps of the team side by side
    const int lane = threadIdx.x & 31, nWarps = nThr >> 5;
    for (int e = tid >> 5; e < nList; e += nWarps) {
      const int row = list[e];
      const HaMin2 m = ha_warp_row_min<HA_UR>(a, row, lane);
      if (lane == 0) { a.rkey[row] = m.best; a.rlb[row] = m.sec; }
    }
    return;
  }
  if (nList > 1 && nCta > 1) {
    // a handful to a few hundred rows: one CTA per row, the CTAs of the team side by side (a row costs a few dependent
    // rounds of loads whoever scans it: what counts is how many rows are scanned at the same time)
    for (int e = cta; e < nList; e += nCta) {
      const int row = list[e];
      HaMin2 m = ha_m2_warp(ha_row_part_min<HA_UR>(a, row, (int)threadIdx.x, (int)blockDim.x));
      __syncthreads();
      if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m.best; smSec[threadIdx.x >> 5] = m.sec; }
      __syncthreads();
      if (threadIdx.x < 32) {
        HaMin2 t;
        const bool have = threadIdx.x < (blockDim.x >> 5);
        t.best = have ? sm[threadIdx.x] : ~0ull; t.sec = have ? smSec[threadIdx.x] : 0xffffffffu;
        t = ha_m2_warp(t);
        if (threadIdx.x == 0) { a.rkey[row] = t.best; a.rlb[row] = t.sec; }
      }
    }
    return;
  }
  for (int e = 0; e < nList; ++e) {
    const int row = list[e];
    const float* R = a.D + (size_t)row * n;
    HaMin2 m; m.best = ~0ull; m.sec = 0xffffffffu;
    for (int j0 = tid; j0 < n; j0 += nThr * HA_U) {
      float v[HA_U]; unsigned char ig[HA_U];
#pragma unroll
      for (int u = 0; u < HA_U; ++u) {
        const int j = j0 + u * nThr;
        const bool in = j < n && j != row;
        ig[u] = in ? a.ign[j] : (unsigned char)1;
        v[u] = in ? R[j] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < HA_U; ++u)
        if (!ig[u]) ha_m2_add(m, ha_key(v[u], j0 + u * nThr));
    }
    m = ha_m2_warp(m);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m.best; smSec[threadIdx.x >> 5] = m.sec; }
    __syncthreads();
    if (threadIdx.x < 32) {
      HaMin2 t;
      const bool have = threadIdx.x < (blockDim.x >> 5);
      t.best = have ? sm[threadIdx.x] : ~0ull; t.sec = have ? smSec[threadIdx.x] : 0xffffffffu;
      t = ha_m2_warp(t);
      if (threadIdx.x == 0 && t.best != ~0ull) {
        // every key that does not end up as the row's minimum passes through here as a candidate for the second minimum
        const ha_u64 old = atomicMin(&a.rkey[row], t.best);
        const ha_u64 loser = old < t.best ? t.best : old;
        atomicMin(&a.rlb[row], min(t.sec, (unsigned int)(loser >> 32)));
      }
    }
  }
}
/// Takes a re-scan result out of rkey[row] / rlb[row] (and re-arms the slots).
__device__ __forceinline__ void ha_take_rescan(const HaArgs& a, int row, int& closest, float& cmin, float& lb) {
  const ha_u64 k = a.rkey[row];
  const unsigned int sec = a.rlb[row];
  a.rkey[row] = ~0ull; a.rlb[row] = 0xffffffffu;
  closest = (k == ~0ull) ? -1 : (int)(unsigned int)k;
  cmin = (k == ~0ull) ? __int_as_float(0x7f800000) : ha_unord((unsigned int)(k >> 32));
  lb = ha_lb_of(sec);
  a.closest[row] = closest; a.cmin[row] = cmin; a.lb2[row] = lb;
}

/// The col-side sequence of SetCdist(C1, k, v_k), k ascending (DynamicMatrix.h:65-113), replayed by one warp: needed only
/// when the minimum of the new row is attained more than once (otherwise the result is that unique minimum).
__device__ __forceinline__ void ha_tie_replay(const HaArgs& a, int C1, int C2) {
  const int lane = threadIdx.x & 31;
  const int n = a.n;
  int c = a.closest[C1];
  if (c == C2) {   // Ignore(C2) re-scanned C1's OLD row (DynamicMatrix.h:116-126)
    ha_u64 best = ~0ull;
    for (int j = lane; j < n; j += 32)
      if (j != C1 && !a.ign[j]) { ha_u64 k = ha_key(a.oold[j], j); best = k < best ? k : best; }
    best = ha_warp_min(best);
    c = (best == ~0ull) ? -1 : (int)(unsigned int)best;
  }
  if (c < 0) { if (lane == 0) { a.closest[C1] = -1; a.cmin[C1] = __int_as_float(0x7f800000); } return; }
  float vc = a.vnew[c], oc = a.oold[c];
  constexpr int TR_U = 4;   // 32-cluster groups whose loads are issued together (they do not depend on the replay's state)
  for (int base0 = 0; base0 < n; base0 += 32 * TR_U) {
    bool validU[TR_U]; float vU[TR_U], oU[TR_U];
#pragma unroll
    for (int u = 0; u < TR_U; ++u) {
      const int k = base0 + 32 * u + lane;
      const bool in = k < n && k != C1;
      const unsigned char ig = in ? a.ign[k] : (unsigned char)1;
      vU[u] = in ? a.vnew[k] : 0.f;
      oU[u] = in ? a.oold[k] : 0.f;
      validU[u] = !ig;
    }
#pragma unroll
    for (int u = 0; u < TR_U; ++u) {
    const int base = base0 + 32 * u;
    if (base >= n) break;
    const int k = base + lane;
    const bool valid = validU[u];
    const float v = valid ? vU[u] : 0.f;
    const float o = valid ? oU[u] : 0.f;
    int start = 0;
    for (;;) {
      const float cur = (c < k) ? vc : oc;   // element c of the half-updated row as SetCdist(C1,k) sees it
      const bool ev = valid && lane >= start && (v < cur || (k == c && v > o));
      const unsigned int mask = __ballot_sync(0xffffffffu, ev);
      if (!mask) break;
      const int L = __ffs(mask) - 1;
      const int kk = base + L;
      const float vL = __shfl_sync(0xffffffffu, v, L);
      const float curL = __shfl_sync(0xffffffffu, cur, L);
      if (vL < curL) c = kk;
      else {   // the closest distance grew: re-scan the row as it stands (new up to kk, old beyond)
        ha_u64 best = ~0ull;
        for (int j = lane; j < n; j += 32)
          if (j != C1 && !a.ign[j]) {
            ha_u64 key = ha_key(j <= kk ? a.vnew[j] : a.oold[j], j);
            best = key < best ? key : best;
          }
        best = ha_warp_min(best);
        c = (int)(unsigned int)best;
      }
      vc = a.vnew[c]; oc = a.oold[c];
      start = L + 1;
      if (start >= 32) break;
    }
    }
  }
  if (lane == 0) { a.closest[C1] = c; a.cmin[C1] = a.vnew[c]; a.lb2[C1] = a.vnew[c]; }
}

/// All merges in one launch.  Launch with ONE cluster of `team` CTAs x 1024 threads.
template <int LINK>
__global__ void __launch_bounds__(1024, 1) hieragglo_kernel(HaArgs a) {
  cg::cluster_group team = cg::this_cluster();
  __shared__ ha_u64 sm[34 + 16];   // block reductions (+ 32 second minima)
  const int nCta = (int)team.num_blocks();
  const int cta = (int)team.block_rank();
  const int tid = cta * blockDim.x + threadIdx.x;
  const int nThr = nCta * blockDim.x;
  const bool leader = (tid == 0);
  const int n = a.n;
  volatile HaCtl* ctl = a.ctl;
  int nClusters = n;
  int prevC1 = -1, prevK = -1; float prevV = 0.f;   // previous merge: C1's closest when the row minimum was unique

  long long t0 = clock64();
#define HA_MARK(i) do { if (leader) { long long t1 = clock64(); a.ctl->clk[i] += t1 - t0; t0 = t1; } } while (0)
  for (int m = 0;; ++m) {
    const int p = m & 1;
    // ---- FindMin (DynamicMatrix.cpp:7-33): lowest column among equal minima.  Rows re-scanned after the previous
    //      update (closest == -2) take their result first; so does the previous C1 when its new row had one minimum.
    ha_u64 best = ~0ull;
    for (int c0 = tid; c0 < n; c0 += nThr * HA_U) {
      unsigned char ig[HA_U]; float cm[HA_U]; int cl[HA_U];
#pragma unroll
      for (int u = 0; u < HA_U; ++u) {
        const int col = c0 + u * nThr;
        const bool in = col < n;
        ig[u] = in ? a.ign[col] : (unsigned char)1;
        cm[u] = in ? a.cmin[col] : 0.f;
        cl[u] = in ? a.closest[col] : -1;
      }
#pragma unroll
      for (int u = 0; u < HA_U; ++u) {
        const int col = c0 + u * nThr;
        if (ig[u]) continue;
        if (col == prevC1) { cm[u] = prevV; cl[u] = prevK; a.closest[col] = prevK; a.cmin[col] = prevV; a.lb2[col] = prevV; }
        else if (cl[u] == -2) { float lbTmp; ha_take_rescan(a, col, cl[u], cm[u], lbTmp); }
        if (cl[u] < 0) continue;
        ha_u64 k = ha_key(cm[u], col);
        best = k < best ? k : best;
      }
    }
    best = ha_block_min(best, sm);
    if (threadIdx.x == 0 && best != ~0ull) atomicMin((ha_u64*)&a.ctl->minKey[p], best);
    team.sync();
    HA_MARK(0);
    const ha_u64 mk = ctl->minKey[p];
    const float minVal = ha_unord((unsigned int)(mk >> 32));
    const int colMin = (int)(unsigned int)mk;
    const int rowMin = a.closest[colMin];
    const int C1 = colMin < rowMin ? colMin : rowMin;
    const int C2 = colMin < rowMin ? rowMin : colMin;
    if (leader) { a.findMin[m] = minVal; ctl->nCalls = m + 1; }
    if ((double)minVal > a.eps) { if (leader) ctl->stopped = 1; break; }
    const int n1 = a.nfr[C1] + a.nfr[C2];
    if (leader) {
      a.mergeInto[m] = C1; a.mergeFrom[m] = C2; ctl->nMerges = m + 1;
      a.ign[C2] = 1;
      ctl->minKey[p ^ 1] = ~0ull; ctl->rowKey[p ^ 1] = ~0ull; ctl->tieCount[p ^ 1] = 0; ctl->nA[p ^ 1] = 0; ctl->nB[p ^ 1] = 0;
    }
    --nClusters;
    // ---- new row of the merged cluster; clusters whose closest was C2 are queued for Ignore()'s re-scan
    best = ~0ull;
    for (int k0 = tid; k0 < n; k0 += nThr * HA_U) {
      unsigned char ig[HA_U]; float o1[HA_U], o2[HA_U], lb[HA_U]; double s1[HA_U], s2[HA_U]; int cl[HA_U], nf[HA_U];
#pragma unroll
      for (int u = 0; u < HA_U; ++u) {
        const int k = k0 + u * nThr;
        const bool in = k < n && k != C1 && k != C2;
        ig[u] = in ? a.ign[k] : (unsigned char)1;
        cl[u] = in ? a.closest[k] : -1;
        lb[u] = in ? a.lb2[k] : 0.f;
        const size_t i1 = in ? (size_t)C1 * n + k : 0, i2 = in ? (size_t)C2 * n + k : 0;
        o1[u] = a.D[i1]; o2[u] = a.D[i2];
        if (LINK == 1) { s1[u] = a.S[i1]; s2[u] = a.S[i2]; nf[u] = in ? a.nfr[k] : 1; }
      }
#pragma unroll
      for (int u = 0; u < HA_U; ++u) {
        const int k = k0 + u * nThr;
        if (ig[u]) continue;
        float v;
        if (LINK == 0) v = o2[u] < o1[u] ? o2[u] : o1[u];
        else if (LINK == 2) v = o2[u] > o1[u] ? o2[u] : o1[u];
        else {
          const double sum = s1[u] + s2[u];
          a.snew[k] = sum;
          v = (float)(sum / (double)(n1 * nf[u]));
        }
        a.vnew[k] = v; a.oold[k] = o1[u];
        if (cl[u] == C2) {
          // Ignore(C2) would re-scan this row and SetCdist(C1, k, v) compare v with what it found: below the bound on
          // every other distance the merged cluster is the closest whatever the scan finds
          if (v < lb[u]) { a.closest[k] = C1; a.cmin[k] = v; }
          else a.listA[atomicAdd((int*)&a.ctl->nA[p], 1)] = k;
        }
        ha_u64 key = ha_key(v, k);
        best = key < best ? key : best;
      }
    }
    best = ha_block_min(best, sm);
    if (threadIdx.x == 0 && best != ~0ull) atomicMin((ha_u64*)&a.ctl->rowKey[p], best);
    team.sync();
    HA_MARK(1);
    // ---- Ignore(C2) (DynamicMatrix.h:116-126): re-scan, on the OLD matrix, every cluster whose closest was C2
    const int nA = ctl->nA[p];
    if (nA > 0) {
      ha_team_rescan(a, a.listA, nA, tid, nThr, sm);
      team.sync();
      if (leader) a.ctl->cnt[0] += nA;
    }
    HA_MARK(2);
    // ---- SetCdist(C1, k, v_k), row side (DynamicMatrix.h:88-101), and the update itself
    const ha_u64 rk = ctl->rowKey[p];
    const unsigned int vminOrd = (unsigned int)(rk >> 32);
    for (int k0 = tid; k0 < n; k0 += nThr * HA_U) {
      unsigned char ig[HA_U]; float v[HA_U], cd[HA_U], lb[HA_U]; int ck[HA_U]; double sn[HA_U];
#pragma unroll
      for (int u = 0; u < HA_U; ++u) {
        const int k = k0 + u * nThr;
        const bool in = k < n && k != C1 && k != C2;
        ig[u] = in ? a.ign[k] : (unsigned char)1;
        v[u] = in ? a.vnew[k] : 0.f;
        cd[u] = in ? a.cmin[k] : 0.f;
        lb[u] = in ? a.lb2[k] : 0.f;
        ck[u] = in ? a.closest[k] : -1;
        if (LINK == 1) sn[u] = in ? a.snew[k] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < HA_U; ++u) {
        const int k = k0 + u * nThr;
        if (ig[u]) continue;
        if (ck[u] == C2) ha_take_rescan(a, k, ck[u], cd[u], lb[u]);   // Ignore()'s re-scan of this row
        if (ck[u] < 0 || v[u] < cd[u]) {
          if (ck[u] >= 0 && ck[u] != C1) a.lb2[k] = fminf(lb[u], cd[u]);   // the old closest is now one of the others
          a.closest[k] = C1; a.cmin[k] = v[u];
        } else if (ck[u] == C1) {
          if (v[u] > cd[u]) {   // its closest distance grew
            if (v[u] < lb[u]) a.cmin[k] = v[u];   // still strictly below every other distance: no re-scan
            else { a.closest[k] = -2; a.listB[atomicAdd((int*)&a.ctl->nB[p], 1)] = k; }
          }
        } else if (v[u] < lb[u]) a.lb2[k] = v[u];   // a distance to a cluster other than the closest one
        const size_t i1 = (size_t)C1 * n + k, i2 = (size_t)k * n + C1;
        a.D[i1] = v[u]; a.D[i2] = v[u];
        if (LINK == 1) { a.S[i1] = sn[u]; a.S[i2] = sn[u]; }
        if (ha_ord(v[u]) == vminOrd) atomicAdd((int*)&a.ctl->tieCount[p], 1);
      }
    }
    if (leader) a.nfr[C1] = n1;
    team.sync();
    HA_MARK(3);
    // ---- re-scans after the update: rows whose closest distance (to C1) grew; C1's own closest
    const int nB = ctl->nB[p];
    const int ties = ctl->tieCount[p];
    if (nB > 0 || ties > 1) {
      ha_team_rescan(a, a.listB, nB, tid, nThr, sm);
      if (ties > 1 && cta == 0 && threadIdx.x < 32) ha_tie_replay(a, C1, C2);
      team.sync();
      if (leader) { a.ctl->cnt[1] += nB; a.ctl->cnt[2] += (ties > 1); }
    }
    HA_MARK(4);
    if (ties == 1) { prevC1 = C1; prevK = (int)(unsigned int)rk; prevV = ha_unord(vminOrd); }
    else prevC1 = -1;
    if (nClusters <= a.target || nClusters == 1) break;
  }
}

/// The symmetric n x n matrix from the cache triangle (src/Matrix.h:110-122): block i copies row i of the triangle
/// (contiguous) to D[i][i+1..] and, scattered, to D[i+1..][i].
__global__ void __launch_bounds__(256) hieragglo_expand_kernel(const float* __restrict__ tri, int n, float* __restrict__ D,
                                                               double* __restrict__ S) {
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const size_t base = (size_t)n * (size_t)i - ((size_t)i * ((size_t)i + 1)) / 2 - (size_t)i - 1;   // + j
    if (threadIdx.x == 0) { D[(size_t)i * n + i] = 0.f; if (S) S[(size_t)i * n + i] = 0.0; }
    for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
      const float v = tri[base + j];
      D[(size_t)i * n + j] = v; D[(size_t)j * n + i] = v;
      if (S) { S[(size_t)i * n + j] = (double)v; S[(size_t)j * n + i] = (double)v; }
    }
  }
}
/// closest / cmin of the initial matrix (one warp per row: lowest index among the minima, as the sequence of SetCdist
/// calls of the initial build leaves it) and the per-cluster state.
__global__ void __launch_bounds__(256) hieragglo_init_kernel(HaArgs a) {
  const int lane = threadIdx.x & 31;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nW = (gridDim.x * blockDim.x) >> 5;
  for (int i = w; i < a.n; i += nW) {
    const HaMin2 m = ha_warp_row_min(a, i, lane);
    if (lane == 0) {
      const ha_u64 k = m.best;
      a.closest[i] = (k == ~0ull) ? -1 : (int)(unsigned int)k;
      a.cmin[i] = (k == ~0ull) ? __int_as_float(0x7f800000) : ha_unord((unsigned int)(k >> 32));
      a.lb2[i] = ha_lb_of(m.sec);
      a.nfr[i] = 1; a.rkey[i] = ~0ull; a.rlb[i] = 0xffffffffu;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.ctl->minKey[0] = a.ctl->minKey[1] = ~0ull; a.ctl->rowKey[0] = a.ctl->rowKey[1] = ~0ull;
    a.ctl->tieCount[0] = a.ctl->tieCount[1] = 0; a.ctl->nA[0] = a.ctl->nA[1] = 0; a.ctl->nB[0] = a.ctl->nB[1] = 0;
    a.ctl->nCalls = 0; a.ctl->nMerges = 0; a.ctl->stopped = 0;
    for (int i = 0; i < 8; ++i) a.ctl->clk[i] = 0;
    for (int i = 0; i < 4; ++i) a.ctl->cnt[i] = 0;
  }
}

}  // namespace b200

// ---------------------------------------------------------------------------------------------------------------------
// Cache consumers of the cluster post-processing (src/Cluster/BestReps.cpp:131-296, src/Cluster/Output.cpp:164-225,
// src/Cluster/Algorithm_HierAgglo.cpp:353-408): sums of cached distances over the members of the clusters.
// ---------------------------------------------------------------------------------------------------------------------
namespace b200 {

__device__ __forceinline__ size_t cache_tri_idx(int n, int a, int b) {   // src/Matrix.h:110-122
  if (a > b) { int t = a; a = b; b = t; }
  return (size_t)n * (size_t)a - ((size_t)a * ((size_t)a + 1)) / 2 + (size_t)b - (size_t)a - 1;
}

/// One thread per listed member p of a cluster: cum[p] = sum over the other members q of d(p, q), added in list order in
/// double exactly as the inner loop of BestReps::FindBestRepFrames_CumulativeDist does (so that equal candidates -- duplicate
/// frames -- compare as in the reference); up[p] / up2[p] = sums of d and d^2 over the members AFTER p (the contribution of
/// row p to the within-cluster average and its standard deviation, Output.cpp:195-225).
__global__ void __launch_bounds__(128) cache_cluster_sums_kernel(const float* __restrict__ tri, int n, const int* __restrict__ members,
                                                                 const int* __restrict__ offsets, const int* __restrict__ clusterOf,
                                                                 int total, double* cum, double* up, double* up2) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= total) return;
  const int c = clusterOf[p], lo = offsets[c], hi = offsets[c + 1], i = members[p];
  double s = 0.0, u = 0.0, u2 = 0.0;
  for (int q = lo; q < hi; ++q) {
    if (q == p) continue;
    const double d = (double)tri[cache_tri_idx(n, i, members[q])];
    s = __dadd_rn(s, d);
    if (q > p) { u = __dadd_rn(u, d); u2 = __dadd_rn(u2, __dmul_rn(d, d)); }
  }
  cum[p] = s;
  if (up) up[p] = u;
  if (up2) up2[p] = u2;
}

/// Linkage between every pair of clusters from ONE pass over the cache triangle (Algorithm_HierAgglo::ClusterDistance for
/// all pairs: the reference walks the frame pairs of every cluster pair, N^2/2 cache reads through virtual calls).
/// label[f] = cluster of cached frame f or -1 (not in a cluster / sieved out and excluded).  table[c1 * K + c2], c1 < c2:
/// minimum and maximum as ordered bits, sum in double, count.
struct LinkCell { unsigned int mn, mx; unsigned long long cnt; double sum; };
__global__ void __launch_bounds__(256) cache_cluster_links_kernel(const float* __restrict__ tri, int n, const int* __restrict__ label,
                                                                  int K, LinkCell* table) {
  extern __shared__ unsigned char smem_links[];
  LinkCell* loc = reinterpret_cast<LinkCell*>(smem_links);
  const bool useSmem = (size_t)K * K * sizeof(LinkCell) <= 40960;
  if (useSmem) {
    for (int x = threadIdx.x; x < K * K; x += blockDim.x) { loc[x].mn = 0xffffffffu; loc[x].mx = 0u; loc[x].cnt = 0ull; loc[x].sum = 0.0; }
    __syncthreads();
  }
  LinkCell* tgt = useSmem ? loc : table;
  for (int i = blockIdx.x; i < n - 1; i += gridDim.x) {
    const int ci = label[i];
    if (ci < 0) continue;
    const size_t base = (size_t)n * (size_t)i - ((size_t)i * ((size_t)i + 1)) / 2 - (size_t)i - 1;   // + j
    for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
      const int cj = label[j];
      if (cj < 0 || cj == ci) continue;
      const float v = tri[base + j];
      LinkCell* cell = tgt + (ci < cj ? (size_t)ci * K + cj : (size_t)cj * K + ci);
      const unsigned int o = ha_ord(v);
      atomicMin(&cell->mn, o); atomicMax(&cell->mx, o);
      atomicAdd(&cell->cnt, 1ull); atomicAdd(&cell->sum, (double)v);
    }
  }
  if (useSmem) {
    __syncthreads();
    for (int x = threadIdx.x; x < K * K; x += blockDim.x) {
      if (loc[x].cnt == 0ull) continue;
      atomicMin(&table[x].mn, loc[x].mn); atomicMax(&table[x].mx, loc[x].mx);
      atomicAdd(&table[x].cnt, loc[x].cnt); atomicAdd(&table[x].sum, loc[x].sum);
    }
  }
}

}  // namespace b200
