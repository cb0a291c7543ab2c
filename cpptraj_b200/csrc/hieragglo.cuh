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
// history-dependent tie case of the merged cluster's own closest index is replayed sequentially.
//
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
};

struct HaArgs {
  float* D;            // cluster-distance triangle (starts as the pairwise cache), src/Matrix.h:110-122 layout
  double* S;           // average linkage: sums of frame-pair distances (same layout), else null
  int n, linkage, target;
  double eps;
  int* closest; float* cmin;        // per cluster: closest cluster and D(i, closest[i])
  unsigned char* ign; int* nfr;     // merged-away flag, frame count
  float* vnew; double* snew; float* oold;   // the merged cluster's new row, its sums, its old row
  int* listA; int* listB;
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
__device__ __forceinline__ size_t ha_idx(int n, int a, int b) {
  if (a > b) { int t = a; a = b; b = t; }
  return (size_t)n * (size_t)a - ((size_t)a * ((size_t)a + 1)) / 2 + (size_t)b - (size_t)a - 1;
}
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

/// DynamicMatrix::updateClosestIdx (DynamicMatrix.h:43-62) for row idx by the whole CTA: lowest index among the minima
/// over the clusters still present.
__device__ __forceinline__ void ha_rescan_row(const HaArgs& a, int idx, ha_u64* sm) {
  ha_u64 best = ~0ull;
  for (int j = threadIdx.x; j < a.n; j += blockDim.x)
    if (j != idx && !a.ign[j]) {
      ha_u64 k = ha_key(a.D[ha_idx(a.n, idx, j)], j);
      best = k < best ? k : best;
    }
  best = ha_block_min(best, sm);
  if (threadIdx.x == 0) {
    if (best == ~0ull) { a.closest[idx] = -1; a.cmin[idx] = __int_as_float(0x7f800000); }
    else { a.closest[idx] = (int)(unsigned int)best; a.cmin[idx] = ha_unord((unsigned int)(best >> 32)); }
  }
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
  for (int base = 0; base < n; base += 32) {
    const int k = base + lane;
    const bool valid = k < n && k != C1 && !a.ign[k];
    const float v = valid ? a.vnew[k] : 0.f;
    const float o = valid ? a.oold[k] : 0.f;
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
  if (lane == 0) { a.closest[C1] = c; a.cmin[C1] = a.vnew[c]; }
}

/// All merges in one launch.  Launch with ONE cluster of `team` CTAs x 1024 threads.
__global__ void __launch_bounds__(1024, 1) hieragglo_kernel(HaArgs a) {
  cg::cluster_group team = cg::this_cluster();
  __shared__ ha_u64 sm[34];
  const int nCta = (int)team.num_blocks();
  const int cta = (int)team.block_rank();
  const int tid = cta * blockDim.x + threadIdx.x;
  const int nThr = nCta * blockDim.x;
  const bool leader = (tid == 0);
  const int n = a.n;
  volatile HaCtl* ctl = a.ctl;
  int nClusters = n;
  int prevC1 = -1, prevK = -1; float prevV = 0.f;   // previous merge: C1's closest when the row minimum was unique

  for (int m = 0;; ++m) {
    const int p = m & 1;
    // ---- FindMin (DynamicMatrix.cpp:7-33): lowest column among equal minima
    ha_u64 best = ~0ull;
    for (int col = tid; col < n; col += nThr) {
      if (a.ign[col]) continue;
      float cm;
      if (col == prevC1) { cm = prevV; a.closest[col] = prevK; a.cmin[col] = prevV; }
      else cm = a.cmin[col];
      if (a.closest[col] < 0 && col != prevC1) continue;
      ha_u64 k = ha_key(cm, col);
      best = k < best ? k : best;
    }
    best = ha_block_min(best, sm);
    if (threadIdx.x == 0 && best != ~0ull) atomicMin((ha_u64*)&a.ctl->minKey[p], best);
    team.sync();
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
    for (int k = tid; k < n; k += nThr) {
      if (k == C1 || k == C2 || a.ign[k]) continue;
      const size_t i1 = ha_idx(n, C1, k), i2 = ha_idx(n, C2, k);
      const float o1 = a.D[i1], o2 = a.D[i2];
      float v;
      if (a.linkage == 0) v = o2 < o1 ? o2 : o1;
      else if (a.linkage == 2) v = o2 > o1 ? o2 : o1;
      else {
        const double s = a.S[i1] + a.S[i2];
        a.snew[k] = s;
        v = (float)(s / (double)(n1 * a.nfr[k]));
      }
      a.vnew[k] = v; a.oold[k] = o1;
      if (a.closest[k] == C2) a.listA[atomicAdd((int*)&a.ctl->nA[p], 1)] = k;
      ha_u64 key = ha_key(v, k);
      best = key < best ? key : best;
    }
    best = ha_block_min(best, sm);
    if (threadIdx.x == 0 && best != ~0ull) atomicMin((ha_u64*)&a.ctl->rowKey[p], best);
    team.sync();
    // ---- Ignore(C2) (DynamicMatrix.h:116-126): re-scan, on the OLD matrix, every cluster whose closest was C2
    const int nA = ctl->nA[p];
    if (nA > 0) {
      for (int e = cta; e < nA; e += nCta) ha_rescan_row(a, a.listA[e], sm);
      team.sync();
    }
    // ---- SetCdist(C1, k, v_k), row side (DynamicMatrix.h:88-101), and the update itself
    const ha_u64 rk = ctl->rowKey[p];
    const unsigned int vminOrd = (unsigned int)(rk >> 32);
    for (int k = tid; k < n; k += nThr) {
      if (k == C1 || k == C2 || a.ign[k]) continue;
      const float v = a.vnew[k];
      const float cd = a.cmin[k];
      const int ck = a.closest[k];
      if (ck < 0 || v < cd) { a.closest[k] = C1; a.cmin[k] = v; }
      else if (ck == C1 && v > cd) a.listB[atomicAdd((int*)&a.ctl->nB[p], 1)] = k;
      const size_t i1 = ha_idx(n, C1, k);
      a.D[i1] = v;
      if (a.linkage == 1) a.S[i1] = a.snew[k];
      if (ha_ord(v) == vminOrd) atomicAdd((int*)&a.ctl->tieCount[p], 1);
    }
    if (leader) a.nfr[C1] = n1;
    team.sync();
    // ---- re-scans after the update: rows whose closest distance (to C1) grew; C1's own closest
    const int nB = ctl->nB[p];
    const int ties = ctl->tieCount[p];
    if (nB > 0 || ties > 1) {
      for (int e = cta; e < nB; e += nCta) ha_rescan_row(a, a.listB[e], sm);
      if (ties > 1 && cta == 0 && threadIdx.x < 32) ha_tie_replay(a, C1, C2);
      team.sync();
    }
    if (ties == 1) { prevC1 = C1; prevK = (int)(unsigned int)rk; prevV = ha_unord(vminOrd); }
    else prevC1 = -1;
    if (nClusters <= a.target || nClusters == 1) break;
  }
}

/// closest / cmin of the initial matrix: one pass over the triangle, (distance, index) minima by 64-bit atomicMin.
__global__ void __launch_bounds__(256) hieragglo_init_kernel(const float* __restrict__ D, int n, ha_u64* keys) {
  __shared__ ha_u64 sm[34];
  for (int i = blockIdx.x; i < n - 1; i += gridDim.x) {
    const size_t base = (size_t)n * (size_t)i - ((size_t)i * ((size_t)i + 1)) / 2 - (size_t)i - 1;   // + j
    ha_u64 best = ~0ull;
    for (int j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
      const float v = D[base + j];
      ha_u64 kr = ha_key(v, j);
      best = kr < best ? kr : best;
      ha_u64 kc = ha_key(v, i);
      if (kc < *(volatile ha_u64*)&keys[j]) atomicMin(&keys[j], kc);
    }
    best = ha_block_min(best, sm);
    if (threadIdx.x == 0 && best != ~0ull) atomicMin(&keys[i], best);
  }
}
__global__ void hieragglo_init2_kernel(HaArgs a, const ha_u64* keys, size_t nElt) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nT = (size_t)gridDim.x * blockDim.x;
  for (size_t i = t; i < (size_t)a.n; i += nT) {
    ha_u64 k = keys[i];
    a.closest[i] = (k == ~0ull) ? -1 : (int)(unsigned int)k;
    a.cmin[i] = (k == ~0ull) ? __int_as_float(0x7f800000) : ha_unord((unsigned int)(k >> 32));
    a.ign[i] = 0; a.nfr[i] = 1;
  }
  if (a.S != nullptr)
    for (size_t e = t; e < nElt; e += nT) a.S[e] = (double)a.D[e];
  if (t == 0) {
    a.ctl->minKey[0] = a.ctl->minKey[1] = ~0ull; a.ctl->rowKey[0] = a.ctl->rowKey[1] = ~0ull;
    a.ctl->tieCount[0] = a.ctl->tieCount[1] = 0; a.ctl->nA[0] = a.ctl->nA[1] = 0; a.ctl->nB[0] = a.ctl->nB[1] = 0;
    a.ctl->nCalls = 0; a.ctl->nMerges = 0; a.ctl->stopped = 0;
  }
}

}  // namespace b200
