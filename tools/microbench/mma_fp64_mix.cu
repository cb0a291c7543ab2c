// Does FP64 CUDA-core work slow down while the tcgen05 int8 tensor pipe is busy (and vice versa)?
// One CTA per SM: thread 0 of warp 0 issues `iters` M128 N256 K32 kind::i8 MMAs back to back (if mmaOn);
// warps 4..4+nFp-1 each run 8 independent DFMA chains of `fpIters` steps (if nFp > 0).  Reports cycles of both.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../cpptraj_b200/csrc -I../../include -o mma_fp64_mix mma_fp64_mix.cu
#include <cstdio>
#include "pair_i8.cuh"
using namespace b200;
__global__ void __launch_bounds__(768, 1) mix(int mmaOn, int iters, int nFp, int fpIters, long long* out, double* sink, int ldOn) {
  extern __shared__ __align__(1024) unsigned char smem_pk[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmemSlot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (3 * I8_BLK_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_pk)[i] = 0x01010101u * (uint32_t)(i & 3);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<1>(smem_u32(&tmemSlot), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmemBase = tmemSlot;
  if (tid == 0 && mmaOn) {
    constexpr uint32_t idesc = umma_idesc_i8(128, 256);
    const uint32_t sA = smem_u32(smem_pk), sB = sA + I8_BLK_BYTES;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
      umma_i8<1>(tmemBase + (uint32_t)((it & 1) * 256), umma_desc(sA + (it & 1) * 256, 128, 512), umma_desc(sB + (it & 1) * 256, 128, 512), idesc, (uint32_t)(it > 1));
    umma_commit<1>(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    out[2 * blockIdx.x] = clock64() - t0;
  }
  if (ldOn && warp < 4) {   // TMEM reads (tcgen05.ld 32x32b.x32) back to back on all four sub-partitions
    int v[32]; int acc = 0;
    const long long t0 = clock64();
    for (int i = 0; i < ldOn; ++i) {
      tmem_ld32(tmemBase + ((uint32_t)(32 * warp) << 16) + (uint32_t)((i & 7) * 32), v);
      tmem_ld_wait();
      acc += v[i & 31];
    }
    if (acc == 0x7fffffff) sink[tid] = acc;
    if (tid == 32) out[2 * blockIdx.x] = clock64() - t0;
  }
  if (warp >= 4 && warp < 4 + nFp) {
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = 1.5 + tid * 1e-3 + k;
    const double a = 1.0000001, b = 1e-9;
    const long long t0 = clock64();
    for (int i = 0; i < fpIters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) x[k] = fma(x[k], a, b);
    }
    const long long t1 = clock64();
    double s = 0; for (int k = 0; k < 8; ++k) s += x[k];
    if (s == 12345.678) sink[tid] = s;
    if (warp == 4 && (tid & 31) == 0) out[2 * blockIdx.x + 1] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_before(); tmem_dealloc<1>(tmemBase, 512); }
}
int main() {
  long long* out; double* sink;
  cudaMalloc(&out, 148 * 16); cudaMalloc(&sink, 8 * 1024);
  const int smem = 3 * I8_BLK_BYTES;
  cudaFuncSetAttribute(mix, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 20000, fpIters = 4000;
  printf("%5s %4s | %12s %14s | %12s %16s\n", "mma", "nFp", "mma cycles", "cyc/MMA", "fp cycles", "DFMA lanes/clk/SM");
  for (int mmaOn : {1, 0})
    for (int nFp : {0, 4, 8, 16, 20}) {
      if (!mmaOn && !nFp) continue;
      cudaMemset(out, 0, 148 * 16);
      mix<<<148, 768, smem>>>(mmaOn, iters, nFp, fpIters, out, sink, 0);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      long long h[296]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
      double m = 0, f = 0; for (int i = 0; i < 148; ++i) { m += h[2 * i]; f += h[2 * i + 1]; } m /= 148; f /= 148;
      printf("%5d %4d | %12.0f %14.1f | %12.0f %16.1f\n", mmaOn, nFp, m, m / iters, f, f > 0 ? (double)nFp * 32 * 8 * fpIters / f : 0.0);
    }
  printf("--- FP64 under TMEM reads only (no MMAs): ld iterations 20000\n");
  for (int nFp : {4, 16}) {
    cudaMemset(out, 0, 148 * 16);
    mix<<<148, 768, smem>>>(0, iters, nFp, fpIters, out, sink, 20000);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    long long h[296]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0, f = 0; for (int i = 0; i < 148; ++i) { m += h[2 * i]; f += h[2 * i + 1]; } m /= 148; f /= 148;
    printf("ld cycles %12.0f (%.1f per x32 load) | nFp %2d fp cycles %12.0f  DFMA lanes/clk/SM %.1f\n", m, m / 20000, nFp, f, (double)nFp * 32 * 8 * fpIters / f);
  }
  return 0;
}
