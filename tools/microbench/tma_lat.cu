// TMA 1-D bulk copy (cp.async.bulk, global -> shared) latency / throughput probe on L2-resident data.
// One thread per CTA issues `n` copies of `S` bytes, waits for all of them on one mbarrier, repeats.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_lat tma_lat.cu ; run: ./tma_lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(32, 1) probe(const unsigned char* src, size_t footprint, int S, int n, int reps, long long* out, int sameAddr) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    const uint32_t b = smem_u32(&bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    size_t off = sameAddr ? 0 : ((size_t)blockIdx.x * 1315423911ull) % footprint;
    long long tot = 0;
    uint32_t phase = 0;
    for (int r = 0; r < reps + 2; ++r) {
      const long long t0 = clock64();
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"((uint32_t)(S * n)) : "memory");
      for (int k = 0; k < n; ++k) {
        off = (off + (size_t)S * 7919) % (footprint - (size_t)S);
        off &= ~(size_t)127;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm) + (uint32_t)(k * S)),
                     "l"(src + off), "r"((uint32_t)S), "r"(b)
                     : "memory");
      }
      asm volatile(
          "{\n.reg .pred p;\nW1:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra.uni D1;\nbra.uni W1;\nD1:\n}\n" ::"r"(b), "r"(phase)
          : "memory");
      phase ^= 1u;
      const long long t1 = clock64();
      if (r >= 2) tot += t1 - t0;
    }
    out[blockIdx.x] = tot / reps;
  }
}
int main() {
  const size_t footprint = 48ull << 20;
  unsigned char* src; long long* out;
  cudaMalloc(&src, footprint); cudaMemset(src, 1, footprint); cudaMalloc(&out, 148 * 8);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int sizes[] = {512, 2048, 8192, 16384, 32768};
  const int ns[] = {1, 2, 4, 8, 16};
  printf("%6s %3s %5s %5s | %10s %10s %10s\n", "S", "n", "ctas", "same", "cyc(avg)", "cyc(max)", "B/clk/SM");
  for (int same = 0; same < 2; ++same)
  for (int ctas : {1, 148})
    for (int S : sizes)
      for (int n : ns) {
        if ((size_t)S * n > 192 * 1024) continue;
        probe<<<ctas, 32, 192 * 1024>>>(src, footprint, S, n, 200, out, same);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        long long h[148]; cudaMemcpy(h, out, ctas * 8, cudaMemcpyDeviceToHost);
        double avg = 0; long long mx = 0; for (int i = 0; i < ctas; ++i) { avg += h[i]; if (h[i] > mx) mx = h[i]; } avg /= ctas;
        printf("%6d %3d %5d %5d | %10.0f %10lld %10.2f\n", S, n, ctas, same, avg, mx, (double)S * n / avg);
      }
  return 0;
}
