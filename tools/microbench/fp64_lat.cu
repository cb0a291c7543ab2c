// Micro-benchmarks behind DESIGN.md's epilogue decisions: dependent-issue latency (one warp, one chain) and
// per-SM throughput (many warps, 8 independent chains each) of the FP64 / conversion instructions the tcgen05
// epilogue uses, on the device it runs on.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int OP> __device__ __forceinline__ double step(double x, double a, double b) {
  if (OP == 0) return fma(x, a, b);              // DFMA
  if (OP == 1) return x + a;                     // DADD
  if (OP == 2) return x * a;                     // DMUL
  if (OP == 3) return 1.0 / x + b;               // DDIV-ish (MUFU.RCP64H + DFMAs)
  if (OP == 4) return (double)(long long)(x) + a;   // F2I.S64.F64 + I2F.F64.S64
  if (OP == 5) return (double)(float)x + a;      // F2F.F32.F64 + F2F.F64.F32
  if (OP == 6) return sqrt(x) + a;               // MUFU.RSQ64H + DFMAs
  if (OP == 7) return (double)__fmaf_rn((float)x, 1.0001f, 0.5f);  // conversions + FFMA
  return x;
}
template <int OP> __global__ void lat(double* out, long long* clk, double a, double b) {
  double x = 1.5 + threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = step<OP>(x, a, b);
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int OP> __global__ void thr(double* out, long long* clk, double a, double b) {
  double x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = 1.5 + threadIdx.x * 1e-3 + k;
  long long t0 = clock64();
  for (int i = 0; i < N / 8; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = step<OP>(x[k], a, b);
  }
  long long t1 = clock64();
  double s = 0; for (int k = 0; k < 8; ++k) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name, double a, double b) {
  double* out; long long* clk; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&clk, 148 * 8);
  long long h[148];
  lat<OP><<<1, 32>>>(out, clk, a, b); cudaDeviceSynchronize();
  lat<OP><<<1, 32>>>(out, clk, a, b); cudaMemcpy(h, clk, 8, cudaMemcpyDeviceToHost);
  double l = (double)h[0] / N;
  thr<OP><<<148, 1024>>>(out, clk, a, b); cudaDeviceSynchronize();
  thr<OP><<<148, 1024>>>(out, clk, a, b); cudaMemcpy(h, clk, 148 * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  // 1024 threads x N ops per SM in c cycles
  printf("%-44s latency %7.1f cyc/op   throughput %6.2f lanes/cyc/SM\n", name, l, 1024.0 * N / c);
  cudaFree(out); cudaFree(clk);
}
int main() {
  run<0>("DFMA", 1.0000001, 1e-9);
  run<1>("DADD", 1e-9, 0);
  run<2>("DMUL", 1.0000001, 0);
  run<3>("1.0/x + b (MUFU.RCP64H + DFMA chain)", 0, 0.5);
  run<4>("(double)(long long)x + a (F2I + I2F.F64.S64)", 0.25, 0);
  run<5>("(double)(float)x + a (F2F both ways)", 0.25, 0);
  run<6>("sqrt(x) + a (MUFU.RSQ64H + DFMA chain)", 2.0, 0);
  run<7>("double->float, FFMA, float->double", 0, 0);
  return 0;
}
