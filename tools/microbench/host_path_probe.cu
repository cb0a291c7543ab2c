// host_path_probe.cu -- what the host side of the rms2d path can do on this box (SURVEY.md section 7 hard parts,
// VERDICT r1 weak 2/3): PCIe copy rates into pinned memory (one GPU and all GPUs at once = the host-ingest ceiling),
// the cost of cudaHostRegister on cpptraj-style pageable buffers, the driver's own pageable copies, and multi-threaded
// memcpy between a pinned stage and pageable memory (fresh = never touched, as cpptraj's new float[] is).
// build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o host_path_probe host_path_probe.cu -lpthread
// usage: host_path_probe [MiB per device = 1024] [max threads = 16]
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <sys/mman.h>
#include <stdint.h>

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

static void par_memcpy(char* d, const char* s, size_t n, int T) {
  if (T <= 1) { memcpy(d, s, n); return; }
  std::vector<std::thread> th;
  const size_t per = ((n + T - 1) / T + 4095) & ~(size_t)4095;
  for (int t = 0; t < T; ++t) {
    const size_t a = (size_t)t * per;
    if (a >= n) break;
    const size_t len = std::min(per, n - a);
    th.emplace_back([=] { memcpy(d + a, s + a, len); });
  }
  for (auto& x : th) x.join();
}
static bool g_thp = false;       // madvise(MADV_HUGEPAGE) on fresh buffers (the library can do that to the caller's output)
static char* fresh(size_t n) {   // pageable, never touched (what new float[n] of this size gives: an anonymous mapping)
  void* p = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (p == MAP_FAILED) return nullptr;
  if (g_thp) {
    const uintptr_t a = ((uintptr_t)p + (2u << 20) - 1) & ~(uintptr_t)((2u << 20) - 1);
    const uintptr_t b = ((uintptr_t)p + n) & ~(uintptr_t)((2u << 20) - 1);
    if (b > a && madvise((void*)a, b - a, MADV_HUGEPAGE) != 0) perror("madvise(MADV_HUGEPAGE)");
  }
  return (char*)p;
}

int main(int argc, char** argv) {
  const size_t MiB = argc > 1 ? atol(argv[1]) : 1024;
  const int maxT = argc > 2 ? atoi(argv[2]) : 16;
  const size_t n = MiB << 20;
  int nd = 0;
  CK(cudaGetDeviceCount(&nd));
  printf("devices %d, %zu MiB per device, host threads available %u\n", nd, MiB, std::thread::hardware_concurrency());
  std::vector<char*> dev(nd), pin(nd);
  std::vector<cudaStream_t> st(nd);
  for (int i = 0; i < nd; ++i) {
    CK(cudaSetDevice(i));
    CK(cudaMalloc(&dev[i], n)); CK(cudaMemset(dev[i], 1, n));
    double t0 = now();
    CK(cudaMallocHost(&pin[i], n));
    if (i == 0) printf("cudaMallocHost %zu MiB: %.1f ms\n", MiB, (now() - t0) * 1e3);
    CK(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
  }
  // ---- 1. pinned copies, one device, then all at once
  for (int dir = 0; dir < 2; ++dir) {
    for (int cnt = 1; cnt <= nd; cnt = (cnt == nd ? nd + 1 : std::min(nd, cnt * 2))) {
      double best = 1e9;
      for (int rep = 0; rep < 3; ++rep) {
        for (int i = 0; i < cnt; ++i) { CK(cudaSetDevice(i)); CK(cudaStreamSynchronize(st[i])); }
        double t0 = now();
        for (int i = 0; i < cnt; ++i) {
          CK(cudaSetDevice(i));
          if (dir == 0) CK(cudaMemcpyAsync(pin[i], dev[i], n, cudaMemcpyDeviceToHost, st[i]));
          else CK(cudaMemcpyAsync(dev[i], pin[i], n, cudaMemcpyHostToDevice, st[i]));
        }
        for (int i = 0; i < cnt; ++i) { CK(cudaSetDevice(i)); CK(cudaStreamSynchronize(st[i])); }
        best = std::min(best, now() - t0);
      }
      printf("pinned %s, %d device(s) at once: %.1f GB/s aggregate (%.1f per device)\n", dir == 0 ? "D2H" : "H2D", cnt,
             cnt * n / best / 1e9, n / best / 1e9);
    }
  }
  // both directions at once on device 0
  {
    cudaStream_t s2; CK(cudaSetDevice(0)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    char* dev2; CK(cudaMalloc(&dev2, n)); char* pin2; CK(cudaMallocHost(&pin2, n));
    double t0 = now();
    CK(cudaMemcpyAsync(pin[0], dev[0], n, cudaMemcpyDeviceToHost, st[0]));
    CK(cudaMemcpyAsync(dev2, pin2, n, cudaMemcpyHostToDevice, s2));
    CK(cudaStreamSynchronize(st[0])); CK(cudaStreamSynchronize(s2));
    printf("pinned D2H + H2D together, device 0: %.1f GB/s each way\n", n / (now() - t0) / 1e9);
    CK(cudaFree(dev2)); CK(cudaFreeHost(pin2));
  }
  CK(cudaSetDevice(0));
  // ---- 2. cudaHostRegister on pageable memory
  for (int touched = 0; touched < 2; ++touched) {
    char* p = fresh(n);
    if (touched) memset(p, 0, n);
    double t0 = now();
    CK(cudaHostRegister(p, n, cudaHostRegisterDefault));
    const double tr = now() - t0;
    t0 = now();
    CK(cudaMemcpyAsync(p, dev[0], n, cudaMemcpyDeviceToHost, st[0])); CK(cudaStreamSynchronize(st[0]));
    const double tc = now() - t0;
    t0 = now();
    CK(cudaHostUnregister(p));
    const double tu = now() - t0;
    printf("cudaHostRegister %s %zu MiB: register %.1f ms (%.2f GB/s), D2H into it %.1f GB/s, unregister %.1f ms\n",
           touched ? "touched" : "fresh  ", MiB, tr * 1e3, n / tr / 1e9, n / tc / 1e9, tu * 1e3);
    munmap(p, n);
  }
  // chunked registration (64 MiB pieces) -- can it be pipelined?
  {
    char* p = fresh(n); memset(p, 0, n);
    const size_t piece = (size_t)64 << 20;
    double t0 = now();
    for (size_t a = 0; a < n; a += piece) CK(cudaHostRegister(p + a, std::min(piece, n - a), cudaHostRegisterDefault));
    const double tr = now() - t0;
    t0 = now();
    for (size_t a = 0; a < n; a += piece) CK(cudaHostUnregister(p + a));
    printf("cudaHostRegister touched, 64 MiB pieces: %.2f GB/s register, %.2f GB/s unregister\n", n / tr / 1e9, n / (now() - t0) / 1e9);
    munmap(p, n);
  }
  // ---- 3. the driver's own pageable copies
  for (int touched = 0; touched < 2; ++touched) {
    char* p = fresh(n);
    if (touched) memset(p, 0, n);
    double t0 = now();
    CK(cudaMemcpy(p, dev[0], n, cudaMemcpyDeviceToHost));
    const double td = now() - t0;
    t0 = now();
    CK(cudaMemcpy(dev[0], p, n, cudaMemcpyHostToDevice));
    printf("driver pageable copy, %s: D2H %.1f GB/s, H2D %.1f GB/s\n", touched ? "touched" : "fresh  ", n / td / 1e9, n / (now() - t0) / 1e9);
    munmap(p, n);
  }
  // ---- 4. threaded memcpy pinned stage <-> pageable (second round: transparent huge pages requested for the fresh buffer)
  for (int thp = 0; thp < 2; ++thp) {
  g_thp = thp != 0;
  if (thp) {
    FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
    char line[128] = "?";
    if (f) { if (!fgets(line, sizeof(line), f)) line[0] = 0; fclose(f); }
    printf("--- with madvise(MADV_HUGEPAGE) on the fresh buffers; /sys/kernel/mm/transparent_hugepage/enabled: %s", line);
  }
  for (int T = (thp ? 4 : 1); T <= maxT; T *= 2) {
    char* p = fresh(n);
    double t0 = now();
    par_memcpy(p, pin[0], n, T);
    const double tf = now() - t0;
    t0 = now();
    par_memcpy(p, pin[0], n, T);
    const double tt = now() - t0;
    t0 = now();
    par_memcpy(pin[0], p, n, T);
    const double ti = now() - t0;
    printf("memcpy %2d threads: pinned->fresh pageable %.1f GB/s, pinned->touched %.1f GB/s, pageable->pinned %.1f GB/s\n", T,
           n / tf / 1e9, n / tt / 1e9, n / ti / 1e9);
    munmap(p, n);
  }
  }
  // ---- 5. the staged pipeline as the library runs it: D2H into a pinned ring of 32 MiB slots, T threads copy out
  for (int T : {4, 8, 16}) {
    if (T > maxT) break;
    char* p = fresh(n);
    const size_t slot = (size_t)32 << 20;
    const int NS = 4;
    cudaEvent_t ev[NS];
    for (int s = 0; s < NS; ++s) CK(cudaEventCreateWithFlags(&ev[s], cudaEventDisableTiming));
    double t0 = now();
    const size_t nslots = (n + slot - 1) / slot;
    for (size_t k = 0; k < nslots + NS; ++k) {
      if (k >= NS) {   // retire slot k - NS
        const size_t j = k - NS; const int s = (int)(j % NS);
        CK(cudaEventSynchronize(ev[s]));
        par_memcpy(p + j * slot, pin[0] + (size_t)s * slot, std::min(slot, n - j * slot), T);
      }
      if (k < nslots) {
        const int s = (int)(k % NS);
        CK(cudaMemcpyAsync(pin[0] + (size_t)s * slot, dev[0] + k * slot, std::min(slot, n - k * slot), cudaMemcpyDeviceToHost, st[0]));
        CK(cudaEventRecord(ev[s], st[0]));
      }
    }
    printf("staged D2H -> fresh pageable (THP requested), 32 MiB slots x %d, %2d copy threads: %.1f GB/s end to end\n", NS, T, n / (now() - t0) / 1e9);
    for (int s = 0; s < NS; ++s) cudaEventDestroy(ev[s]);
    munmap(p, n);
  }
  return 0;
}
