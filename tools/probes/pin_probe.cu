// pin_probe.cu -- how long does it take to page-lock 1 GiB of fresh host memory, three ways?
// nvcc -O2 -o pin_probe pin_probe.cu && ./pin_probe
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <sys/mman.h>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  const size_t n = 1ull << 30;
  cudaFree(0);
  void* d;
  cudaMalloc(&d, n);
  for (int rep = 0; rep < 2; ++rep) {
    double t0 = now();
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, n, cudaHostAllocDefault);
    double t1 = now();
    cudaMemcpy(d, p, n, cudaMemcpyHostToDevice);
    double t2 = now();
    printf("cudaHostAlloc 1 GiB: %.3f s (%s), first H2D %.3f s\n", t1 - t0, cudaGetErrorString(e), t2 - t1);
    cudaFreeHost(p);
    t0 = now();
    void* q = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    int mr = madvise(q, n, MADV_HUGEPAGE);
    e = cudaHostRegister(q, n, cudaHostRegisterDefault);
    t1 = now();
    cudaMemcpy(d, q, n, cudaMemcpyHostToDevice);
    t2 = now();
    printf("mmap + MADV_HUGEPAGE(%d) + cudaHostRegister 1 GiB: %.3f s (%s), first H2D %.3f s\n", mr, t1 - t0, cudaGetErrorString(e), t2 - t1);
    cudaHostUnregister(q);
    munmap(q, n);
    t0 = now();
    q = mmap(nullptr, n, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_POPULATE, -1, 0);
    t1 = now();
    e = cudaHostRegister(q, n, cudaHostRegisterDefault);
    t2 = now();
    printf("mmap MAP_POPULATE %.3f s + cudaHostRegister %.3f s (%s)\n", t1 - t0, t2 - t1, cudaGetErrorString(e));
    cudaHostUnregister(q);
    munmap(q, n);
    t0 = now();
    q = malloc(n);
    memset(q, 0, n);
    t1 = now();
    cudaMemcpy(d, q, n, cudaMemcpyHostToDevice);
    t2 = now();
    cudaMemcpy(q, d, n, cudaMemcpyDeviceToHost);
    double t3 = now();
    printf("pageable: touch %.3f s, H2D %.3f s (%.1f GB/s), D2H %.3f s (%.1f GB/s)\n", t1 - t0, t2 - t1, n / 1e9 / (t2 - t1), t3 - t2, n / 1e9 / (t3 - t2));
    free(q);
  }
  FILE* f = fopen("/sys/kernel/mm/transparent_hugepage/enabled", "r");
  if (f) { char b[128] = {0}; fgets(b, 127, f); printf("THP: %s", b); fclose(f); }
  return 0;
}
