// Probe: does a pitched host->device copy that skips the T coordinate of each 128-byte
// Element (96 of every 128 bytes) beat the flat copy over PCIe?  Build:
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/h2d_2d_probe.cu -o gpurun_out/h2d_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main() {
  const size_t n = 1u << 24;
  uint8_t *h, *d;
  CK(cudaHostAlloc(&h, n * 128, cudaHostAllocDefault));
  memset(h, 1, n * 128);
  CK(cudaMalloc(&d, n * 128));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int mode = 0; mode < 5; mode++) {
    float best = 1e9f;
    for (int it = 0; it < 4; it++) {
      CK(cudaEventRecord(a, st));
      if (mode == 0) CK(cudaMemcpyAsync(d, h, n * 128, cudaMemcpyHostToDevice, st));
      if (mode == 1) CK(cudaMemcpy2DAsync(d, 96, h, 128, 96, n, cudaMemcpyHostToDevice, st));
      if (mode == 2) CK(cudaMemcpy2DAsync(d, 128, h, 128, 96, n, cudaMemcpyHostToDevice, st));
      if (mode == 3) CK(cudaMemcpyAsync(d, h, n * 96, cudaMemcpyHostToDevice, st));
      if (mode == 4) {  // 16 pitched copies of 2^20 rows, as the chunked pipeline would issue
        for (int c = 0; c < 16; c++)
          CK(cudaMemcpy2DAsync(d + (size_t)c * (n / 16) * 96, 96, h + (size_t)c * (n / 16) * 128, 128, 96,
                               n / 16, cudaMemcpyHostToDevice, st));
      }
      CK(cudaEventRecord(b, st));
      CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      if (ms < best) best = ms;
    }
    const char* nm[] = {"flat 128 B/row", "2D 96 of 128 -> packed 96", "2D 96 of 128 -> pitch 128",
                        "flat 96 B/row (lower bound)", "2D 96 of 128, 16 chunks"};
    printf("%-32s %8.2f ms  %6.1f GB/s payload  %6.1f Mrows/s\n", nm[mode], best,
           (mode == 0 ? 128.0 : 96.0) * n / best / 1e6, n / best / 1e3);
  }
  return 0;
}
