#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) kk(uint32_t* out, uint32_t seed, int iters) {
  uint32_t b = seed * 3 + threadIdx.x * 5 + 7;
  uint32_t x[16];
#pragma unroll
  for (int j = 0; j < 16; j++) x[j] = j * seed + threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int j = 0; j < 8; j++)
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                     : "+r"(x[2 * j]), "+r"(x[2 * j + 1]) : "r"(x[(2 * j + 3) & 15]), "r"(b));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 16; j++) s ^= x[j];
  if (s == 0x1234567u) out[0] = s;
}
int main() {
  uint32_t* d; cudaMalloc(&d, 4);
  cudaEvent_t e[402]; for (int i=0;i<402;i++) cudaEventCreate(&e[i]);
  int grid = 148*8;
  for (int i=0;i<400;i++){ cudaEventRecord(e[i]); kk<<<grid,256>>>(d, 12345u+i, 2048);} cudaEventRecord(e[400]);
  cudaDeviceSynchronize();
  for (int i=0;i<400;i+= (i<20?1:20)) { float ms; cudaEventElapsedTime(&ms,e[i],e[i+1]); printf("launch %3d: %.3f ms  %.1f Gop/s\n", i, ms, (double)grid*256*2048*64/ms/1e6); }
  return 0;
}
