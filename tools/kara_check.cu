// Bit-exactness of the Karatsuba multiplication (fq_mul_kara, an experiment kept behind
// -DD377_MUL_KARATSUBA) against the row-interleaved one the library uses, on 2^22 random and
// extreme operand pairs.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3
//   -Idecaf377_b200/csrc -o tools/kara_check tools/kara_check.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "fq.cuh"

__global__ void k(unsigned long long* bad, uint32_t seed) {
  uint32_t s = seed + (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
  auto next = [&]() { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; };
  fq_t a, b;
  for (int it = 0; it < 64; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) { a.l[i] = next(); b.l[i] = next(); }
    const uint32_t mode = next() & 7;
    if (mode == 0) { for (int i = 0; i < 4; i++) a.l[i] = 0xffffffffu; }
    if (mode == 1) { for (int i = 4; i < 8; i++) b.l[i] = 0xffffffffu; }
    if (mode == 2) { for (int i = 0; i < 8; i++) a.l[i] = 0xffffffffu; }
    // storage-class operands (< 2q): keep the static bounds honest
    a.l[7] &= 0x1fffffffu; b.l[7] &= 0x1fffffffu;
    auto x = fq_reduce(fq_mul_cios(a, b));
    auto y = fq_reduce(fq_mul_kara(a, b));
    bool same = true;
#pragma unroll
    for (int i = 0; i < 8; i++) same = same && x.l[i] == y.l[i];
    if (!same) atomicAdd(bad, 1ull);
  }
}

int main() {
  unsigned long long *d, h = 0;
  cudaMalloc(&d, 8);
  cudaMemset(d, 0, 8);
  k<<<512, 128>>>(d, 12345u);
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("kara_check: %d pairs, %llu mismatches (%s)\n", 512 * 128 * 64, h, cudaGetErrorString(cudaGetLastError()));
  return h != 0;
}
