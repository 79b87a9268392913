// Arithmetic-only microbenchmarks of the field / point cores in csrc/fq.cuh and
// csrc/point.cuh: how close a register-resident loop of fq_mul / fq_sqr / bucket
// additions gets to the IMAD.WIDE issue rate (the roofline denominator) when no memory
// traffic is involved.  The gap between these figures and k_msm_accumulate is what the
// gathers and the bookkeeping cost; the gap between these figures and 100 % is what the
// ALU glue between the multiply rows costs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -Idecaf377_b200/csrc -o tools/ub_field tools/ub_field.cu
// (profiles/r1_ub_field_before_lazy.txt holds the figures of the eagerly reduced field core
// this file measured first: 85 % for the bucket addition against 95 % after.)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "point.cuh"


#ifndef UB_MINB
#define UB_MINB 4
#endif

template <int V>
__global__ void __launch_bounds__(128, UB_MINB) k(uint32_t* io, int iters) {
  __shared__ uint32_t tab[16 * 32 + 8];
  for (int i = threadIdx.x; i < 16 * 32 + 8; i += blockDim.x) tab[i] = io[i] & 0x0fffffffu;
  __syncthreads();
  const uint32_t* src = io + 8 * (threadIdx.x & 31);
  fq_t x = fq_load(src), y = fq_load(src + 256);
  x.l[7] &= 0x0fffffffu; y.l[7] &= 0x0fffffffu;   // < q: keeps the static bounds honest
  if (V == 0) {  // one dependent multiplication chain per thread
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 4; r++) x = fq_mul(x, y);
    }
  } else if (V == 1) {  // two independent chains
    fq_t z = y;
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 2; r++) { x = fq_mul(x, y); z = fq_mul(z, y); }
    }
    x = fq_fold(fq_add(x, z));
  } else if (V == 2) {  // squaring chain
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 4; r++) x = fq_sqr(x);
    }
  } else if (V == 3 || V == 4 || V == 6) {  // bucket addition, operand from shared memory
    pt_t acc;
    acc.x = x; acc.y = y; acc.z = fq_fold(fq_add(x, y)); acc.t = fq_mul(x, y);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
      const uint32_t* p = tab + ((it * 7 + threadIdx.x) & 15) * 32;
      const bool neg = (V == 4) ? false : ((it ^ threadIdx.x) & 1);
      const int o = neg ? 8 : 0;
      if (V == 6) {  // 7M mixed addition, sign by address (k_msm_accumulate<true>)
        fq_r ymx, ypx, kt;
#pragma unroll
        for (int j = 0; j < 8; j++) { ymx.l[j] = p[o + j]; ypx.l[j] = p[8 - o + j]; kt.l[j] = p[16 + o + j]; }
        acc = pt_add_affine<true>(acc, ymx, ypx, kt);
      } else {       // 8M cached projective addition (k_msm_accumulate<false>)
        cached_t c;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          c.ymx.l[j] = p[o + j]; c.ypx.l[j] = p[8 - o + j]; c.kt.l[j] = p[16 + j]; c.z2.l[j] = p[24 + j];
        }
        acc = pt_add_cached<true, true>(acc, c, neg);
      }
    }
    x = fq_fold(fq_add(fq_fold(fq_add(acc.x, acc.y)), fq_fold(fq_add(acc.z, acc.t))));
  } else if (V == 5) {  // doubling chain
    pt_t acc;
    acc.x = x; acc.y = y; acc.z = fq_fold(fq_add(x, y)); acc.t = fq_mul(x, y);
#pragma unroll 1
    for (int it = 0; it < iters; it++) acc = pt_dbl<true>(acc);
    x = fq_fold(fq_add(fq_fold(fq_add(acc.x, acc.y)), fq_fold(fq_add(acc.z, acc.t))));
  }
  if (x.l[0] == 0x12345u && x.l[3] == 77u) fq_store(io + 8 * threadIdx.x, x);
}

template <int V>
void run(const char* name, double wide_per_iter, int iters, int sms, double peak) {
  uint32_t* d;
  cudaMalloc(&d, 65536);
  cudaMemset(d, 0x5a, 65536);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, k<V>);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<V>, 128, 0);
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  int grid = sms * occ * 4;
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(t0);
    k<V><<<grid, 128>>>(d, iters);
    cudaEventRecord(t1);
    cudaEventSynchronize(t1);
    float ms;
    cudaEventElapsedTime(&ms, t0, t1);
    if (rep && ms < best) best = ms;
  }
  double rate = (double)grid * 128 * iters * wide_per_iter / (best * 1e-3);
  printf("%-46s regs %3d occ %d CTAs  %8.3f ms  %8.1f G wide/s  %5.1f%% of %.0f\n", name, fa.numRegs, occ,
         best, rate / 1e9, 100.0 * rate / peak, peak / 1e9);
  cudaFree(d);
}

// same peak probe as the library's d377_imad_peak
__global__ void __launch_bounds__(256) k_peak(uint32_t* out, uint32_t seed) {
  uint32_t x[16], b = seed * 3 + threadIdx.x * 5 + 7;
#pragma unroll
  for (int j = 0; j < 16; j++) x[j] = j * seed + threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < 2048; it++) {
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
  if (s == 0x1234567) out[0] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  uint32_t* d;
  cudaMalloc(&d, 4);
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(t0);
    k_peak<<<sms * 8, 256>>>(d, 12345u + rep);
    cudaEventRecord(t1);
    cudaEventSynchronize(t1);
    float ms;
    cudaEventElapsedTime(&ms, t0, t1);
    if (rep && ms < best) best = ms;
  }
  double peak = (double)sms * 8 * 256 * 2048 * 8 * 8 / (best * 1e-3);
  printf("%s, %d SMs; IMAD.WIDE peak %.1f G/s (min CTAs/SM hint %d)\n", p.name, sms, peak / 1e9, UB_MINB);
  run<0>("fq_mul, 1 chain", 4 * 120, 2000, sms, peak);
  run<1>("fq_mul, 2 chains", 4 * 120, 2000, sms, peak);
  run<2>("fq_sqr, 1 chain", 4 * 92, 2000, sms, peak);
  run<3>("pt_add_cached 8M (sign varies), smem operand", 8 * 120, 1000, sms, peak);
  run<4>("pt_add_cached 8M (sign fixed), smem operand", 8 * 120, 1000, sms, peak);
  run<6>("pt_add_affine 7M (sign by address), smem operand", 7 * 120, 1000, sms, peak);
  run<5>("pt_dbl (4S + 4M)", 4 * 92 + 4 * 120, 1000, sms, peak);
  return 0;
}
