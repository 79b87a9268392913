// Instruction-throughput microbenchmarks for the integer multiply paths used by fq.cuh.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench.cu ; run on B200.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define REPS 8

template <int V>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 3 + threadIdx.x * 5 + 7;
  uint32_t x[16];
#pragma unroll
  for (int j = 0; j < 16; j++) x[j] = j * seed + (V == 11 ? threadIdx.x * 2654435761u + blockIdx.x : 0u);
#pragma unroll 1
  for (int it = 0; it < ((V == 9 || V == 10) ? 0 : ITERS); it++) {
#pragma unroll
    for (int r = 0; r < REPS; r++) {
      if (V == 0) {  // plain IMAD.WIDE, 8 independent 64-bit accumulators
#pragma unroll
        for (int j = 0; j < 8; j++) {
          uint64_t acc = ((uint64_t)x[2 * j + 1] << 32) | x[2 * j];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x[(2 * j + 3) & 15]), "r"(b));
          x[2 * j] = (uint32_t)acc; x[2 * j + 1] = (uint32_t)(acc >> 32);
        }
      } else if (V == 1) {  // two carry chains of 4 IMAD.WIDE (.cc / .X), as in fq_mul
        a ^= x[15];
        asm volatile(
            "mad.lo.cc.u32 %0, %16, %17, %0;\n\tmadc.hi.cc.u32 %1, %16, %17, %1;\n\t"
            "madc.lo.cc.u32 %2, %16, %17, %2;\n\tmadc.hi.cc.u32 %3, %16, %17, %3;\n\t"
            "madc.lo.cc.u32 %4, %16, %17, %4;\n\tmadc.hi.cc.u32 %5, %16, %17, %5;\n\t"
            "madc.lo.cc.u32 %6, %16, %17, %6;\n\tmadc.hi.u32 %7, %16, %17, %7;\n\t"
            "mad.lo.cc.u32 %8, %16, %17, %8;\n\tmadc.hi.cc.u32 %9, %16, %17, %9;\n\t"
            "madc.lo.cc.u32 %10, %16, %17, %10;\n\tmadc.hi.cc.u32 %11, %16, %17, %11;\n\t"
            "madc.lo.cc.u32 %12, %16, %17, %12;\n\tmadc.hi.cc.u32 %13, %16, %17, %13;\n\t"
            "madc.lo.cc.u32 %14, %16, %17, %14;\n\tmadc.hi.u32 %15, %16, %17, %15;"
            : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]),
              "+r"(x[8]), "+r"(x[9]), "+r"(x[10]), "+r"(x[11]), "+r"(x[12]), "+r"(x[13]), "+r"(x[14]), "+r"(x[15])
            : "r"(a), "r"(b));
      } else if (V == 2) {  // mad.lo only (IMAD), 8 independent
#pragma unroll
        for (int j = 0; j < 8; j++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[j]) : "r"(x[(j + 1) & 7]), "r"(b));
      } else if (V == 3) {  // mad.hi only (IMAD.HI), 8 independent
#pragma unroll
        for (int j = 0; j < 8; j++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x[j]) : "r"(x[(j + 1) & 7] | 0x80000000u), "r"(b));
      } else if (V == 4 || V == 11) {  // carry-out only (each pair independent: mad.lo.cc + madc.hi)
#pragma unroll
        for (int j = 0; j < 8; j++)
          asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                       : "+r"(x[2 * j]), "+r"(x[2 * j + 1]) : "r"(x[(2 * j + 3) & 15]), "r"(b));
      } else if (V == 5) {  // IADD3 carry chains: 8-limb add.cc chains x2
        asm volatile(
            "add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %10;\n\taddc.cc.u32 %3, %3, %11;\n\t"
            "addc.cc.u32 %4, %4, %12;\n\taddc.cc.u32 %5, %5, %13;\n\taddc.cc.u32 %6, %6, %14;\n\taddc.u32 %7, %7, %15;"
            : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7])
            : "r"(x[8]), "r"(x[9]), "r"(x[10]), "r"(x[11]), "r"(x[12]), "r"(x[13]), "r"(x[14]), "r"(x[15]));
      } else if (V == 6) {  // plain IMAD.WIDE interleaved with independent ALU ops (LOP3/SHF)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          uint64_t acc = ((uint64_t)x[2 * j + 1] << 32) | x[2 * j];
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x[(2 * j + 3) & 7]), "r"(b));
          x[2 * j] = (uint32_t)acc; x[2 * j + 1] = (uint32_t)(acc >> 32);
          x[8 + 2 * j] = (x[8 + 2 * j] >> 3) ^ a;
          x[9 + 2 * j] = (x[9 + 2 * j] & b) + 5;
        }
      } else if (V == 7) {  // 64-bit add chain via add.cc.u64? use IADD3-based 64-bit adds
#pragma unroll
        for (int j = 0; j < 8; j++) {
          uint64_t acc = ((uint64_t)x[2 * j + 1] << 32) | x[2 * j];
          acc += ((uint64_t)b << 32) | a;
          asm volatile("" : "+l"(acc));
          x[2 * j] = (uint32_t)acc; x[2 * j + 1] = (uint32_t)(acc >> 32);
        }
      } else if (V == 8) {  // IMAD.WIDE carry-in only chain end (madc.lo + madc.hi w/o cc) preceded by add.cc
#pragma unroll
        for (int j = 0; j < 8; j++)
          asm volatile("add.cc.u32 %2, %2, %3;\n\tmadc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                       : "+r"(x[2 * j]), "+r"(x[2 * j + 1]), "+r"(a) : "r"(b));
      }
    }
  }
  if (V == 9 || V == 10) {
    uint64_t acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = ((uint64_t)x[2 * j + 1] << 32) | x[2 * j];
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
      for (int r = 0; r < REPS; r++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if (V == 9)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"((uint32_t)acc[(j + 1) & 7]), "r"(b));
          else
            asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(acc[j]) : "r"((uint32_t)acc[(j + 1) & 7] ^ (uint32_t)(acc[j] >> 32)), "r"(b));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) { x[2 * j] = (uint32_t)acc[j]; x[2 * j + 1] = (uint32_t)(acc[j] >> 32); }
  }
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 16; j++) s ^= x[j];
  if (s == 0x1234567) out[0] = s + a;
}

template <int V>
double run(const char* name, double ops_per_rep, int sms, double clk_ghz) {
  uint32_t* d; cudaMalloc(&d, 4);
  cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1);
  int grid = sms * 8;
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(t0);
    k<V><<<grid, 256>>>(d, 12345u + rep);
    cudaEventRecord(t1); cudaEventSynchronize(t1);
    float ms; cudaEventElapsedTime(&ms, t0, t1);
    if (rep && ms < best) best = ms;
  }
  double ops = (double)grid * 256 * ITERS * REPS * ops_per_rep;
  double rate = ops / (best * 1e-3);
  printf("%-44s %8.3f ms  %9.1f Gop/s  %6.2f op/clk/SM @%.2f GHz\n", name, best, rate / 1e9, rate / sms / (clk_ghz * 1e9), clk_ghz);
  cudaFree(d);
  return rate;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; double clk = p.clockRate / 1e6;
  printf("%s, %d SMs, clockRate %.3f GHz\n", p.name, sms, clk);
  run<0>("IMAD.WIDE.U32 plain (8 indep)", 8, sms, clk);
  run<1>("IMAD.WIDE.U32 carry chains 2x4 (.cc/.X)", 8, sms, clk);
  run<2>("IMAD (mad.lo) 8 indep", 8, sms, clk);
  run<3>("IMAD.HI (mad.hi) 8 indep", 8, sms, clk);
  run<4>("IMAD.WIDE carry-out only pairs", 8, sms, clk);
  run<5>("IADD3 carry chain (8 adds)", 8, sms, clk);
  run<6>("4 IMAD.WIDE + 8 ALU interleaved (count 4)", 4, sms, clk);
  run<7>("64-bit add (8 indep)", 8, sms, clk);
  run<8>("add.cc + IMAD.WIDE.X carry-in (count 8)", 8, sms, clk);
  run<11>("IMAD.WIDE carry-out pairs, per-thread data", 8, sms, clk);
  run<4>("IMAD.WIDE carry-out only pairs (again)", 8, sms, clk);
  run<9>("IMAD.WIDE.U32 64-bit addend (u64 accs)", 8, sms, clk);
  run<10>("mul.wide.u32 (no addend) + xor", 8, sms, clk);
  return 0;
}
