// Batch codec / Elligator / scalar-mul kernels (one element per thread) and the
// C ABI of include/decaf377_b200.h.  Reference items replaced are cited per
// kernel; the device arithmetic lives in fq.cuh / isqrt.cuh / point.cuh.
#include <cstdarg>
#include <cstring>
#include <vector>

#include "engine.h"
#include "point.cuh"

namespace d377 {

// ---------------------------------------------------------------------------
// engine plumbing
// ---------------------------------------------------------------------------
static thread_local std::string g_err;

Engine& engine() {
  static Engine e;
  return e;
}

void set_error(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
  return D377_ERR_CUDA;
}

int ensure(DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return D377_OK;
  if (b.p) D377_CUDA(cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  size_t cap = bytes + (bytes >> 3) + 256;
  D377_CUDA(cudaMalloc(&b.p, cap));
  b.cap = cap;
  return D377_OK;
}

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
constexpr int kCodecBlock = 128;  // 8 isqrt slots * 32 B * 128 = 32 KB shared / CTA

// Encoding::vartime_decompress, ark_curve/encoding.rs:32-83
__global__ void __launch_bounds__(kCodecBlock)
k_decompress(const uint8_t* __restrict__ enc, size_t n, uint8_t* __restrict__ out,
             uint8_t* __restrict__ ok) {
  extern __shared__ uint32_t smem[];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  isqrt_smem_t sm = isqrt_smem(smem);
  fq_t s = fq_load(enc + 32 * i);
  pt_t p;
  bool good = pt_decompress(p, s, sm);
  p = pt_select(good, p, pt_identity());
  pt_store(out + 128 * i, p);
  if (ok) ok[i] = good ? 1 : 0;
}

// Element::vartime_compress, ark_curve/encoding.rs:116-128
__global__ void __launch_bounds__(kCodecBlock)
k_compress(const uint8_t* __restrict__ in, size_t n, uint8_t* __restrict__ enc) {
  extern __shared__ uint32_t smem[];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  isqrt_smem_t sm = isqrt_smem(smem);
  pt_t p = pt_load(in + 128 * i);
  fq_store(enc + 32 * i, pt_compress_to_field(p, sm));
}

// Element::encode_to_curve / hash_to_curve, ark_curve/elligator.rs:67-76
template <bool kHash, bool kEncode>
__global__ void __launch_bounds__(kCodecBlock)
k_elligator(const uint8_t* __restrict__ r1, const uint8_t* __restrict__ r2, size_t n,
            uint8_t* __restrict__ out) {
  extern __shared__ uint32_t smem[];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  isqrt_smem_t sm = isqrt_smem(smem);
  // from_le_bytes_mod_order on 32 bytes == to_mont of the raw 256-bit value
  fq_t a = fq_mul(fq_const(FQ_R2), fq_load(r1 + 32 * i));
  pt_t p = pt_elligator(a, sm);
  if (kHash) {
    fq_t b = fq_mul(fq_const(FQ_R2), fq_load(r2 + 32 * i));
    pt_t q = pt_elligator(b, sm);
    p = pt_add(p, q);
  }
  if (kEncode)
    fq_store(out + 32 * i, pt_compress_to_field(p, sm));
  else
    pt_store(out + 128 * i, p);
}

// &Element * &Fr, ark_curve/ops/projective.rs:106-191
template <int kFmt, bool kEncode>
__global__ void __launch_bounds__(kCodecBlock)
k_scalar_mul(const uint8_t* __restrict__ pts, const uint8_t* __restrict__ scalars, size_t n,
             uint8_t* __restrict__ out, uint8_t* __restrict__ ok) {
  extern __shared__ uint32_t smem[];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  isqrt_smem_t sm = isqrt_smem(smem);
  pt_t p;
  if (kFmt == D377_PT_ELEMENT) {
    p = pt_load(pts + 128 * i);
  } else if (kFmt == D377_PT_AFFINE) {
    p.x = fq_load(pts + 64 * i);
    p.y = fq_load(pts + 64 * i + 32);
    p.z = fq_one();
    p.t = fq_mul(p.x, p.y);
  } else {
    bool good = pt_decompress(p, fq_load(pts + 32 * i), sm);
    p = pt_select(good, p, pt_identity());
    if (ok) ok[i] = good ? 1 : 0;
  }
  fq_t k = fq_load(scalars + 32 * i);
  pt_t r = pt_scalar_mul(p, k);
  if (kEncode)
    fq_store(out + 32 * i, pt_compress_to_field(r, sm));
  else
    pt_store(out + 128 * i, r);
}

__global__ void k_add(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n,
                      uint8_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pt_store(out + 128 * i, pt_add(pt_load(a + 128 * i), pt_load(b + 128 * i)));
}

// PartialEq, element/projective.rs:65-70: x1*y2 == x2*y1
__global__ void k_eq(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, size_t n,
                     uint8_t* __restrict__ eq) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fq_t x1 = fq_load(a + 128 * i), y1 = fq_load(a + 128 * i + 32);
  fq_t x2 = fq_load(b + 128 * i), y2 = fq_load(b + 128 * i + 32);
  eq[i] = fq_eq(fq_mul(x1, y2), fq_mul(x2, y1)) ? 1 : 0;
}

// field-layer test entry (rows a2-a4)
__global__ void k_fq_op(int op, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                        size_t n, uint8_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fq_t x = fq_load(a + 32 * i);
  fq_t y = b ? fq_load(b + 32 * i) : fq_zero();
  fq_t r;
  switch (op) {
    case 0: r = fq_mul(x, y); break;
    case 1: r = fq_sqr(x); break;
    case 2: r = fq_add(x, y); break;
    case 3: r = fq_sub(x, y); break;
    case 4: r = fq_neg(x); break;
    case 5: r = fq_to_mont(x); break;
    case 6: r = fq_from_mont(x); break;
    default: r = fq_mul(fq_const(FQ_R2), x); break;
  }
  fq_store(out + 32 * i, r);
}

__global__ void __launch_bounds__(kCodecBlock)
k_fq_isqrt(const uint8_t* __restrict__ x, size_t n, uint8_t* __restrict__ out,
           uint8_t* __restrict__ wsq) {
  extern __shared__ uint32_t smem[];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  isqrt_smem_t sm = isqrt_smem(smem);
  fq_t r;
  bool s = fq_isqrt(r, fq_load(x + 32 * i), sm);
  fq_store(out + 32 * i, r);
  wsq[i] = s ? 1 : 0;
}

// ---- fixed-base tables ------------------------------------------------------
// T[w][j] = (j+1) * 2^(16 w) * G in cached affine form, w < 16, j < 2^15.
constexpr int kFbC = 16;
constexpr int kFbW = 16;
constexpr int kFbK = 1 << (kFbC - 1);

D377_DI fq_t fq_inv(const fq_t& x) {
  // x^(q-2), plain MSB-first square-and-multiply; table building only.
  const uint32_t e[8] = {0xffffffffu, Q1 - 1u, Q2, Q3, Q4, Q5, Q6, Q7};  // q - 2
  fq_t acc = fq_one();
#pragma unroll 1
  for (int i = 252; i >= 0; i--) {
    acc = fq_sqr(acc);
    uint32_t w = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) w = (i >> 5) == j ? e[j] : w;
    if ((w >> (i & 31)) & 1u) acc = fq_mul(acc, x);
  }
  return acc;
}

__global__ void k_fb_bases(pt_t* bases) {
  pt_t p;
  p.x = fq_const(FQ_BX);
  p.y = fq_const(FQ_BY);
  p.z = fq_one();
  p.t = fq_const(FQ_BT);
  for (int w = 0; w < kFbW; w++) {
    bases[w] = p;
    for (int k = 0; k < kFbC; k++) p = pt_dbl(p);
  }
}

__global__ void k_fb_fill(const pt_t* __restrict__ bases, niels_t* __restrict__ table) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)kFbW * kFbK) return;
  int w = (int)(idx / kFbK);
  uint32_t m = (uint32_t)(idx % kFbK) + 1;
  pt_t base = bases[w];
  pt_t acc = pt_identity();
#pragma unroll 1
  for (int i = kFbC - 1; i >= 0; i--) {
    acc = pt_dbl(acc);
    if ((m >> i) & 1u) acc = pt_add(acc, base);
  }
  fq_t zi = fq_inv(acc.z);
  table[idx] = niels_from_affine(fq_mul(acc.x, zi), fq_mul(acc.y, zi));
}

D377_DI niels_t niels_load(const niels_t* p) {
  const uint4* v = reinterpret_cast<const uint4*>(p);
  uint4 q[6];
#pragma unroll
  for (int i = 0; i < 6; i++) q[i] = __ldg(v + i);
  niels_t n;
  n.ymx.l[0] = q[0].x; n.ymx.l[1] = q[0].y; n.ymx.l[2] = q[0].z; n.ymx.l[3] = q[0].w;
  n.ymx.l[4] = q[1].x; n.ymx.l[5] = q[1].y; n.ymx.l[6] = q[1].z; n.ymx.l[7] = q[1].w;
  n.ypx.l[0] = q[2].x; n.ypx.l[1] = q[2].y; n.ypx.l[2] = q[2].z; n.ypx.l[3] = q[2].w;
  n.ypx.l[4] = q[3].x; n.ypx.l[5] = q[3].y; n.ypx.l[6] = q[3].z; n.ypx.l[7] = q[3].w;
  n.kt.l[0] = q[4].x; n.kt.l[1] = q[4].y; n.kt.l[2] = q[4].z; n.kt.l[3] = q[4].w;
  n.kt.l[4] = q[5].x; n.kt.l[5] = q[5].y; n.kt.l[6] = q[5].z; n.kt.l[7] = q[5].w;
  return n;
}

// Element::GENERATOR * s with signed 16-bit windows over the table above.
template <bool kEncode>
__global__ void __launch_bounds__(kCodecBlock)
k_fixed_base(const niels_t* __restrict__ table, const uint8_t* __restrict__ scalars, size_t n,
             uint8_t* __restrict__ out) {
  extern __shared__ uint32_t smem[];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fq_t k = fq_load(scalars + 32 * i);
  pt_t acc = pt_identity();
  uint32_t carry = 0;
#pragma unroll 1
  for (int w = 0; w < kFbW; w++) {
    uint32_t limb = k.l[w >> 1];
    uint32_t raw = ((w & 1) ? (limb >> 16) : (limb & 0xffffu)) + carry;
    carry = raw > (uint32_t)kFbK ? 1u : 0u;
    int32_t d = (int32_t)raw - (int32_t)(carry << kFbC);
    uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
    niels_t nl = niels_identity();
    if (mag) nl = niels_cneg(niels_load(table + (size_t)w * kFbK + (mag - 1)), d < 0);
    acc = pt_add_niels(acc, nl);
  }
  // a carry out of the top window only happens for scalars >= 2^255 (never canonical)
  if (kEncode) {
    isqrt_smem_t sm = isqrt_smem(smem);
    fq_store(out + 32 * i, pt_compress_to_field(acc, sm));
  } else {
    pt_store(out + 128 * i, acc);
  }
}

// ---- IMAD.WIDE issue-rate microbenchmark --------------------------------
// Eight independent IMAD.WIDE.U32 (with carry-out, the form fq_mul issues) per
// step; every multiplicand is another accumulator's limb so that ptxas can
// neither hoist nor strength-reduce the products.  Counts 8 wide multiply-adds
// per step; 2048 resident threads per SM.  The achieved rate depends on the
// register-bank pattern ptxas happens to pick for the four source registers, so
// a few operand arrangements are timed and the best one is reported.
template <int kVariant>
__global__ void __launch_bounds__(256) k_imad_peak(uint32_t* out, uint32_t seed, int iters) {
  uint32_t b0 = seed * 3 + threadIdx.x * 5 + 7, b1 = b0 ^ 0x5bd1e995u;
  uint32_t x[16];
#pragma unroll
  for (int j = 0; j < 16; j++) x[j] = j * seed + threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const int src = kVariant == 0 ? ((2 * j + 3) & 15) : kVariant == 1 ? ((2 * j + 2) & 15)
                                                                          : ((2 * j + 5) & 15);
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                     : "+r"(x[2 * j]), "+r"(x[2 * j + 1])
                     : "r"(x[src]), "r"(kVariant == 3 ? ((j & 1) ? b1 : b0) : b0));
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 16; j++) s ^= x[j];
  if (s == 0x1234567u) out[0] = s + b1;
}

// ---------------------------------------------------------------------------
// host side of the API
// ---------------------------------------------------------------------------
static size_t codec_smem() { return ISQRT_SMEM_WORDS(kCodecBlock) * sizeof(uint32_t); }

static int check_fmt(int f) { return f == D377_OUT_ELEMENT || f == D377_OUT_ENCODING; }
static size_t pt_bytes(int fmt) {
  return fmt == D377_PT_ELEMENT ? 128 : fmt == D377_PT_ENCODING ? 32 : 64;
}
static size_t out_bytes(int fmt) { return fmt == D377_OUT_ENCODING ? 32 : 128; }

static int ensure_fb_table() {
  Engine& e = engine();
  if (e.fb_table) return D377_OK;
  pt_t* bases = nullptr;
  niels_t* table = nullptr;
  D377_CUDA(cudaMalloc(&bases, sizeof(pt_t) * kFbW));
  D377_CUDA(cudaMalloc(&table, sizeof(niels_t) * (size_t)kFbW * kFbK));
  k_fb_bases<<<1, 1, 0, e.stream>>>(bases);
  D377_LAUNCHED();
  k_fb_fill<<<grid_for((size_t)kFbW * kFbK, 128), 128, 0, e.stream>>>(bases, table);
  D377_LAUNCHED();
  D377_CUDA(cudaGetLastError());
  D377_CUDA(cudaStreamSynchronize(e.stream));
  D377_CUDA(cudaFree(bases));
  e.fb_table = table;
  return D377_OK;
}

}  // namespace d377

using namespace d377;

#define LOCK() std::lock_guard<std::recursive_mutex> _lk(engine().mu)

extern "C" {

int d377_init(int device) {
  Engine& e = engine();
  LOCK();
  if (e.ready && e.device == device) return D377_OK;
  if (e.ready) d377_shutdown();
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0) {
    set_error("no CUDA device available (%s); decaf377_b200 has no CPU fallback",
              ce == cudaSuccess ? "device count is 0" : cudaGetErrorString(ce));
    return D377_ERR_CUDA;
  }
  if (device < 0 || device >= count) {
    set_error("device %d out of range (have %d)", device, count);
    return D377_ERR_INVALID_ARG;
  }
  D377_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  D377_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major,
              prop.minor);
    return D377_ERR_CUDA;
  }
  e.sm_count = prop.multiProcessorCount;
  D377_CUDA(cudaStreamCreateWithFlags(&e.stream, cudaStreamNonBlocking));
  // d_small / h_small layout: [0,160) result of the synchronous calls, [512,640) tmp,
  // [2048,4096) MSM chunk partials, [4096,4100) status word, [4352 + 256 k, ...) slot k
  D377_CUDA(cudaMalloc(&e.d_small, 8192));
  D377_CUDA(cudaMallocHost(&e.h_small, 8192));
  D377_CUDA(cudaStreamCreateWithFlags(&e.copy_stream, cudaStreamNonBlocking));
  for (int k = 0; k < Engine::kSlots; k++) {
    D377_CUDA(cudaEventCreateWithFlags(&e.ev_h2d[k], cudaEventDisableTiming));
    D377_CUDA(cudaEventCreateWithFlags(&e.ev_done[k], cudaEventDisableTiming));
    e.slot_busy[k] = false;
  }
  e.device = device;
  e.ready = true;
  e.launches = 0;
  return D377_OK;
}

int d377_shutdown(void) {
  Engine& e = engine();
  LOCK();
  if (!e.ready) return D377_OK;
  cudaSetDevice(e.device);
  cudaStreamSynchronize(e.stream);
  for (DevBuf* b : {&e.in0, &e.in1, &e.out0, &e.out1, &e.msm_ws, &e.slot_sc[0], &e.slot_sc[1],
                    &e.slot_pt[0], &e.slot_pt[1]}) {
    if (b->p) cudaFree(b->p);
    b->p = nullptr;
    b->cap = 0;
  }
  if (e.fb_table) cudaFree(e.fb_table);
  e.fb_table = nullptr;
  if (e.d_small) cudaFree(e.d_small);
  if (e.h_small) cudaFreeHost(e.h_small);
  e.d_small = e.h_small = nullptr;
  for (int k = 0; k < Engine::kSlots; k++) {
    if (e.ev_h2d[k]) cudaEventDestroy(e.ev_h2d[k]);
    if (e.ev_done[k]) cudaEventDestroy(e.ev_done[k]);
    e.ev_h2d[k] = e.ev_done[k] = nullptr;
  }
  if (e.copy_stream) cudaStreamDestroy(e.copy_stream);
  e.copy_stream = nullptr;
  cudaStreamDestroy(e.stream);
  e.stream = nullptr;
  e.ready = false;
  return D377_OK;
}

void* d377_stream(void) { return (void*)engine().stream; }

int d377_sync(void) {
  D377_REQUIRE_READY();
  D377_CUDA(cudaStreamSynchronize(engine().stream));
  return D377_OK;
}

const char* d377_last_error(void) { return g_err.c_str(); }

uint64_t d377_launch_count(void) { return engine().launches.load(); }

int d377_msm_stage_info(float ms[8], int* c, int* W, uint64_t* n) {
  D377_REQUIRE_READY();
  if (!ms) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  return msm_stage_info(ms, c, W, n);
}

int d377_msm_set_window(int c) {
  if (c != 0 && (c < 4 || c > 24)) {
    set_error("window width %d out of range [4, 24]", c);
    return D377_ERR_INVALID_ARG;
  }
  engine().msm_window_override = c;
  return D377_OK;
}

// ---- device-pointer entry points -------------------------------------------

int d377_batch_decompress_dev(const uint8_t* enc, size_t n, uint8_t* elements, uint8_t* ok) {
  D377_REQUIRE_READY();
  if (n == 0) return D377_OK;
  if (!enc || !elements) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  k_decompress<<<grid_for(n, kCodecBlock), kCodecBlock, codec_smem(), e.stream>>>(enc, n, elements, ok);
  D377_LAUNCHED();
  D377_CUDA(cudaGetLastError());
  return D377_OK;
}

int d377_batch_compress_dev(const uint8_t* elements, size_t n, uint8_t* enc) {
  D377_REQUIRE_READY();
  if (n == 0) return D377_OK;
  if (!enc || !elements) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  k_compress<<<grid_for(n, kCodecBlock), kCodecBlock, codec_smem(), e.stream>>>(elements, n, enc);
  D377_LAUNCHED();
  D377_CUDA(cudaGetLastError());
  return D377_OK;
}

int d377_batch_encode_to_curve_dev(const uint8_t* r, size_t n, uint8_t* out, int out_format) {
  D377_REQUIRE_READY();
  if (!check_fmt(out_format)) { set_error("bad out_format %d", out_format); return D377_ERR_INVALID_ARG; }
  if (n == 0) return D377_OK;
  if (!r || !out) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  dim3 g(grid_for(n, kCodecBlock));
  if (out_format == D377_OUT_ENCODING)
    k_elligator<false, true><<<g, kCodecBlock, codec_smem(), e.stream>>>(r, nullptr, n, out);
  else
    k_elligator<false, false><<<g, kCodecBlock, codec_smem(), e.stream>>>(r, nullptr, n, out);
  D377_LAUNCHED();
  D377_CUDA(cudaGetLastError());
  return D377_OK;
}

int d377_batch_hash_to_curve_dev(const uint8_t* r1, const uint8_t* r2, size_t n, uint8_t* out,
                                 int out_format) {
  D377_REQUIRE_READY();
  if (!check_fmt(out_format)) { set_error("bad out_format %d", out_format); return D377_ERR_INVALID_ARG; }
  if (n == 0) return D377_OK;
  if (!r1 || !r2 || !out) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  dim3 g(grid_for(n, kCodecBlock));
  if (out_format == D377_OUT_ENCODING)
    k_elligator<true, true><<<g, kCodecBlock, codec_smem(), e.stream>>>(r1, r2, n, out);
  else
    k_elligator<true, false><<<g, kCodecBlock, codec_smem(), e.stream>>>(r1, r2, n, out);
  D377_LAUNCHED();
  D377_CUDA(cudaGetLastError());
  return D377_OK;
}

int d377_batch_scalar_mul_dev(const uint8_t* points, int point_format, const uint8_t* scalars,
                              size_t n, uint8_t* out, int out_format, uint8_t* ok) {
  D377_REQUIRE_READY();
  if (!check_fmt(out_format) || point_format < 0 || point_format > 2) {
    set_error("bad format (%d, %d)", point_format, out_format);
    return D377_ERR_INVALID_ARG;
  }
  if (n == 0) return D377_OK;
  if (!points || !scalars || !out) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  dim3 g(grid_for(n, kCodecBlock));
  size_t sm = codec_smem();
#define SM_LAUNCH(F, E) k_scalar_mul<F, E><<<g, kCodecBlock, sm, e.stream>>>(points, scalars, n, out, ok)
  bool enc = out_format == D377_OUT_ENCODING;
  switch (point_format) {
    case D377_PT_ELEMENT: if (enc) SM_LAUNCH(D377_PT_ELEMENT, true); else SM_LAUNCH(D377_PT_ELEMENT, false); break;
    case D377_PT_ENCODING: if (enc) SM_LAUNCH(D377_PT_ENCODING, true); else SM_LAUNCH(D377_PT_ENCODING, false); break;
    default: if (enc) SM_LAUNCH(D377_PT_AFFINE, true); else SM_LAUNCH(D377_PT_AFFINE, false); break;
  }
#undef SM_LAUNCH
  D377_LAUNCHED();
  D377_CUDA(cudaGetLastError());
  return D377_OK;
}

int d377_fixed_base_mul_dev(const uint8_t* scalars, size_t n, uint8_t* out, int out_format) {
  D377_REQUIRE_READY();
  if (!check_fmt(out_format)) { set_error("bad out_format %d", out_format); return D377_ERR_INVALID_ARG; }
  if (n == 0) return D377_OK;
  if (!scalars || !out) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  int rc = ensure_fb_table();
  if (rc) return rc;
  Engine& e = engine();
  dim3 g(grid_for(n, kCodecBlock));
  const niels_t* tab = (const niels_t*)e.fb_table;
  if (out_format == D377_OUT_ENCODING)
    k_fixed_base<true><<<g, kCodecBlock, codec_smem(), e.stream>>>(tab, scalars, n, out);
  else
    k_fixed_base<false><<<g, kCodecBlock, codec_smem(), e.stream>>>(tab, scalars, n, out);
  D377_LAUNCHED();
  D377_CUDA(cudaGetLastError());
  return D377_OK;
}

int d377_batch_add_dev(const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
  D377_REQUIRE_READY();
  if (n == 0) return D377_OK;
  if (!a || !b || !out) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  k_add<<<grid_for(n, 128), 128, 0, engine().stream>>>(a, b, n, out);
  D377_LAUNCHED();
  D377_CUDA(cudaGetLastError());
  return D377_OK;
}

int d377_batch_element_eq_dev(const uint8_t* a, const uint8_t* b, size_t n, uint8_t* eq) {
  D377_REQUIRE_READY();
  if (n == 0) return D377_OK;
  if (!a || !b || !eq) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  k_eq<<<grid_for(n, 128), 128, 0, engine().stream>>>(a, b, n, eq);
  D377_LAUNCHED();
  D377_CUDA(cudaGetLastError());
  return D377_OK;
}

int d377_element_sum_dev(const uint8_t* elements, size_t n, uint8_t* out_element,
                         uint8_t* out_encoding) {
  D377_REQUIRE_READY();
  return element_sum_dev(elements, n, out_element, out_encoding);
}

int d377_msm_dev(const uint8_t* scalars, const uint8_t* points, int point_format, size_t n,
                 uint8_t* out_element, uint8_t* out_encoding) {
  D377_REQUIRE_READY();
  return msm_dev(scalars, points, point_format, n, out_element, out_encoding);
}

// ---- host-pointer entry points -------------------------------------------
// Stage through the engine's device buffers; inputs go up and results come
// back on the engine stream, then the call blocks until they have landed.

#define H2D(dst, src, bytes) D377_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, e.stream))
#define D2H(dst, src, bytes) D377_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, e.stream))
#define TRY(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

int d377_batch_decompress(const uint8_t* enc, size_t n, uint8_t* elements, uint8_t* ok) {
  D377_REQUIRE_READY();
  if (n == 0) return D377_OK;
  if (!enc || !elements) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  LOCK();
  TRY(ensure(e.in0, n * 32));
  TRY(ensure(e.out0, n * 128));
  TRY(ensure(e.out1, n));
  H2D(e.in0.p, enc, n * 32);
  TRY(d377_batch_decompress_dev((uint8_t*)e.in0.p, n, (uint8_t*)e.out0.p, (uint8_t*)e.out1.p));
  D2H(elements, e.out0.p, n * 128);
  if (ok) D2H(ok, e.out1.p, n);
  D377_CUDA(cudaStreamSynchronize(e.stream));
  return D377_OK;
}

int d377_batch_compress(const uint8_t* elements, size_t n, uint8_t* enc) {
  D377_REQUIRE_READY();
  if (n == 0) return D377_OK;
  if (!enc || !elements) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  LOCK();
  TRY(ensure(e.in0, n * 128));
  TRY(ensure(e.out0, n * 32));
  H2D(e.in0.p, elements, n * 128);
  TRY(d377_batch_compress_dev((uint8_t*)e.in0.p, n, (uint8_t*)e.out0.p));
  D2H(enc, e.out0.p, n * 32);
  D377_CUDA(cudaStreamSynchronize(e.stream));
  return D377_OK;
}

int d377_batch_encode_to_curve(const uint8_t* r, size_t n, uint8_t* out, int out_format) {
  D377_REQUIRE_READY();
  if (!check_fmt(out_format)) { set_error("bad out_format %d", out_format); return D377_ERR_INVALID_ARG; }
  if (n == 0) return D377_OK;
  if (!r || !out) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  LOCK();
  size_t ob = out_bytes(out_format);
  TRY(ensure(e.in0, n * 32));
  TRY(ensure(e.out0, n * ob));
  H2D(e.in0.p, r, n * 32);
  TRY(d377_batch_encode_to_curve_dev((uint8_t*)e.in0.p, n, (uint8_t*)e.out0.p, out_format));
  D2H(out, e.out0.p, n * ob);
  D377_CUDA(cudaStreamSynchronize(e.stream));
  return D377_OK;
}

int d377_batch_hash_to_curve(const uint8_t* r1, const uint8_t* r2, size_t n, uint8_t* out,
                             int out_format) {
  D377_REQUIRE_READY();
  if (!check_fmt(out_format)) { set_error("bad out_format %d", out_format); return D377_ERR_INVALID_ARG; }
  if (n == 0) return D377_OK;
  if (!r1 || !r2 || !out) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  LOCK();
  size_t ob = out_bytes(out_format);
  TRY(ensure(e.in0, n * 32));
  TRY(ensure(e.in1, n * 32));
  TRY(ensure(e.out0, n * ob));
  H2D(e.in0.p, r1, n * 32);
  H2D(e.in1.p, r2, n * 32);
  TRY(d377_batch_hash_to_curve_dev((uint8_t*)e.in0.p, (uint8_t*)e.in1.p, n, (uint8_t*)e.out0.p, out_format));
  D2H(out, e.out0.p, n * ob);
  D377_CUDA(cudaStreamSynchronize(e.stream));
  return D377_OK;
}

int d377_batch_scalar_mul(const uint8_t* points, int point_format, const uint8_t* scalars,
                          size_t n, uint8_t* out, int out_format, uint8_t* ok) {
  D377_REQUIRE_READY();
  if (!check_fmt(out_format) || point_format < 0 || point_format > 2) {
    set_error("bad format (%d, %d)", point_format, out_format);
    return D377_ERR_INVALID_ARG;
  }
  if (n == 0) return D377_OK;
  if (!points || !scalars || !out) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  LOCK();
  size_t pb = pt_bytes(point_format), ob = out_bytes(out_format);
  TRY(ensure(e.in0, n * pb));
  TRY(ensure(e.in1, n * 32));
  TRY(ensure(e.out0, n * ob));
  TRY(ensure(e.out1, n));
  H2D(e.in0.p, points, n * pb);
  H2D(e.in1.p, scalars, n * 32);
  D377_CUDA(cudaMemsetAsync(e.out1.p, 1, n, e.stream));
  TRY(d377_batch_scalar_mul_dev((uint8_t*)e.in0.p, point_format, (uint8_t*)e.in1.p, n,
                                (uint8_t*)e.out0.p, out_format, (uint8_t*)e.out1.p));
  D2H(out, e.out0.p, n * ob);
  if (ok) D2H(ok, e.out1.p, n);
  D377_CUDA(cudaStreamSynchronize(e.stream));
  return D377_OK;
}

int d377_fixed_base_mul(const uint8_t* scalars, size_t n, uint8_t* out, int out_format) {
  D377_REQUIRE_READY();
  if (!check_fmt(out_format)) { set_error("bad out_format %d", out_format); return D377_ERR_INVALID_ARG; }
  if (n == 0) return D377_OK;
  if (!scalars || !out) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  LOCK();
  size_t ob = out_bytes(out_format);
  TRY(ensure(e.in0, n * 32));
  TRY(ensure(e.out0, n * ob));
  H2D(e.in0.p, scalars, n * 32);
  TRY(d377_fixed_base_mul_dev((uint8_t*)e.in0.p, n, (uint8_t*)e.out0.p, out_format));
  D2H(out, e.out0.p, n * ob);
  D377_CUDA(cudaStreamSynchronize(e.stream));
  return D377_OK;
}

int d377_batch_add(const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
  D377_REQUIRE_READY();
  if (n == 0) return D377_OK;
  if (!a || !b || !out) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  LOCK();
  TRY(ensure(e.in0, n * 128));
  TRY(ensure(e.in1, n * 128));
  TRY(ensure(e.out0, n * 128));
  H2D(e.in0.p, a, n * 128);
  H2D(e.in1.p, b, n * 128);
  TRY(d377_batch_add_dev((uint8_t*)e.in0.p, (uint8_t*)e.in1.p, n, (uint8_t*)e.out0.p));
  D2H(out, e.out0.p, n * 128);
  D377_CUDA(cudaStreamSynchronize(e.stream));
  return D377_OK;
}

int d377_batch_element_eq(const uint8_t* a, const uint8_t* b, size_t n, uint8_t* eq) {
  D377_REQUIRE_READY();
  if (n == 0) return D377_OK;
  if (!a || !b || !eq) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  LOCK();
  TRY(ensure(e.in0, n * 128));
  TRY(ensure(e.in1, n * 128));
  TRY(ensure(e.out0, n));
  H2D(e.in0.p, a, n * 128);
  H2D(e.in1.p, b, n * 128);
  TRY(d377_batch_element_eq_dev((uint8_t*)e.in0.p, (uint8_t*)e.in1.p, n, (uint8_t*)e.out0.p));
  D2H(eq, e.out0.p, n);
  D377_CUDA(cudaStreamSynchronize(e.stream));
  return D377_OK;
}

static int small_results_back(uint8_t* out_element, uint8_t* out_encoding) {
  Engine& e = engine();
  D2H(e.h_small, e.d_small, 160);
  D377_CUDA(cudaStreamSynchronize(e.stream));
  if (out_element) memcpy(out_element, e.h_small, 128);
  if (out_encoding) memcpy(out_encoding, e.h_small + 128, 32);
  return D377_OK;
}

int d377_element_sum(const uint8_t* elements, size_t n, uint8_t out_element[128],
                     uint8_t out_encoding[32]) {
  D377_REQUIRE_READY();
  if (n && !elements) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  LOCK();
  TRY(ensure(e.in0, n * 128 + 128));
  if (n) H2D(e.in0.p, elements, n * 128);
  TRY(element_sum_dev((uint8_t*)e.in0.p, n, e.d_small, e.d_small + 128));
  return small_results_back(out_element, out_encoding);
}

int d377_msm_submit(const uint8_t* scalars, const uint8_t* points, int point_format, size_t n,
                    int slot) {
  D377_REQUIRE_READY();
  if (point_format < 0 || point_format > 2) { set_error("bad point_format %d", point_format); return D377_ERR_INVALID_ARG; }
  if (slot < 0 || slot >= Engine::kSlots) { set_error("slot %d out of range", slot); return D377_ERR_INVALID_ARG; }
  if (n && (!scalars || !points)) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  LOCK();
  if (e.slot_busy[slot]) { set_error("slot %d still in flight: call d377_msm_wait first", slot); return D377_ERR_INVALID_ARG; }
  size_t pb = pt_bytes(point_format);
  TRY(ensure(e.slot_sc[slot], n * 32 + 32));
  TRY(ensure(e.slot_pt[slot], n * pb + 128));
  uint8_t* dres = e.d_small + 4352 + 256 * slot;
  if (n) {
    // inputs go up on the copy stream so that they overlap the MSM of the other slot
    D377_CUDA(cudaMemcpyAsync(e.slot_sc[slot].p, scalars, n * 32, cudaMemcpyHostToDevice, e.copy_stream));
    D377_CUDA(cudaMemcpyAsync(e.slot_pt[slot].p, points, n * pb, cudaMemcpyHostToDevice, e.copy_stream));
  }
  D377_CUDA(cudaEventRecord(e.ev_h2d[slot], e.copy_stream));
  D377_CUDA(cudaStreamWaitEvent(e.stream, e.ev_h2d[slot], 0));
  D377_CUDA(cudaMemsetAsync(dres + 192, 0, 4, e.stream));
  TRY(msm_enqueue((uint8_t*)e.slot_sc[slot].p, (uint8_t*)e.slot_pt[slot].p, point_format, n, dres,
                  dres + 128, (uint32_t*)(dres + 192)));
  D377_CUDA(cudaMemcpyAsync(e.h_small + 4352 + 256 * slot, dres, 256, cudaMemcpyDeviceToHost, e.stream));
  D377_CUDA(cudaEventRecord(e.ev_done[slot], e.stream));
  e.slot_busy[slot] = true;
  return D377_OK;
}

int d377_msm_wait(int slot, uint8_t out_element[128], uint8_t out_encoding[32]) {
  D377_REQUIRE_READY();
  if (slot < 0 || slot >= Engine::kSlots) { set_error("slot %d out of range", slot); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  LOCK();
  if (!e.slot_busy[slot]) { set_error("slot %d has no MSM in flight", slot); return D377_ERR_INVALID_ARG; }
  D377_CUDA(cudaEventSynchronize(e.ev_done[slot]));
  e.slot_busy[slot] = false;
  const uint8_t* h = e.h_small + 4352 + 256 * slot;
  uint32_t flags;
  memcpy(&flags, h + 192, 4);
  TRY(msm_check_flags(flags));
  if (out_element) memcpy(out_element, h, 128);
  if (out_encoding) memcpy(out_encoding, h + 128, 32);
  return D377_OK;
}

int d377_msm(const uint8_t* scalars, const uint8_t* points, int point_format, size_t n,
             uint8_t out_element[128], uint8_t out_encoding[32]) {
  D377_REQUIRE_READY();
  Engine& e = engine();
  LOCK();
  int slot = e.slot_busy[0] ? 1 : 0;
  TRY(d377_msm_submit(scalars, points, point_format, n, slot));
  return d377_msm_wait(slot, out_element, out_encoding);
}

int d377_fq_batch_op(int op, const uint8_t* a, const uint8_t* b, size_t n, uint8_t* out) {
  D377_REQUIRE_READY();
  if (op < 0 || op > 7) { set_error("bad op %d", op); return D377_ERR_INVALID_ARG; }
  if (n == 0) return D377_OK;
  if (!a || !out) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  bool binary = (op == 0 || op == 2 || op == 3);
  if (binary && !b) { set_error("op %d needs b", op); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  LOCK();
  TRY(ensure(e.in0, n * 32));
  TRY(ensure(e.in1, n * 32));
  TRY(ensure(e.out0, n * 32));
  H2D(e.in0.p, a, n * 32);
  if (binary) H2D(e.in1.p, b, n * 32);
  k_fq_op<<<grid_for(n, 128), 128, 0, e.stream>>>(op, (uint8_t*)e.in0.p,
                                                  binary ? (uint8_t*)e.in1.p : nullptr, n,
                                                  (uint8_t*)e.out0.p);
  D377_LAUNCHED();
  D377_CUDA(cudaGetLastError());
  D2H(out, e.out0.p, n * 32);
  D377_CUDA(cudaStreamSynchronize(e.stream));
  return D377_OK;
}

int d377_fq_batch_isqrt(const uint8_t* x, size_t n, uint8_t* out, uint8_t* was_square) {
  D377_REQUIRE_READY();
  if (n == 0) return D377_OK;
  if (!x || !out || !was_square) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  LOCK();
  TRY(ensure(e.in0, n * 32));
  TRY(ensure(e.out0, n * 32));
  TRY(ensure(e.out1, n));
  H2D(e.in0.p, x, n * 32);
  k_fq_isqrt<<<grid_for(n, kCodecBlock), kCodecBlock, codec_smem(), e.stream>>>(
      (uint8_t*)e.in0.p, n, (uint8_t*)e.out0.p, (uint8_t*)e.out1.p);
  D377_LAUNCHED();
  D377_CUDA(cudaGetLastError());
  D2H(out, e.out0.p, n * 32);
  D2H(was_square, e.out1.p, n);
  D377_CUDA(cudaStreamSynchronize(e.stream));
  return D377_OK;
}

int d377_imad_peak(double* gimad_per_s) {
  D377_REQUIRE_READY();
  if (!gimad_per_s) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  Engine& e = engine();
  LOCK();
  const int iters = 2048, block = 256;
  const int grid = e.sm_count * 8;
  cudaEvent_t t0, t1;
  D377_CUDA(cudaEventCreate(&t0));
  D377_CUDA(cudaEventCreate(&t1));
  double best = 0;
  for (int rep = 0; rep < 12; rep++) {
    D377_CUDA(cudaEventRecord(t0, e.stream));
    switch (rep & 3) {
      case 0: k_imad_peak<0><<<grid, block, 0, e.stream>>>((uint32_t*)e.d_small, 12345u + rep, iters); break;
      case 1: k_imad_peak<1><<<grid, block, 0, e.stream>>>((uint32_t*)e.d_small, 12345u + rep, iters); break;
      case 2: k_imad_peak<2><<<grid, block, 0, e.stream>>>((uint32_t*)e.d_small, 12345u + rep, iters); break;
      default: k_imad_peak<3><<<grid, block, 0, e.stream>>>((uint32_t*)e.d_small, 12345u + rep, iters); break;
    }
    D377_CUDA(cudaEventRecord(t1, e.stream));
    D377_CUDA(cudaEventSynchronize(t1));
    float ms = 0;
    D377_CUDA(cudaEventElapsedTime(&ms, t0, t1));
    double ops = (double)grid * block * iters * 64.0;
    double g = ops / (ms * 1e-3) / 1e9;
    if (rep >= 4 && g > best) best = g;
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  *gimad_per_s = best;
  return D377_OK;
}

}  // extern "C"
