// Pippenger multiscalar multiplication  Q = sum_i s_i * P_i  on one GPU.
//
// Replaces Element::vartime_multiscalar_mul (reference
// src/ark_curve/element/projective.rs:99-117, a serial fold of scalar-muls) and
// the ark-ec default `VariableBaseMSM::msm` the reference inherits at
// src/ark_curve/element.rs:37 (not in the reference tree).  The result is the
// same group element; tests compare its 32-byte encoding with the oracle's.
//
// Pipeline (no host synchronisation inside; the scalar side 2-4 runs on a second stream,
// window group by window group, under the accumulation of the group before):
//   1. k_msm_normalize / k_msm_points*   input points -> 128-byte bucket operands
//                      (affine (y-x, y+x, 2dxy, -2dxy), or cached projective for small n)
//   2. k_msm_count     scalars -> signed c-bit digits; per-(window,|digit|) bucket
//                      histogram with fire-and-forget atomics
//   3. scan            bucket counts -> offsets
//   4. k_msm_scatter   counting-sort scatter of (point index | sign) by bucket,
//                      window-major so the target stays in L2
//   5. k_msm_accumulate  every thread adds a fixed-length run of the sorted list
//                      (load balance independent of the scalar distribution);
//                      runs covering a whole bucket store the bucket sum, pieces
//                      of buckets that straddle threads are stitched by
//   6. k_msm_seg_reduce / k_msm_seg_reduce_warp (log-depth, so one huge bucket costs no
//                      serial walk)
//   7. k_msm_bucket_reduce_ap  per 16-bucket segment: running sums (A, P)
//   8. k_wsum_tree     weighted tree over the segments -> one sum per window;
//      k_finish        Horner over the windows, optional fused compress.
#include <algorithm>
#include <cmath>
#include <vector>

#include "engine.h"
#include "point.cuh"

namespace d377 {

D377_DI cached_t cached_load(const cached_t* p) {
  const uint8_t* b = reinterpret_cast<const uint8_t*>(p);
  cached_t c;
  c.ymx = fq_load(b);
  c.ypx = fq_load(b + 32);
  c.kt = fq_load(b + 64);
  c.z2 = fq_load(b + 96);
  return c;
}

D377_DI void cached_store(cached_t* p, const cached_t& c) {
  uint8_t* b = reinterpret_cast<uint8_t*>(p);
  fq_store(b, c.ymx);
  fq_store(b + 32, c.ypx);
  fq_store(b + 64, c.kt);
  fq_store(b + 96, c.z2);
}

// Affine bucket operand as the accumulation kernel reads it: 128 bytes, one cache line,
// (y-x | y+x | 2d*x*y | -2d*x*y), all canonical.  Adding -P reads the first pair in the
// opposite order and the fourth field instead of the third, so the sign of a bucket entry
// costs no instruction beyond the address computation; and a 128-byte aligned record is one
// DRAM fetch, where a 96-byte one straddles two 128-byte lines half of the time (ncu:
// 1.97x the algorithmic bytes with 96-byte records).
struct aff4_t {
  fq_r ymx, ypx, kt, nkt;
};

D377_DI void aff4_store(aff4_t* p, const niels_t& c) {
  uint8_t* b = reinterpret_cast<uint8_t*>(p);
  fq_store(b, c.ymx);
  fq_store(b + 32, c.ypx);
  fq_store(b + 64, c.kt);
  fq_store(b + 96, fq_reduce(fq_neg(c.kt)));
}

D377_DI pt_t ptv_load(const pt_t* p) { return pt_load(reinterpret_cast<const uint8_t*>(p)); }
D377_DI void ptv_store(pt_t* p, const pt_t& v) { pt_store(reinterpret_cast<uint8_t*>(p), v); }

// ---- 1. points -> cached ---------------------------------------------------
constexpr int kBlk = 128;

template <int kFmt>
__global__ void __launch_bounds__(kBlk)
k_msm_points(const uint8_t* __restrict__ pts, size_t n, cached_t* __restrict__ out,
             uint32_t* __restrict__ flags) {
  extern __shared__ uint32_t smem[];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pt_t p;
  if (kFmt == D377_PT_ELEMENT) {
    p = pt_load_wire(pts + 128 * i);
  } else if (kFmt == D377_PT_XYZ) {
    // (X : Y : Z) without T: (XZ : YZ : Z^2 : XY) is the same point in extended coordinates
    fq_t x = fq_load_wire(pts + 96 * i), y = fq_load_wire(pts + 96 * i + 32), z = fq_load_wire(pts + 96 * i + 64);
    p.x = fq_mul(x, z);
    p.y = fq_mul(y, z);
    p.z = fq_sqr(z);
    p.t = fq_mul(x, y);
  } else if (kFmt == D377_PT_AFFINE) {
    p.x = fq_load_wire(pts + 64 * i);
    p.y = fq_load_wire(pts + 64 * i + 32);
    p.z = fq_one();
    p.t = fq_mul(p.x, p.y);
  } else {
    isqrt_smem_t sm = isqrt_smem(smem);
    bool good = pt_decompress(p, fq_load_raw(pts + 32 * i), sm);
    if (!good) {
      atomicOr(flags, 2u);
      p = pt_identity();
    }
  }
  cached_store(out + i, cached_from(p));
}

// Inputs that already have Z = 1 (AffinePoint, Encoding) -> canonical (y-x, y+x, 2d*x*y),
// 96 B each: the bucket additions then cost 7 multiplications instead of 8.
template <int kFmt>
__global__ void __launch_bounds__(kBlk)
k_msm_points_affine(const uint8_t* __restrict__ pts, size_t n, aff4_t* __restrict__ out,
                    uint32_t* __restrict__ flags) {
  extern __shared__ uint32_t smem[];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fq_t x, y;
  if (kFmt == D377_PT_AFFINE) {
    x = fq_load_wire(pts + 64 * i);
    y = fq_load_wire(pts + 64 * i + 32);
  } else {
    isqrt_smem_t sm = isqrt_smem(smem);
    pt_t p;
    bool good = pt_decompress(p, fq_load_raw(pts + 32 * i), sm);
    if (!good) {
      atomicOr(flags, 2u);
      p = pt_identity();
    }
    x = p.x;
    y = p.y;
  }
  aff4_store(out + i, niels_from_affine(x, y));
}

// Element inputs (projective) -> the same 96 B affine form, by Montgomery's trick with ONE
// field inversion per CTA: thread t multiplies the Z of its strided elements
// {t, t + T, ...} (prefix products parked in `scratch`, n x 32 B), the CTA combines the
// per-thread products (warp shuffles, then the eight warp totals through shared memory),
// warp 0 alone inverts the CTA product while the other warps wait at the barrier (the
// multiply pipe is free for other CTAs meanwhile), and every thread unwinds its own
// chain backwards.  Per element: 3 M (trick) + 2 M (x, y) + 2 M (2d x y); the inversion
// costs ~380 M per CTA.  (normalize_batch of the reference, ark_curve/element.rs:74-81,
// is the same computation; see k_normalize for the ABI entry point.)
constexpr int kNormBlk = 256;

// kStride: 128 for Elements (X||Y||Z||T), 96 for the T-less D377_PT_XYZ records; T is
// never read -- the affine form recomputes 2d*x*y from x and y.
// kXY: write the AffinePoint wire image x || y (64 B, canonical; Z = 0 gives (0, 0)) instead
// of the bucket operand -- the batch_normalize entry point (ark_curve/element.rs:74-81).
template <int kStride, bool kXY = false>
__global__ void __launch_bounds__(kNormBlk)
k_msm_normalize(const uint8_t* __restrict__ pts, size_t n, size_t T, uint8_t* __restrict__ scratch,
                void* __restrict__ out_v, bool gcd_inv) {
  aff4_t* out = reinterpret_cast<aff4_t*>(out_v);
  uint8_t* out_xy = reinterpret_cast<uint8_t*>(out_v);
  __shared__ fq_t sh[kNormBlk / 32 + 1];
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  fq_t acc = fq_one();
  size_t cnt = 0;
#pragma unroll 1
  for (size_t i = t; i < n; i += T, cnt++) {
    fq_t z = fq_load_wire(pts + (size_t)kStride * i + 64);
    fq_store(scratch + 32 * i, acc);
    // Z = 0 never occurs for a curve point; keep the chain alive anyway
    acc = fq_mul(acc, fq_select(fq_is_zero(z), fq_t(fq_one()), z));
  }
  // inclusive prefix / suffix products of `acc` across the warp
  fq_t pre = acc, suf = acc;
#pragma unroll 1
  for (int o = 1; o < 32; o <<= 1) {
    fq_t y = fq_shfl_up(pre, o);
    fq_t m = fq_mul(pre, y);
    pre = fq_select(lane >= o, m, pre);
    fq_t w = fq_shfl_down(suf, o);
    fq_t m2 = fq_mul(suf, w);
    suf = fq_select(lane + o < 32, m2, suf);
  }
  if (lane == 31) sh[wid] = pre;
  __syncthreads();
  if (wid == 0) {
    fq_t tot = sh[0];
#pragma unroll 1
    for (int v = 1; v < kNormBlk / 32; v++) tot = fq_mul(tot, sh[v]);
    fq_t inv = gcd_inv ? fq_inv_vartime(tot) : fq_inv(tot);
    if (lane == 0) sh[kNormBlk / 32] = inv;
  }
  __syncthreads();
  // 1 / (this warp's total) = inv(CTA total) * product of the other warps' totals
  fq_t inv = sh[kNormBlk / 32];
#pragma unroll 1
  for (int v = 0; v < kNormBlk / 32; v++)
    if (v != wid) inv = fq_mul(inv, sh[v]);
  // 1 / acc = that * (product of the lanes before) * (product of the lanes after)
  fq_t ex_pre = fq_shfl_up(pre, 1), ex_suf = fq_shfl_down(suf, 1);
  inv = fq_mul(inv, fq_select(lane == 0, fq_t(fq_one()), ex_pre));
  inv = fq_mul(inv, fq_select(lane == 31, fq_t(fq_one()), ex_suf));
#pragma unroll 1
  for (size_t k = cnt; k-- > 0;) {
    const size_t i = t + k * T;
    fq_t z = fq_load_wire(pts + (size_t)kStride * i + 64);
    const bool zz = fq_is_zero(z);
    fq_t zi = fq_mul(inv, fq_load_rw(scratch + 32 * i));
    inv = fq_mul(inv, fq_select(zz, fq_t(fq_one()), z));
    fq_t x = fq_mul(fq_load_wire(pts + (size_t)kStride * i), zi);
    fq_t y = fq_mul(fq_load_wire(pts + (size_t)kStride * i + 32), zi);
    if (kXY) {
      fq_store_canon(out_xy + 64 * i, fq_select(zz, fq_t(fq_zero()), x));
      fq_store_canon(out_xy + 64 * i + 32, fq_select(zz, fq_t(fq_zero()), y));
    } else {
      niels_t nl = niels_from_affine(x, y);
      aff4_store(out + i, zz ? niels_identity() : nl);
    }
  }
}

// d377_batch_normalize: same kernel, one resident wave of CTAs (or fewer for small batches).
void launch_normalize(const uint8_t* el, size_t n, uint8_t* scratch, uint8_t* out, cudaStream_t st) {
  Engine& e = engine();
  size_t per = n >> 17;
  per = per < 1 ? 1 : per;
  size_t T = ((n + per - 1) / per + kNormBlk - 1) / kNormBlk * kNormBlk;
  T = std::min(T, (size_t)e.sm_count * 3 * kNormBlk);
  k_msm_normalize<128, true><<<(unsigned)(T / kNormBlk), kNormBlk, 0, st>>>(el, n, T, scratch, out,
                                                                          e.tune_gcd_inv != 0);
}

// ---- 2./4. signed-digit recoding ------------------------------------------
struct MsmGeom {
  int c;        // window width
  int W;        // number of windows = ceil(252 / c)
  uint32_t K;   // buckets per window = 2^(c-1)
};

// bits [w*c, w*c + c) of the 256-bit little-endian scalar
D377_DI uint32_t scalar_window(const fq_raw_t& s, int w, int c) {
  int bit = w * c;
  int limb = bit >> 5, off = bit & 31;
  uint64_t v = s.l[limb];
  if (limb + 1 < 8) v |= (uint64_t)s.l[limb + 1] << 32;
  return (uint32_t)(v >> off) & ((1u << c) - 1u);
}

// Pass 1 (one thread per scalar): range check, signed-digit recoding, bucket histogram.
// For every (window, scalar) it records the entry's bucket id (sign in bit 31, 0xffffffff
// for a zero digit) in `dig` -- 4 bytes, the only per-entry state the sort keeps -- and
// bumps the bucket's count with a fire-and-forget reduction (RED: no return value, so no
// round trip to L2 stalls the warp).
// The kernel handles the windows [wa, wb) of one window group (`counts` and `dig` are the
// group's own arrays, bucket ids are local to the group); the signed-digit carry into
// window wa is recomputed from the windows below it.
__global__ void __launch_bounds__(256, 8)   // <= 32 registers: fits beside 4 accumulation CTAs
k_msm_count(const uint8_t* __restrict__ scalars, size_t n, MsmGeom g, int wa, int wb,
            uint32_t* __restrict__ counts, uint32_t* __restrict__ dig, uint32_t* __restrict__ flags) {
  // grid-stride: when the kernel shares the SMs with an accumulation it is launched with
  // one CTA per SM (the registers an accumulation CTA set leaves over) and walks the batch
#pragma unroll 1
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    fq_raw_t s = fq_load_raw(scalars + 32 * i);
    const bool ok = fr_raw_is_canonical(s);
    if (!ok) atomicOr(flags, 1u);  // contributes nothing; the call reports D377_ERR_SCALAR_RANGE
    uint32_t carry = 0;
#pragma unroll 1
    for (int w = 0; w < wa; w++) carry = (scalar_window(s, w, g.c) + carry) > g.K ? 1u : 0u;
    uint32_t* row = dig + i;
#pragma unroll 1
    for (int w = wa; w < wb; w++, row += n) {
      uint32_t raw = scalar_window(s, w, g.c) + carry;
      carry = raw > g.K ? 1u : 0u;
      int32_t d = (int32_t)raw - (int32_t)(carry << g.c);
      uint32_t e = 0xffffffffu;
      if (d != 0 && ok) {
        uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
        uint32_t id = (uint32_t)(w - wa) * g.K + (mag - 1);
        atomicAdd(&counts[id], 1u);   // result unused: compiles to RED.E.ADD
        e = id | (d < 0 ? 0x80000000u : 0u);
      }
      *row = e;
    }
  }
}

// Pass 2: counting-sort scatter.  `cursor` starts as a copy of the bucket offsets; every
// entry claims its slot with one atomicAdd on its bucket's cursor and stores (point index |
// sign) there.  The walk is window-major: one window at a time keeps the destination region
// (n * 4 B) and its cursors resident in L2, so the random 4-byte stores merge there instead
// of becoming DRAM read-modify-writes.  Every thread moves kScatterIlp entries so that as
// many (digit -> atomic -> store) chains are in flight.
constexpr int kScatterIlp = 4;

__global__ void __launch_bounds__(256, 8)
k_msm_scatter(const uint32_t* __restrict__ dig, size_t n, uint32_t rows, uint32_t* __restrict__ cursor,
              uint32_t* __restrict__ sorted) {
  const size_t tiles_per_row = (n + 256 * kScatterIlp - 1) / (256 * kScatterIlp);
  const size_t total = tiles_per_row * rows;
  // row-major walk: all CTAs work on the same window at (almost) the same time
#pragma unroll 1
  for (size_t tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const size_t r = tile / tiles_per_row;
    const size_t base = (tile - r * tiles_per_row) * (256 * kScatterIlp) + threadIdx.x;
    const uint32_t* row = dig + r * n;
    uint32_t e[kScatterIlp], pos[kScatterIlp];
#pragma unroll
    for (int k = 0; k < kScatterIlp; k++) {
      const size_t i = base + (size_t)k * 256;
      e[k] = i < n ? row[i] : 0xffffffffu;
    }
#pragma unroll
    for (int k = 0; k < kScatterIlp; k++)
      pos[k] = e[k] != 0xffffffffu ? atomicAdd(&cursor[e[k] & 0x7fffffffu], 1u) : 0u;
#pragma unroll
    for (int k = 0; k < kScatterIlp; k++) {
      const size_t i = base + (size_t)k * 256;
      if (e[k] != 0xffffffffu) sorted[pos[k]] = (uint32_t)i | (e[k] & 0x80000000u);
    }
  }
}

// ---- 3. exclusive scan (three small kernels) ---------------------------------
// 256-thread CTAs with few registers: they have to fit beside a resident accumulation
constexpr int kScanBlock = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanBlock * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t warp_sums[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t s = lane < kScanBlock / 32 ? warp_sums[lane] : 0u;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    warp_sums[lane] = s;
  }
  __syncthreads();
  uint32_t before = wid ? warp_sums[wid - 1] : 0u;
  *total = warp_sums[31];
  __syncthreads();
  return before + x - v;
}

__global__ void __launch_bounds__(kScanBlock, 8)
k_scan_tiles(uint32_t* __restrict__ data, size_t n, uint32_t* __restrict__ tile_sums) {
  size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems], sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    v[k] = base + k < n ? data[base + k] : 0u;
    sum += v[k];
  }
  uint32_t total;
  uint32_t ex = block_exclusive_scan(sum, &total);
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    if (base + k < n) data[base + k] = ex;
    ex += v[k];
  }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanBlock)
k_scan_sums(uint32_t* __restrict__ tile_sums, size_t ntiles, uint32_t* __restrict__ grand_total) {
  uint32_t running = 0;
  for (size_t base = 0; base < ntiles; base += kScanBlock) {
    size_t i = base + threadIdx.x;
    uint32_t v = i < ntiles ? tile_sums[i] : 0u;
    uint32_t total;
    uint32_t ex = block_exclusive_scan(v, &total);
    if (i < ntiles) tile_sums[i] = running + ex;
    running += total;
  }
  if (threadIdx.x == 0) *grand_total = running;
}

__global__ void __launch_bounds__(kScanBlock, 8)
k_scan_apply(uint32_t* __restrict__ data, size_t n, const uint32_t* __restrict__ tile_sums,
             uint32_t* __restrict__ copy) {
  size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  uint32_t add = tile_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; k++)
    if (base + k < n) {
      uint32_t v = data[base + k] + add;
      data[base + k] = v;
      copy[base + k] = v;   // the scatter's cursors
    }
}

// ---- 5. bucket accumulation -------------------------------------------------
// `offsets` has nb + 1 entries (offsets[nb] = total number of sorted entries).
// Thread t owns sorted[t*L, (t+1)*L).
// kAffine: `pts` holds canonical affine records (aff4_t) and an addition costs 7
// multiplications; otherwise cached projective records (cached_t, 8 multiplications).
// Both are 128 bytes.
// 112 registers: four CTAs (57 344 registers) leave room for one 256-thread, 32-register
// sort CTA of the second stream on every SM.
template <bool kAffine>
__global__ void __maxnreg__(112)
k_msm_accumulate(const void* __restrict__ pts_v, const uint32_t* __restrict__ sorted,
                 const uint32_t* __restrict__ offsets, uint32_t nb, int L,
                 pt_t* __restrict__ bsum, pt_t* __restrict__ part, int32_t* __restrict__ part_bucket,
                 int32_t bucket_base) {
  const uint8_t* pts = reinterpret_cast<const uint8_t*>(pts_v);
  constexpr uint32_t kRec = 128u;
  // (n W < 2^32 entries, so thread ids and positions fit 32 bits; the few scalars the loop
  // keeps live are chosen so that the kernel fits its 112 registers without a spill)
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t lo, hi;
  {
    const uint32_t total = offsets[nb];
    const uint64_t lo64 = (uint64_t)t * (uint64_t)L;
    if (lo64 >= total) return;
    lo = (uint32_t)lo64;
    hi = (uint32_t)min((uint64_t)total, lo64 + (uint64_t)L);
  }
  // bucket containing position lo: the last b with offsets[b] <= lo
  uint32_t b = 0;
  {
    uint32_t z = nb;  // invariant: offsets[b] <= lo < offsets[z]
    while (z - b > 1) {
      uint32_t m = b + ((z - b) >> 1);
      if (offsets[m] <= lo) b = m; else z = m;
    }
  }
  // does the current bucket begin inside this thread's range?
  bool starts_here = offsets[b] >= lo;
  uint32_t next = offsets[b + 1];
  pt_t acc = pt_identity();
  // Software pipeline on the sorted list and on the operands.  The index of entry pos + 2
  // is loaded while entry pos is added.  Affine records: the first two fields of the operand
  // of entry pos + 1 are loaded into REGISTERS before the addition of entry pos starts (16
  // more live registers), so the gather has a whole addition (~5 us) to arrive -- with a
  // prefetch into L1 and a load at the point of use, 9 % of the stall samples were the first
  // multiply of an addition waiting for its operand (profiles/; the kernel being pipe-bound,
  // removing them is worth < 1 % of its time).  Projective records (four fields) keep the
  // L1 prefetch: a register copy would cost a resident CTA.
  uint32_t e_next = sorted[lo];
  uint32_t e_next2 = lo + 1 < hi ? sorted[lo + 1] : 0u;
  fq_r n_ymx, n_ypx;
  if (kAffine) {
    const uint8_t* rec = pts + (size_t)(e_next & 0x7fffffffu) * kRec;
    const int o = (e_next >> 31) ? 32 : 0;
    n_ymx = fq_assume<1000>(fq_load_stream(rec + o));
    n_ypx = fq_assume<1000>(fq_load_stream(rec + (32 - o)));
  }
#pragma unroll 1
  for (uint32_t pos = lo; pos < hi; pos++) {
    const uint32_t e = e_next;
    e_next = e_next2;
    if (pos + 2 < hi) e_next2 = sorted[pos + 2];
    if (kAffine) {
      // written canonical by aff4_store; the sign of the entry picks the load addresses of
      // (y-x, y+x) and of (2dxy, -2dxy): no selects on limbs
      const fq_r ymx = n_ymx, ypx = n_ypx;
      // third field of the current record: its line came into L1 with the two fields above,
      // and it is not needed before the third multiplication
      const fq_r kt = fq_assume<1000>(
          fq_load_stream(pts + (size_t)(e & 0x7fffffffu) * kRec + 64 + ((e >> 31) ? 32 : 0)));
      if (pos + 1 < hi) {
        const uint8_t* rec = pts + (size_t)(e_next & 0x7fffffffu) * kRec;
        const int o = (e_next >> 31) ? 32 : 0;
        n_ymx = fq_assume<1000>(fq_load_stream(rec + o));
        n_ypx = fq_assume<1000>(fq_load_stream(rec + (32 - o)));
      }
      acc = pt_add_affine<true>(acc, ymx, ypx, kt);
    } else {
      if (pos + 1 < hi)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pts + (size_t)(e_next & 0x7fffffffu) * kRec));
      const bool neg = (e >> 31) != 0;
      const uint8_t* rec = pts + (size_t)(e & 0x7fffffffu) * kRec;
      const int o = neg ? 32 : 0;
      cached_t c;
      c.ymx = fq_load_stream(rec + o);
      c.ypx = fq_load_stream(rec + (32 - o));
      c.kt = fq_load_stream(rec + 64);
      c.z2 = fq_load_stream(rec + 96);
      acc = pt_add_cached<true, true>(acc, c, neg);
    }
    const bool bucket_ends = (pos + 1 == next);
    if (bucket_ends || pos + 1 == hi) {
      // The slot address is recomputed HERE from the special registers (opaque to the
      // optimiser): hoisted out of the loop it is a 64-bit value the compiler keeps alive
      // across every addition, and at 112 registers that meant a spill reloaded per entry.
      uint32_t tt = blockIdx.x * blockDim.x + threadIdx.x;
      asm volatile("" : "+r"(tt));
      if (starts_here && bucket_ends) {
        ptv_store(bsum + b, acc);
      } else if (!starts_here) {
        // piece of a bucket that began in an earlier thread's range
        ptv_store(part + 2 * (size_t)tt, acc);
        part_bucket[2 * (size_t)tt] = bucket_base + (int32_t)b;
      } else {
        // bucket begins here and continues into later ranges: this thread owns it
        ptv_store(part + 2 * (size_t)tt + 1, acc);
        part_bucket[2 * (size_t)tt + 1] = bucket_base + (int32_t)b;
      }
      acc = pt_identity();
      if (bucket_ends && pos + 1 < hi) {
        do {
          b++;
          next = offsets[b + 1];
        } while (next <= pos + 1);
        starts_here = true;   // every later bucket begins at or after pos + 1 > lo
      }
    }
  }
}

// ---- 5b. the same accumulation with a TMA-staged operand stream (experiment) ---------
// The north star asks for "a TMA-staged point stream".  The stream of a bucket sort is a
// gather of 128-byte records at random addresses, so the staging unit is one record per
// lane: every lane issues `cp.async.bulk` (1-D bulk copy, the TMA unit: UBLKCP in SASS) of
// the NEXT entry's record into its own shared-memory slot while the current entry is added;
// a warp shares one mbarrier per stage (one lane arms it with the bytes its lanes are about
// to request, every lane's copy completes on it, every lane waits on its phase).  Two
// stages of 128 x 144 B (the slots are padded by 16 B so that the 16-byte reads of a quarter
// warp fall into distinct banks) = 36 KB per CTA.  It frees the 16 registers of the operand
// prefetch and replaces three LDG.256 per addition by six LDS.128.  Selected with
// D377_MSM_ACC_TMA=1; measured against the register-staged kernel in DESIGN.md 3.4.
constexpr int kTmaSlot = 144;
D377_DI uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

D377_DI fq_r fq_lds(uint32_t addr) {
  fq_r r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]) : "r"(addr));
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7]) : "r"(addr + 16));
  return r;
}

__global__ void __maxnreg__(112)
k_msm_accumulate_tma(const void* __restrict__ pts_v, const uint32_t* __restrict__ sorted,
                     const uint32_t* __restrict__ offsets, uint32_t nb, int L,
                     pt_t* __restrict__ bsum, pt_t* __restrict__ part, int32_t* __restrict__ part_bucket,
                     int32_t bucket_base) {
  extern __shared__ __align__(128) uint8_t tma_smem[];
  const uint8_t* pts = reinterpret_cast<const uint8_t*>(pts_v);
  constexpr uint32_t kRec = 128u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // [stage][thread] slots, then [warp][stage] barriers
  const uint32_t slot0 = smem_u32(tma_smem) + threadIdx.x * kTmaSlot;
  const uint32_t slot_stride = blockDim.x * kTmaSlot;
  const uint32_t bar0 = smem_u32(tma_smem) + 2 * slot_stride + warp * 16;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t total = offsets[nb];
  const uint64_t lo64 = (uint64_t)t * (uint64_t)L;
  // every lane of a warp runs the same L iterations (the barrier protocol is warp-wide);
  // lanes past the end of the list carry no entries
  const bool live = lo64 < total;
  const uint32_t lo = live ? (uint32_t)lo64 : 0u;
  const uint32_t hi = live ? (uint32_t)min((uint64_t)total, lo64 + (uint64_t)L) : 0u;
  uint32_t b = 0;
  if (live) {
    uint32_t z = nb;
    while (z - b > 1) {
      uint32_t m = b + ((z - b) >> 1);
      if (offsets[m] <= lo) b = m; else z = m;
    }
  }
  bool starts_here = live && offsets[b] >= lo;
  uint32_t next = live ? offsets[b + 1] : 0u;
  pt_t acc = pt_identity();
  uint32_t e_next = live ? sorted[lo] : 0u;
  uint32_t e_next2 = lo + 1 < hi ? sorted[lo + 1] : 0u;
  // issue the copy of entry `e` into stage `st` (warp-wide: arm, then every lane with an entry copies)
  auto issue = [&](uint32_t e, bool has, int st) {
    const uint32_t nact = __popc(__ballot_sync(0xffffffffu, has));
    if (nact == 0) return false;
    const uint32_t bar = bar0 + 8 * st;
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nact * kRec) : "memory");
    __syncwarp();
    if (has) {
      const uint8_t* src = pts + (size_t)(e & 0x7fffffffu) * kRec;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(slot0 + st * slot_stride), "l"(src), "r"(kRec), "r"(bar) : "memory");
    }
    return true;
  };
  uint32_t phase0 = 0, phase1 = 0;
  bool armed = issue(e_next, lo < hi, 0);
#pragma unroll 1
  for (int it = 0; it < L; it++) {
    const uint32_t pos = lo + (uint32_t)it;
    const int st = it & 1;
    const bool valid = pos < hi;
    const uint32_t e = e_next;
    e_next = e_next2;
    if (pos + 2 < hi) e_next2 = sorted[pos + 2];
    // next entry into the other stage, then wait for this one
    const bool armed_next = issue(e_next, pos + 1 < hi, st ^ 1);
    if (!armed) break;   // warp-uniform: no lane has an entry left
    {
      const uint32_t bar = bar0 + 8 * st;
      const uint32_t ph = st ? phase1 : phase0;
      uint32_t done = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(ph) : "memory");
      }
      if (st) phase1 ^= 1u; else phase0 ^= 1u;
    }
    armed = armed_next;
    if (valid) {
      const uint32_t slot = slot0 + st * slot_stride;
      const uint32_t o = (e >> 31) ? 32u : 0u;
      const fq_r ymx = fq_lds(slot + o), ypx = fq_lds(slot + (32u - o)), kt = fq_lds(slot + 64u + o);
      acc = pt_add_affine<true>(acc, ymx, ypx, kt);
      const bool bucket_ends = (pos + 1 == next);
      if (bucket_ends || pos + 1 == hi) {
        uint32_t tt = blockIdx.x * blockDim.x + threadIdx.x;
        asm volatile("" : "+r"(tt));
        if (starts_here && bucket_ends) {
          ptv_store(bsum + b, acc);
        } else if (!starts_here) {
          ptv_store(part + 2 * (size_t)tt, acc);
          part_bucket[2 * (size_t)tt] = bucket_base + (int32_t)b;
        } else {
          ptv_store(part + 2 * (size_t)tt + 1, acc);
          part_bucket[2 * (size_t)tt + 1] = bucket_base + (int32_t)b;
        }
        acc = pt_identity();
        if (bucket_ends && pos + 1 < hi) {
          do {
            b++;
            next = offsets[b + 1];
          } while (next <= pos + 1);
          starts_here = true;
        }
      }
    }
  }
}

// ---- 6. stitch buckets that straddle accumulation ranges -------------------
// `keys`/`pts` is a list of slots, each either empty (key < 0) or a partial sum
// of bucket `key`; all pieces of one bucket are consecutive non-empty slots.
// Every thread folds kSegG slots.  A run of equal keys whose neighbours on both
// sides (looked up across the range boundary) belong to other buckets is the
// whole bucket and is stored; a run that may continue in an adjacent range is
// handed to the next level (two output slots per thread), so a bucket that
// holds N entries is finished after ~log_16(N/L) levels whatever the scalar
// distribution is.
constexpr int kSegG = 32;
constexpr int kSegLook = 64;

__global__ void __launch_bounds__(kBlk)
k_msm_seg_reduce(const int32_t* __restrict__ keys, const pt_t* __restrict__ pts, size_t nslots,
                 pt_t* __restrict__ bsum, int32_t* __restrict__ out_keys, pt_t* __restrict__ out_pts,
                 size_t nthreads) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nthreads) return;
  const size_t lo = t * kSegG, hi = min(nslots, lo + (size_t)kSegG);
  int32_t cur = -1;
  int nruns = 0;
  bool out0_used = false;
  pt_t acc = pt_identity();
  // left neighbour: nearest non-empty slot before lo (bounded look-back)
  int32_t left = -2;  // -2: unknown (treat as possibly the same bucket)
  if (lo == 0) {
    left = -1;
  } else {
    for (size_t s = lo, k = 0; s-- > 0 && k < (size_t)kSegLook; k++) {
      int32_t v = keys[s];
      if (v >= 0) { left = v; break; }
      if (s == 0) left = -1;
    }
  }
  out_keys[2 * t] = -1;
  out_keys[2 * t + 1] = -1;
  for (size_t s = lo; s <= hi; s++) {
    int32_t k = s < hi ? keys[s] : -3;  // -3: end-of-range sentinel
    if (k == -1) continue;
    if (k != cur) {
      if (cur >= 0) {
        // close run `cur`
        nruns++;
        bool left_done = nruns > 1 || (left != -2 && left != cur);
        bool right_done;
        if (k != -3) {
          right_done = true;  // another bucket follows inside this range
        } else {
          // look ahead past hi
          int32_t right = -2;
          if (hi >= nslots) {
            right = -1;
          } else {
            for (size_t q = hi, c = 0; q < nslots && c < (size_t)kSegLook; q++, c++) {
              int32_t v = keys[q];
              if (v >= 0) { right = v; break; }
              if (q + 1 == nslots) right = -1;
            }
          }
          right_done = (right != -2 && right != cur);
        }
        if (left_done && right_done) {
          ptv_store(bsum + cur, acc);
        } else {
          size_t o = out0_used ? 2 * t + 1 : 2 * t;
          out0_used = true;
          ptv_store(out_pts + o, acc);
          out_keys[o] = cur;
        }
      }
      if (k == -3) break;
      cur = k;
      acc = ptv_load(pts + s);
    } else {
      acc = pt_add(acc, ptv_load(pts + s));
    }
  }
}

// Warp-parallel form of the same stitching step: one slot per lane, a segmented inclusive
// scan over the warp's 32 slots (empty slots are transparent), so a level costs at most five
// dependent point additions instead of up to 31 -- and usually one, because a scan step
// that no lane needs is skipped by a warp vote (the common layout is [tail piece][head
// piece] pairs: only the distance-1 step has work).  Segments that touch the warp's first
// or last slot and whose neighbour beyond it is, or may be, the same bucket go to the next
// level (two output slots per warp); everything else is a finished bucket.
D377_DI pt_t pt_shfl_up(const pt_t& p, int delta) {
  pt_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.x.l[i] = __shfl_up_sync(0xffffffffu, p.x.l[i], delta);
    r.y.l[i] = __shfl_up_sync(0xffffffffu, p.y.l[i], delta);
    r.z.l[i] = __shfl_up_sync(0xffffffffu, p.z.l[i], delta);
    r.t.l[i] = __shfl_up_sync(0xffffffffu, p.t.l[i], delta);
  }
  return r;
}

// nearest non-empty key in keys[from], keys[from + step], ... (step = +1 / -1), looking at
// most kSegLook slots; -1: ran off the list (no neighbour), -2: not found (unknown)
D377_DI int32_t seg_neighbour(const int32_t* __restrict__ keys, size_t nslots, size_t from, bool forward,
                              bool exists, int lane) {
  if (!exists) return -1;
  int32_t res = -2;
#pragma unroll 1
  for (int round = 0; round < kSegLook / 32 && res == -2; round++) {
    const size_t d = (size_t)round * 32 + lane;
    bool in_range;
    size_t s;
    if (forward) { s = from + d; in_range = s < nslots; }
    else { in_range = d <= from; s = from - d; }
    int32_t k = in_range ? keys[s] : -1;
    uint32_t found = __ballot_sync(0xffffffffu, in_range && k >= 0);
    uint32_t off = __ballot_sync(0xffffffffu, !in_range);
    if (found) {
      res = __shfl_sync(0xffffffffu, k, __ffs(found) - 1);
    } else if (off) {
      res = -1;
    }
  }
  return res;
}

__global__ void __launch_bounds__(kBlk)
k_msm_seg_reduce_warp(const int32_t* __restrict__ keys, const pt_t* __restrict__ pts, size_t nslots,
                      pt_t* __restrict__ bsum, int32_t* __restrict__ out_keys, pt_t* __restrict__ out_pts,
                      size_t nwarps) {
  const size_t gw = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= nwarps) return;   // warp-uniform
  const size_t lo = gw * 32, s = lo + lane;
  const int32_t key = s < nslots ? keys[s] : -1;
  const bool ne = key >= 0;
  const uint32_t nemask = __ballot_sync(0xffffffffu, ne);
  if (lane < 2) out_keys[2 * gw + lane] = -1;
  if (nemask == 0) return;
  __syncwarp();
  pt_t p = ne ? ptv_load(pts + s) : pt_identity();
  // key of the last non-empty slot at or before this lane
  const uint32_t below = nemask & (0xffffffffu >> (31 - lane));
  int32_t fkey = __shfl_sync(0xffffffffu, key, below ? 31 - __clz(below) : 0);
  if (!below) fkey = -1;
  const int32_t pkey = __shfl_up_sync(0xffffffffu, fkey, 1);
  const bool head = ne && (lane == 0 || pkey != key);
  const uint32_t H = __ballot_sync(0xffffffffu, head);
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    // lane i takes the partial sum ending at lane i - d unless a segment starts in (i - d, i]
    const bool need = lane >= d && ((H >> (lane - d + 1)) & ((1u << d) - 1u)) == 0;
    if (!__any_sync(0xffffffffu, need)) continue;
    pt_t q = pt_shfl_up(p, d);
    pt_t sum = pt_add(p, q);
    p = pt_select(need, sum, p);
  }
  const bool is_end = fkey >= 0 && (lane == 31 || ((H >> (lane + 1)) & 1u));
  const uint32_t hs = H & (0xffffffffu >> (31 - lane));   // heads at or before this lane
  const bool first = (hs & (hs - 1)) == 0;                // this lane's segment is the warp's first
  const bool last = lane == 31;
  const int32_t left = seg_neighbour(keys, nslots, lo - 1, false, lo > 0, lane);
  const int32_t right = seg_neighbour(keys, nslots, lo + 32, true, lo + 32 < nslots, lane);
  if (is_end) {
    const bool left_done = !first || (left != -2 && left != fkey);
    const bool right_done = !last || (right != -2 && right != fkey);
    if (left_done && right_done) {
      ptv_store(bsum + fkey, p);
    } else {
      const size_t o = first ? 2 * gw : 2 * gw + 1;
      ptv_store(out_pts + o, p);
      out_keys[o] = fkey;
    }
  }
}

// ---- 7. bucket reduction ------------------------------------------------------
// For window w the sum  sum_j (j + 1) * B_j  over its K buckets, segment by segment and
// window group by window group (the groups finish at different times, see msm_once).
// Same running sums without the per-segment scalar multiplication.  Segment s of a window
// yields A_s = sum_j (j - base + 1) * B_j and P_s = Lseg * sum_j B_j (log2 Lseg doublings);
// the window sum  sum_s (A_s + s * P_s)  is then formed by k_wsum_tree, which carries the
// weight in the node itself: a node that covers 2^k segments holds
//   A = sum_s (A_s + (s - first) * P_s)   and   P = 2^k * sum_s P_s,
// so two neighbours combine as  A = A_l + A_r + P_r,  P = 2 (P_l + P_r)  at every level.
// Segments can therefore be short (many threads, short dependent chains) at no extra cost.
// The bucket offsets (consulted for "is this bucket empty") belong to the window groups.
struct GroupTab {
  int ng;
  int wlo[8], whi[8];
  const uint32_t* offs[8];
};

__global__ void __launch_bounds__(kBlk)
k_msm_bucket_reduce_ap(const pt_t* __restrict__ bsum, GroupTab gt, MsmGeom g, uint32_t Lseg, int log_lseg,
                       uint32_t S, pt_t* __restrict__ out_a, pt_t* __restrict__ out_p) {
  // out_a / out_p: [W][S]
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)g.W * S) return;
  const int w = (int)(idx / S);
  const uint32_t s = (uint32_t)(idx % S);
  int grp = 0;
  while (grp + 1 < gt.ng && !(w >= gt.wlo[grp] && w < gt.whi[grp])) grp++;
  const uint32_t* __restrict__ offsets = gt.offs[grp] + (size_t)(w - gt.wlo[grp]) * g.K;
  const pt_t* __restrict__ bw = bsum + (size_t)w * g.K;
  uint32_t base = s * Lseg;
  uint32_t end = min(base + Lseg, g.K);
  pt_t run = pt_identity(), acc = pt_identity();
  bool any = false;
#pragma unroll 1
  for (uint32_t j = end; j-- > base;) {
    if (offsets[j + 1] > offsets[j]) {
      run = pt_add(run, ptv_load(bw + j));
      any = true;
    }
    if (any) acc = pt_add(acc, run);
  }
  ptv_store(out_a + idx, acc);
  if (any) {
#pragma unroll 1
    for (int k = 0; k < log_lseg; k++) run = pt_dbl<true>(run);
  }
  ptv_store(out_p + idx, run);
}

// ---- 8. tree sums and the final combine ------------------------------------
// in: [groups][len] points; out[groups][ceil(len/G)]: every thread folds G consecutive
// points (used while the list is long enough to fill the machine that way)
// kWire: `in` is a caller's buffer (untrusted limbs, pt_load_wire)
template <bool kWire = false>
__global__ void __launch_bounds__(kBlk)
k_sum_groups(const pt_t* __restrict__ in, uint32_t groups, uint32_t len, uint32_t G,
             pt_t* __restrict__ out) {
  uint32_t olen = (len + G - 1) / G;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)groups * olen) return;
  uint32_t gi = (uint32_t)(idx / olen), o = (uint32_t)(idx % olen);
  uint32_t lo = o * G, hi = min(lo + G, len);
  auto ld = [&](size_t k) {
    const uint8_t* q = reinterpret_cast<const uint8_t*>(in + k);
    return kWire ? pt_load_wire(q) : pt_load(q);
  };
  pt_t acc = ld((size_t)gi * len + lo);
#pragma unroll 1
  for (uint32_t j = lo + 1; j < hi; j++) acc = pt_add(acc, ld((size_t)gi * len + j));
  ptv_store(out + idx, acc);
}

D377_DI fq_t fq_shfl(const fq_t& v, int src) {
  fq_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(0xffffffffu, v.l[i], src);
  return r;
}

D377_DI pt_t pt_shfl_down(const pt_t& p, int delta) {
  pt_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.x.l[i] = __shfl_down_sync(0xffffffffu, p.x.l[i], delta);
    r.y.l[i] = __shfl_down_sync(0xffffffffu, p.y.l[i], delta);
    r.z.l[i] = __shfl_down_sync(0xffffffffu, p.z.l[i], delta);
    r.t.l[i] = __shfl_down_sync(0xffffffffu, p.t.l[i], delta);
  }
  return r;
}

// Latency-oriented sum for short lists: one point per thread, log-depth shuffle tree
// (5 dependent additions per 32 points instead of 31), then the four warp results of a
// CTA through shared memory.  grid = (ceil(len / 128), groups); out[groups][gridDim.x].
__global__ void __launch_bounds__(kBlk)
k_sum_tree(const pt_t* __restrict__ in, uint32_t len, pt_t* __restrict__ out) {
  __shared__ pt_t warp_sum[kBlk / 32];
  const uint32_t gi = blockIdx.y;
  const uint32_t j = blockIdx.x * kBlk + threadIdx.x;
  pt_t p = j < len ? ptv_load(in + (size_t)gi * len + j) : pt_identity();
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) p = pt_add(p, pt_shfl_down(p, d));
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) warp_sum[wid] = p;
  __syncthreads();
  if (wid == 0) {
    p = lane < kBlk / 32 ? warp_sum[lane] : pt_identity();
#pragma unroll 1
    for (int d = kBlk / 64; d >= 1; d >>= 1) p = pt_add(p, pt_shfl_down(p, d));
    if (lane == 0) ptv_store(out + (size_t)gi * gridDim.x + blockIdx.x, p);
  }
}

// Four lanes share one point operation: every lane holds the whole point, computes one
// of the four independent products of each half of the formula and the results are
// exchanged by shuffles, so a doubling costs the latency of 1S + 1M instead of 4S + 4M
// and an addition 3M instead of 9M.  Used by the serial Horner tail only.
D377_DI pt_t pt_gather4(const fq_t& m, int base) {
  pt_t r;
  r.x = fq_shfl(m, base + 0);
  r.y = fq_shfl(m, base + 1);
  r.z = fq_shfl(m, base + 2);
  r.t = fq_shfl(m, base + 3);
  return r;
}

D377_DI pt_t pt_dbl4(const pt_t& p, int role, int base) {
  auto in = fq_select(role == 0, p.x, fq_select(role == 1, p.y, fq_select(role == 2, p.z, fq_add(p.x, p.y))));
  pt_t s = pt_gather4(fq_fold(fq_sqr(in)), base);  // s.x = X^2, s.y = Y^2, s.z = Z^2, s.t = (X+Y)^2
  // same sign arrangement as pt_dbl: F' = C - G, H' = A + B
  auto c = fq_dbl(s.z);
  auto hh = fq_add(s.x, s.y);
  auto e = fq_fold(fq_sub(s.t, hh));
  auto g = fq_fold(fq_sub(s.y, s.x));
  auto ff = fq_sub(c, g);
  // role 0: X3 = E F', 1: Y3 = G H', 2: Z3 = F' G, 3: T3 = E H'
  auto u = fq_select(role == 0 || role == 3, e, fq_select(role == 1, g, ff));
  auto v = fq_select(role == 0, ff, fq_select(role == 2, g, hh));
  return pt_gather4(fq_fold(fq_mul(u, v)), base);
}

D377_DI pt_t pt_add4(const pt_t& p, const pt_t& o, int role, int base) {
  // role 0: A = (Y1-X1)(Y2-X2), 1: B = (Y1+X1)(Y2+X2), 2: D = 2 Z1 Z2, 3: C = K T1 T2
  auto u = fq_select(role == 0, fq_sub(p.y, p.x), fq_select(role == 1, fq_add(p.y, p.x),
           fq_select(role == 2, fq_dbl(p.z), p.t)));
  auto v = fq_select(role == 0, fq_sub(o.y, o.x), fq_select(role == 1, fq_add(o.y, o.x),
           fq_select(role == 2, o.z, fq_fold(fq_mul_small<D377_K>(o.t)))));
  pt_t m = pt_gather4(fq_fold(fq_mul(u, v)), base);  // m.x = A, m.y = B, m.z = D, m.t = C
  auto e = fq_sub(m.y, m.x), f = fq_sub(m.z, m.t);
  auto g = fq_add(m.z, m.t), h = fq_add(m.y, m.x);
  auto uu = fq_select(role == 0 || role == 3, e, fq_select(role == 1, g, f));
  auto vv = fq_select(role == 0, f, fq_select(role == 2, g, h));
  return pt_gather4(fq_fold(fq_mul(uu, vv)), base);
}

// One CTA folds kWT consecutive nodes of one group (missing ones count as empty).  A lone
// warp is bound by its own issue rate (one IMAD.WIDE per 4 clocks), not by the pipe, so a
// node is computed by FOUR lanes (pt_add4 / pt_dbl4: 11 multiplications per lane and node
// instead of 35); levels exchange nodes through shared memory.
// grid = (ceil(len / kWT), groups); out[groups][gridDim.x].
constexpr int kWT = 128;

D377_DI void wnode_combine4(pt_t& a, pt_t& p, const pt_t& al, const pt_t& pl, const pt_t& ar,
                            const pt_t& pr, int role, int base) {
  a = pt_add4(pt_add4(al, ar, role, base), pr, role, base);
  p = pt_dbl4(pt_add4(pl, pr, role, base), role, base);
}

__global__ void __launch_bounds__(2 * kWT)
k_wsum_tree(const pt_t* __restrict__ a_in, const pt_t* __restrict__ p_in, uint32_t len,
            pt_t* __restrict__ a_out, pt_t* __restrict__ p_out) {
  __shared__ pt_t sa[2][kWT / 2], sp[2][kWT / 2];
  const uint32_t gi = blockIdx.y;
  const int lane = threadIdx.x & 31, role = lane & 3, base = lane & ~3;
  const uint32_t grp = threadIdx.x >> 2, wid = threadIdx.x >> 5;
  const size_t row = (size_t)gi * len;
  const uint32_t l = blockIdx.x * kWT + 2 * grp, r = l + 1;
  pt_t a, p;
  {
    pt_t al = l < len ? ptv_load(a_in + row + l) : pt_identity();
    pt_t pl = l < len ? ptv_load(p_in + row + l) : pt_identity();
    pt_t ar = r < len ? ptv_load(a_in + row + r) : pt_identity();
    pt_t pr = r < len ? ptv_load(p_in + row + r) : pt_identity();
    wnode_combine4(a, p, al, pl, ar, pr, role, base);
  }
  int cur = 0;
  if (role == 0) {
    sa[0][grp] = a;
    sp[0][grp] = p;
  }
  __syncthreads();
#pragma unroll 1
  for (uint32_t n = kWT / 4; n >= 1; n >>= 1) {
    // whole warps only: the four-lane operations shuffle with a full mask
    if (wid * 8 < n) {
      const uint32_t j = min(grp, n - 1);
      wnode_combine4(a, p, sa[cur][2 * j], sp[cur][2 * j], sa[cur][2 * j + 1], sp[cur][2 * j + 1], role, base);
      if (role == 0 && grp < n) {
        sa[cur ^ 1][grp] = a;
        sp[cur ^ 1][grp] = p;
      }
    }
    __syncthreads();
    cur ^= 1;
  }
  if (threadIdx.x == 0) {
    const size_t o = (size_t)gi * gridDim.x + blockIdx.x;
    ptv_store(a_out + o, a);
    ptv_store(p_out + o, p);
  }
}

// Horner over window sums: Q = sum_w 2^(c w) * S_w ; W = 0 means identity.  One warp;
// lanes work in groups of four (all groups compute the same thing, lane 0 stores).
__global__ void __launch_bounds__(32)
k_finish(const pt_t* __restrict__ wsums, int W, int c, uint8_t* __restrict__ out_element,
         uint8_t* __restrict__ out_encoding) {
  extern __shared__ uint32_t smem[];
  const int lane = threadIdx.x & 31, role = lane & 3, base = lane & ~3;
  pt_t r = pt_identity();
  if (W > 0) {
    r = ptv_load(wsums + (W - 1));
#pragma unroll 1
    for (int w = W - 2; w >= 0; w--) {
#pragma unroll 1
      for (int k = 0; k < c; k++) r = pt_dbl4(r, role, base);
      r = pt_add4(r, ptv_load(wsums + w), role, base);
    }
  }
  if (lane == 0) D377_DBG_POINT(r);
  if (out_element && lane == 0) pt_store_canon(out_element, r);
  if (out_encoding) {
    isqrt_smem_t sm = isqrt_smem(smem);
    fq_t s = pt_compress_to_field(r, sm);
    if (lane == 0) fq_store(out_encoding, s);
  }
}


// ---- many small MSMs in one call ------------------------------------------------------
// Segment j of `prods` (the products s_i * P_i, 128 B each, written by k_scalar_mul) is
// summed by one warp: lanes stride over the segment, then a shuffle tree.  `okp` (may be
// null) holds the decode status of encoding inputs; a segment with an invalid encoding
// reports ok = 0 (and still gets the sum of its valid pairs, the invalid ones count as the
// identity).  Empty segments give the identity (Element::default()).
__global__ void __launch_bounds__(kBlk)
k_seg_sum(const uint8_t* __restrict__ prods, const uint8_t* __restrict__ okp, const uint32_t* __restrict__ offs,
          size_t nseg, uint8_t* __restrict__ out_el, uint8_t* __restrict__ ok_out) {
  const size_t seg = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (seg >= nseg) return;   // warp-uniform
  const uint32_t lo = offs[seg], hi = offs[seg + 1];
  pt_t acc = pt_identity();
  bool good = true;
#pragma unroll 1
  for (uint32_t i = lo + lane; i < hi; i += 32) {
    acc = pt_add(acc, pt_load(prods + 128 * (size_t)i));
    if (okp) good = good && okp[i] != 0;
  }
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) acc = pt_add(acc, pt_shfl_down(acc, d));
  const bool all_good = __all_sync(0xffffffffu, good);
  if (lane == 0) {
    D377_DBG_POINT(acc);
    pt_store_canon(out_el + 128 * seg, acc);
    if (ok_out) ok_out[seg] = all_good ? 1 : 0;
  }
}

void launch_seg_sum(const uint8_t* prods, const uint8_t* okp, const uint32_t* offs, size_t nseg,
                    uint8_t* out_el, uint8_t* ok_out, cudaStream_t st) {
  k_seg_sum<<<grid_for(nseg * 32, kBlk), kBlk, 0, st>>>(prods, okp, offs, nseg, out_el, ok_out);
  D377_LAUNCHED();
}

D377_DBG_READER(msm_debug_counts)

// ---------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------
static MsmGeom choose_geom(Engine& e, size_t n) {
  int best_c = 4;
  double best = 1e300;
  for (int c = 4; c <= 22; c++) {
    if (e.msm_window_override && c != e.msm_window_override) continue;
    int W = (252 + c - 1) / c;
    double K = std::ldexp(1.0, c - 1);
    // accumulate: 8M per (point, window); reduce: 2 adds (9M) per bucket + the tree;
    // fixed per-window latency of the serial tails.
    double cost = (double)W * (double)n * 8.0 + (double)W * K * 32.0 + (double)W * 3000.0;
    if (cost < best) { best = cost; best_c = c; }
  }
  // From 2^20 pairs the choice is the measured one (tools/tune_msm.py sweeps on B200,
  // gpurun_out/s4n_tune.log): the model above is right up to 2^24 but prefers c = 21 a
  // factor two too early -- with 2^20 buckets per window the accumulation itself slows down
  // (a warp handles a bucket end in almost every iteration) and the bucket reduction takes
  // 3.5 ms.  c = 19, 20 never win (no or one window fewer, twice / four times the buckets).
  // Re-swept in round 2 with the tails hidden under the next MSM (tools/sweep_c_r2.sh,
  // profiles/r2_msm_window_sweep_pipelined.txt): same winners -- 2^20 2.59 / 2.46 / 2.47 ms for
  // c = 15 / 16 / 17, 2^21 4.31 / 4.16 / 4.19 for 16 / 17 / 18, 2^22 7.56 / 7.36 / 8.47 for
  // 17 / 18 / 19, 2^24 26.8 / 30.5 / 28.3 / 27.8 for 18 / 19 / 20 / 21.
  if (!e.msm_window_override && n >= ((size_t)1 << 20)) {
    if (n < ((size_t)3 << 19)) best_c = 16;        // 2^20:        16 (3.53 ms; 17: 3.54, 15: 3.60)
    else if (n < ((size_t)3 << 20)) best_c = 17;   // 2^21:        17 (5.34 ms; 16: 5.46, 18: 5.45)
    else if (n < ((size_t)3 << 24)) best_c = 18;   // 2^22 - 2^25: 18 (2^25: 55.7 ms; 21: 59.8)
    else best_c = 21;                              // 2^26:        21 (113.3 ms; 18: 114.4)
  }
  MsmGeom g;
  g.c = best_c;
  g.W = (252 + best_c - 1) / best_c;
  g.K = 1u << (best_c - 1);
  return g;
}

static size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Streams and events of the pipeline, created on first use (the engine's device is current).
static int msm_state_init(Engine& e) {
  MsmState& ms = e.msm;
  if (ms.ready) return D377_OK;
  int least = 0, greatest = 0;
  D377_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
  D377_CUDA(cudaStreamCreateWithPriority(&ms.sort_stream, cudaStreamNonBlocking, greatest));
  // The tail kernels are short and latency-bound; at the greatest priority their CTAs take
  // the first slots that the long accumulation CTAs of the next MSM vacate instead of
  // queueing behind them.
  D377_CUDA(cudaStreamCreateWithPriority(&ms.tail_stream, cudaStreamNonBlocking,
                                         e.tune_tail_prio ? greatest : least));
  // (Measured back to back, greatest against least priority: 2.73 / 3.04 ms at 2^20,
  // 7.9 / 8.2 ms at 2^22, 27.5 / 29.6 ms at 2^24 with the single-group sort prefetch.)
  D377_CUDA(cudaStreamCreateWithPriority(&ms.points_stream, cudaStreamNonBlocking,
                                         e.tune_points_prio ? greatest : least));
  for (int k = 0; k < 2; k++) D377_CUDA(cudaEventCreateWithFlags(&ms.ev_points_done[k], cudaEventDisableTiming));
  D377_CUDA(cudaEventCreateWithFlags(&ms.ev_fork, cudaEventDisableTiming));
  D377_CUDA(cudaEventCreateWithFlags(&ms.ev_acc_done, cudaEventDisableTiming));
  D377_CUDA(cudaEventCreateWithFlags(&ms.ev_join, cudaEventDisableTiming));
  for (int k = 0; k < 2; k++) D377_CUDA(cudaEventCreateWithFlags(&ms.ev_tail_done[k], cudaEventDisableTiming));
  // per-group events carry timestamps: d377_msm_timeline reads them
  for (int k = 0; k < kMaxGroups; k++) {
    D377_CUDA(cudaEventCreate(&ms.ev_sorted[k]));
    D377_CUDA(cudaEventCreate(&ms.ev_acc0[k]));
    D377_CUDA(cudaEventCreate(&ms.ev_acc[k]));
  }
  D377_CUDA(cudaEventCreate(&ms.ev_sort0));
  D377_CUDA(cudaEventCreate(&ms.ev_sort1));
  for (int k = 0; k <= kMsmStages; k++) D377_CUDA(cudaEventCreate(&ms.ev_stage[k]));
  ms.ready = true;
  return D377_OK;
}

// The stream on which MSM results become complete.  Whoever asks is about to enqueue work
// there, so the engine stream has to be ordered behind it at the next join.
cudaStream_t result_stream(Engine& e) {
  if (!e.tune_tail_overlap || msm_state_init(e) != D377_OK) return e.stream;
  e.msm.tail_pending = true;
  return e.msm.tail_stream;
}

int msm_join(Engine& e) {
  MsmState& ms = e.msm;
  if (!ms.ready || !ms.tail_pending) return D377_OK;
  D377_CUDA(cudaEventRecord(ms.ev_join, ms.tail_stream));
  D377_CUDA(cudaStreamWaitEvent(e.stream, ms.ev_join, 0));
  ms.tail_pending = false;
  return D377_OK;
}

static int finish(cudaStream_t st, const pt_t* wsums, int W, int c, uint8_t* out_element, uint8_t* out_encoding) {
  k_finish<<<1, 32, ISQRT_SMEM_WORDS(32) * sizeof(uint32_t), st>>>(wsums, W, c, out_element, out_encoding);
  D377_LAUNCHED();
  D377_CUDA(cudaGetLastError());
  return D377_OK;
}

// reduce [groups][len] -> [groups][1] in place inside two ping-pong buffers
static int tree_sum(cudaStream_t st, pt_t*& cur, pt_t*& other, uint32_t groups, uint32_t len) {
  while (len > 1) {
    uint32_t olen;
    if ((size_t)groups * len >= ((size_t)1 << 18)) {
      const uint32_t G = 32;
      olen = (len + G - 1) / G;
      size_t total = (size_t)groups * olen;
      k_sum_groups<false><<<grid_for(total, kBlk), kBlk, 0, st>>>(cur, groups, len, G, other);
    } else {
      olen = (len + kBlk - 1) / kBlk;
      k_sum_tree<<<dim3(olen, groups), kBlk, 0, st>>>(cur, len, other);
    }
    D377_LAUNCHED();
    D377_CUDA(cudaGetLastError());
    std::swap(cur, other);
    len = olen;
  }
  return D377_OK;
}

// Sum<Element> (element/projective.rs:131-140) on stream `st` (the engine stream or the
// result stream; each has its own scratch).
int element_sum_on(Engine& e, cudaStream_t st, const uint8_t* elements, size_t n, uint8_t* out_element,
                   uint8_t* out_encoding) {
  if (n == 0) return finish(st, nullptr, 0, 0, out_element, out_encoding);
  if (n > 0xffffffffull) { set_error("element_sum: n too large"); return D377_ERR_INVALID_ARG; }
  DevBuf& ws = e.sum_ws[st == e.stream ? 0 : 1];
  size_t half = align_up(((n + 31) / 32) * sizeof(pt_t));
  int rc = ensure(ws, 2 * half);
  if (rc) return rc;
  pt_t* a = (pt_t*)ws.p;
  pt_t* b = (pt_t*)((uint8_t*)ws.p + half);
  // first level reads the caller's buffer
  const uint32_t G = 32;
  uint32_t len = (uint32_t)n;
  uint32_t olen = (len + G - 1) / G;
  k_sum_groups<true><<<grid_for(olen, kBlk), kBlk, 0, st>>>((const pt_t*)elements, 1, len, G, a);
  D377_LAUNCHED();
  D377_CUDA(cudaGetLastError());
  rc = tree_sum(st, a, b, 1, olen);
  if (rc) return rc;
  return finish(st, a, 1, 0, out_element, out_encoding);
}

// Prepared bases (D377_PT_BASES): any input format -> the affine bucket operands, once.
// Projective inputs are always batch-normalised here, whatever the batch size.
int msm_bases_prepare(const uint8_t* points, int point_format, size_t n, uint8_t* records) {
  Engine& e = engine();
  if (n == 0) return D377_OK;
  aff4_t* aff = (aff4_t*)records;
  uint32_t* dflags = (uint32_t*)(e.d_small + kSmallFlags);
  uint32_t* hflags = (uint32_t*)(e.h_small + kSmallFlags);
  D377_CUDA(cudaMemsetAsync(dflags, 0, 4, e.stream));
  if (point_format == D377_PT_ELEMENT || point_format == D377_PT_XYZ) {
    int rc = ensure(e.scratch, n * 32);
    if (rc) return rc;
    size_t per = n >> 17;
    per = per < 1 ? 1 : per;
    size_t T = ((n + per - 1) / per + kNormBlk - 1) / kNormBlk * kNormBlk;
    T = std::min(T, (size_t)e.sm_count * 3 * kNormBlk);
    if (point_format == D377_PT_ELEMENT)
      k_msm_normalize<128><<<(unsigned)(T / kNormBlk), kNormBlk, 0, e.stream>>>(points, n, T, (uint8_t*)e.scratch.p, aff, e.tune_gcd_inv != 0);
    else
      k_msm_normalize<96><<<(unsigned)(T / kNormBlk), kNormBlk, 0, e.stream>>>(points, n, T, (uint8_t*)e.scratch.p, aff, e.tune_gcd_inv != 0);
  } else if (point_format == D377_PT_AFFINE) {
    k_msm_points_affine<D377_PT_AFFINE><<<grid_for(n, kBlk), kBlk, 0, e.stream>>>(points, n, aff, dflags);
  } else {
    k_msm_points_affine<D377_PT_ENCODING><<<grid_for(n, kBlk), kBlk, ISQRT_SMEM_WORDS(kBlk) * 4, e.stream>>>(points, n, aff, dflags);
  }
  D377_LAUNCHED();
  D377_CUDA(cudaGetLastError());
  D377_CUDA(cudaMemcpyAsync(hflags, dflags, 4, cudaMemcpyDeviceToHost, e.stream));
  D377_CUDA(cudaStreamSynchronize(e.stream));
  return msm_check_flags(*hflags);
}

// One Pippenger.  Head (point conversion, sort, bucket accumulation) on the engine and sort
// streams; tail (stitch, bucket reduction, weighted tree, Horner, compress) on the tail
// stream, over one of two tail-side workspace sets: the head of the next MSM starts right
// behind this MSM's last accumulation and the ~1 ms of latency-bound tail kernels run under
// it.  The result is complete on result_stream(e).
// `scalars_ready`: nullptr = the scalars are ordered on the engine stream like everything
// else (the sort stream forks from it); otherwise the scalars depend on nothing but that
// event (or on nothing at all if *scalars_ready is null), and the scalar side of this MSM
// is NOT ordered behind the engine stream: it starts as soon as its workspace set is free,
// i.e. under the bucket accumulation of the previous MSM.
static int msm_once(const uint8_t* scalars, const uint8_t* points, int point_format, size_t n,
                    uint8_t* out_element, uint8_t* out_encoding, uint32_t* flags /* device, OR-ed */,
                    const cudaEvent_t* scalars_ready = nullptr, const cudaEvent_t* points_ready = nullptr) {
  Engine& e = engine();
  // scalars as the reference holds them in memory (Montgomery limbs): converted on the sort
  // stream into the workspace set before the digits are cut
  const bool sc_mont = (point_format & D377_SCALARS_MONTGOMERY) != 0;
  point_format &= ~D377_SCALARS_MONTGOMERY;
  int rc = msm_state_init(e);
  if (rc) return rc;
  MsmState& ms = e.msm;
  MsmGeom g = choose_geom(e, n);
  const size_t nb = (size_t)g.W * g.K;
  const size_t max_entries = n * (size_t)g.W;
  // Run length per accumulate thread.  Every run boundary splits a bucket into pieces
  // that k_msm_seg_reduce has to stitch (one extra addition each), so runs are as long
  // as the thread count allows: >= ~2 full waves of 128-thread CTAs on 148 SMs.
  int L = max_entries >= ((size_t)1 << 27) ? 128 : max_entries >= ((size_t)1 << 25) ? 64 : 32;
  if (e.tune_acc_run > 0) L = e.tune_acc_run;
  if (max_entries >= 0xfffffff0ull) { set_error("msm chunk too large"); return D377_ERR_INVALID_ARG; }
  // Buckets per bucket-reduce thread.  With the weighted tree (k_wsum_tree) a segment
  // costs no scalar multiplication, so segments can be short: many threads, short
  // dependent chains.
  // With the tails hidden under the next MSM's head, their pipe work matters
  // more than their latency: longer segments (fewer doublings and tree nodes) win once a
  // window has many buckets -- measured back to back: 2^22 7.64 / 7.49 / 7.42 ms and 2^24
  // 27.08 / 26.97 / 27.00 ms with 16 / 32 / 64, 2^20 (c = 16) 2.43 / 2.48 / 2.64 ms.
  uint32_t Lseg = g.c >= 18 ? 64 : g.c == 17 ? 32 : 16;
  if (e.tune_reduce_seg > 0) Lseg = (uint32_t)e.tune_reduce_seg;
  int log_lseg = 0;
  while ((2u << log_lseg) <= Lseg) log_lseg++;
  Lseg = 1u << log_lseg;   // power of two: P_s = Lseg * R_s by doublings
  const uint32_t S = (g.K + Lseg - 1) / Lseg;

  // Window groups: group k+1 is sorted (sort stream) while group k is accumulated (engine
  // stream).  Small MSMs run as one group: the extra launches would cost more than the
  // overlap hides.
  int ngroups = 1;
  if (g.W >= 6) {
    // measured on B200 (tools/tune_msm.py)
    if (max_entries >= ((size_t)3 << 26)) ngroups = std::min(3, g.W / 3);
    else if (max_entries >= ((size_t)1 << 24)) ngroups = 2;
  }
  // When this MSM's sort can run under the accumulation of the MSM before it (its scalars
  // are ready and that accumulation is still in flight), there is nothing left for window
  // groups to hide: one group, a third of the launches (2^24 back to back: 28.14 -> 27.76 ms).
  const bool may_prefetch = scalars_ready != nullptr && e.tune_tail_overlap && e.tune_sort_prefetch;
  if (may_prefetch && ms.tail_used[0] && cudaEventQuery(ms.ev_acc_done) == cudaErrorNotReady) ngroups = 1;
  cudaGetLastError();
  if (e.tune_groups > 0) ngroups = std::min(std::min(e.tune_groups, kMaxGroups), g.W);
  // Group sizes in processing order, weights 2, 3, 4, ...: the first sort has only the
  // point conversion to hide under, every later one the accumulation of the group before
  // it.  wlo[k] .. whi[k] are the windows of the k-th group processed.
  int wlo[kMaxGroups], whi[kMaxGroups];
  {
    int wsum = 0, acc = 0, cut[kMaxGroups + 1];
    // three groups: 2 : 5 : 7 (measured: the first sort then ends with the normalisation
    // and every later one just before the accumulation that waits for it)
    auto weight = [&](int k) { return ngroups == 3 ? (k == 0 ? 2 : k == 1 ? 5 : 7) : k + 2; };
    for (int k = 0; k < ngroups; k++) wsum += weight(k);
    cut[0] = 0;
    for (int k = 0; k < ngroups; k++) {
      acc += weight(k);
      cut[k + 1] = std::max(cut[k] + 1, (int)((long long)acc * g.W / wsum));
    }
    cut[ngroups] = g.W;
    for (int k = ngroups - 1; k > 0; k--) cut[k] = std::min(cut[k], cut[k + 1] - 1);
    // D377_MSM_GW="3,8,3": explicit group sizes in processing order (experiments)
    if (const char* gwenv = getenv("D377_MSM_GW")) {
      int sz[kMaxGroups], ng = 0, tot = 0;
      for (const char* q = gwenv; *q && ng < kMaxGroups;) {
        sz[ng] = atoi(q);
        tot += sz[ng++];
        while (*q && *q != ',') q++;
        if (*q == ',') q++;
      }
      if (tot == g.W) {
        ngroups = ng;
        cut[0] = 0;
        for (int k = 0; k < ng; k++) cut[k + 1] = cut[k] + sz[k];
      }
    }
    for (int k = 0; k < ngroups; k++) {
      whi[k] = g.W - cut[k];
      wlo[k] = g.W - cut[k + 1];
    }
  }
  // sort kernels beside a resident accumulation: one 256-thread CTA per SM
  const unsigned sort_cap = ngroups > 1 || may_prefetch
                                ? (unsigned)e.sm_count * (unsigned)std::max(1, e.tune_sort_ctas) : 0u;
  auto capped = [&](size_t want) { return (unsigned)(sort_cap ? std::min<size_t>(want, sort_cap) : want); };
  size_t g_nthreads[kMaxGroups], g_tbase[kMaxGroups + 1], g_ntiles[kMaxGroups];
  g_tbase[0] = 0;
  for (int k = 0; k < ngroups; k++) {
    size_t ent_k = (size_t)(whi[k] - wlo[k]) * n;
    g_nthreads[k] = (ent_k + L - 1) / L;
    g_tbase[k + 1] = g_tbase[k] + g_nthreads[k];
    g_ntiles[k] = ((size_t)(whi[k] - wlo[k]) * g.K + 1 + kScanTile - 1) / kScanTile;
  }
  const size_t nthreads = g_tbase[ngroups];

  // Bucket additions against affine records cost 7 M instead of 8: free when the input
  // already has Z = 1; Element inputs are normalised first when the batch is large enough
  // for the one-inversion-per-CTA trick to pay (7 M + one inversion per CTA against W
  // multiplications saved).
  const bool projective = point_format == D377_PT_ELEMENT || point_format == D377_PT_XYZ;
  const bool prepared = point_format == D377_PT_BASES;   // `points` already holds the bucket operands
  bool affine = !projective;
  size_t norm_per = 0, norm_T = 0;
  if (projective) {
    norm_per = n >> 17;                       // >= 2^17 threads stay busy
    if (norm_per > 64) norm_per = 64;
    int want = e.tune_normalize;              // D377_MSM_NORMALIZE: -1 off, 1 force, 0 auto
    if (want > 0 && norm_per < 4) norm_per = 4;
    if (want >= 0 && norm_per >= (size_t)e.tune_norm_min_per) affine = true;
    if (want > 0) affine = true;
    if (affine) norm_T = ((n + norm_per - 1) / norm_per + kNormBlk - 1) / kNormBlk * kNormBlk;
    // One wave: every CTA is resident from the start (3 per SM at 70 registers), so the
    // kernel has no ragged last wave and the CTA-wide inversion bubble occurs once.
    if (affine && e.tune_norm_wave > 0)
      norm_T = std::min(norm_T, (size_t)e.sm_count * (size_t)e.tune_norm_wave * kNormBlk);
  }

  // ---- workspaces -----------------------------------------------------------------
  // Two complete sets (bucket operands, digits, sorted lists, offsets, bucket sums, piece
  // slots): the conversion of the points and the sort of the NEXT MSM run under this MSM's
  // bucket accumulation, and the tail of this MSM under the next MSM's accumulation.
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
  // The other set is only needed while the previous MSM is still in flight: a caller that
  // waits for every result (d377_msm_dev, d377_msm) keeps using set 0 and pays for one.
  int set = 0;
  if (e.tune_tail_overlap && ms.tail_used[ms.cur_set] &&
      cudaEventQuery(ms.ev_tail_done[ms.cur_set]) == cudaErrorNotReady)
    set = ms.cur_set ^ 1;
  cudaGetLastError();
  ms.cur_set = set;
  size_t o_cached = carve(prepared ? 0 : n * sizeof(cached_t));
  size_t o_norm = carve(affine && norm_T ? n * 32 : 0);
  size_t o_dig = carve(max_entries * 4);
  size_t o_sorted = carve(max_entries * 4);
  size_t o_sc = carve(sc_mont ? n * 32 : 0);
  size_t o_counts[kMaxGroups], o_cursor[kMaxGroups], o_tiles[kMaxGroups];
  for (int k = 0; k < ngroups; k++) {
    o_counts[k] = carve(((size_t)(whi[k] - wlo[k]) * g.K + 1) * 4);
    o_cursor[k] = carve(((size_t)(whi[k] - wlo[k]) * g.K + 1) * 4);
    o_tiles[k] = carve(g_ntiles[k] * 4 + 4);
  }
  size_t o_bsum = carve(nb * sizeof(pt_t));
  size_t o_part = carve(2 * nthreads * sizeof(pt_t));
  size_t o_pb = carve(2 * nthreads * 4);
  // second and third stitch level (the levels ping-pong between them)
  const size_t nthreads2 = (2 * nthreads + kSegG - 1) / kSegG;
  const size_t nthreads3 = (2 * nthreads2 + kSegG - 1) / kSegG;
  size_t o_part2 = carve(2 * nthreads2 * sizeof(pt_t));
  size_t o_pb2 = carve(2 * nthreads2 * 4);
  size_t o_part3 = carve(2 * nthreads3 * sizeof(pt_t));
  size_t o_pb3 = carve(2 * nthreads3 * 4);
  const size_t seg_n = (size_t)g.W * S, seg_n2 = (size_t)g.W * ((S + kWT - 1) / kWT);
  size_t o_seg_a = carve(seg_n * sizeof(pt_t));
  size_t o_seg_p = carve(seg_n * sizeof(pt_t));
  size_t o_seg_a2 = carve(seg_n2 * sizeof(pt_t));
  size_t o_seg_p2 = carve(seg_n2 * sizeof(pt_t));
  rc = ensure(ms.tail_ws[set], off);
  if (rc) return rc;

  uint8_t* tw = (uint8_t*)ms.tail_ws[set].p;
  uint8_t* ws = tw;
  cached_t* cached = (cached_t*)(ws + o_cached);
  const aff4_t* aff_in = prepared ? (const aff4_t*)points : (const aff4_t*)(ws + o_cached);
  aff4_t* aff = (aff4_t*)(ws + o_cached);
  uint32_t* dig = (uint32_t*)(tw + o_dig);
  uint32_t* sorted = (uint32_t*)(tw + o_sorted);
  pt_t* bsum = (pt_t*)(tw + o_bsum);
  pt_t* part = (pt_t*)(tw + o_part);
  int32_t* pb = (int32_t*)(tw + o_pb);
  cudaStream_t st = e.stream, ss = ms.sort_stream;
  // point conversion / normalisation: on the engine stream, or -- when the inputs are known to
  // be ready -- on a stream of its own, so that it runs under the previous MSM's accumulation
  // (its CTA-wide inversion bubble and its two passes over the points are then hidden; the
  // engine stream holds nothing but bucket accumulations)
  // Measured back to back (ms per MSM at 2^20 / 2^21 / 2^22 / 2^24): on the engine stream
  // 2.44 / 4.16 / 7.44 / 27.03, on its own stream at the least priority 2.57 / 4.09 / 7.36 /
  // 26.81, at the greatest 2.75 / 4.19 / 7.43 / 27.01 -- the conversion is multiply-pipe work
  // like the accumulation it runs under, so only its bubbles are hidden: ~1 % from 2^21 pairs
  // on one GPU.  In the 8-GPU strong-scaling loop (2^21 pairs per GPU, an all-gather per
  // step) it costs 4 % instead (4.28 -> 4.45 ms), so it is used from 2^22 pairs.
  const bool pts_prefetch = may_prefetch && e.tune_points_prefetch && !prepared &&
                            n >= ((size_t)1 << 22);
  cudaStream_t ps = pts_prefetch ? ms.points_stream : st;
  cudaStream_t ts = e.tune_tail_overlap ? ms.tail_stream : e.stream;

  ms.last_geom.c = g.c;
  ms.last_geom.W = g.W;
  ms.last_n = n;
  ms.last_affine = affine;
  ms.last_groups = ngroups;
  ms.stage_valid = true;
  auto stage_mark = [&](int i, cudaStream_t s) { return cudaEventRecord(ms.ev_stage[i], s); };
  D377_CUDA(stage_mark(0, st));
  // The scalar side needs its inputs and its workspace set (the tail that last used the set
  // has to be done).  By default the inputs are ordered on the engine stream, so the sort
  // stream forks from it -- behind the previous MSM's accumulation.  With `scalars_ready`
  // the only dependency is that event.
  const bool prefetch = may_prefetch;
  if (prefetch) {
    if (*scalars_ready) D377_CUDA(cudaStreamWaitEvent(ss, *scalars_ready, 0));
  } else {
    D377_CUDA(cudaEventRecord(ms.ev_fork, st));
    D377_CUDA(cudaStreamWaitEvent(ss, ms.ev_fork, 0));
  }
  // (the points of a host-buffer MSM arrive after its scalars: the sort above only waits for
  // the scalars, the point conversion for the whole chunk)
  if (pts_prefetch) {
    const cudaEvent_t* pr = points_ready ? points_ready : scalars_ready;
    if (*pr) D377_CUDA(cudaStreamWaitEvent(ps, *pr, 0));
  }
  if (ms.tail_used[set]) {
    D377_CUDA(cudaStreamWaitEvent(ss, ms.ev_tail_done[set], 0));
    D377_CUDA(cudaStreamWaitEvent(st, ms.ev_tail_done[set], 0));
    if (pts_prefetch) D377_CUDA(cudaStreamWaitEvent(ps, ms.ev_tail_done[set], 0));
  }
  D377_CUDA(cudaEventRecord(ms.ev_sort0, ss));
  if (sc_mont) {
    launch_fr_from_mont(scalars, n, tw + o_sc, ss);
    scalars = tw + o_sc;
  }

  // ---- scalar side, all groups, on the sort stream (2, 3, 4) ----
  for (int k = 0; k < ngroups; k++) {
    const int wa = wlo[k], wb = whi[k];
    const size_t nbk = (size_t)(wb - wa) * g.K;
    uint32_t* counts = (uint32_t*)(tw + o_counts[k]);
    uint32_t* tiles = (uint32_t*)(tw + o_tiles[k]);
    uint32_t* cursor = (uint32_t*)(tw + o_cursor[k]);
    D377_CUDA(cudaMemsetAsync(counts, 0, (nbk + 1) * 4, ss));
    k_msm_count<<<capped(grid_for(n, 256)), 256, 0, ss>>>(scalars, n, g, wa, wb, counts, dig + (size_t)wa * n, flags);
    D377_LAUNCHED();
    k_scan_tiles<<<(unsigned)g_ntiles[k], kScanBlock, 0, ss>>>(counts, nbk + 1, tiles);
    D377_LAUNCHED();
    k_scan_sums<<<1, kScanBlock, 0, ss>>>(tiles, g_ntiles[k], tiles + g_ntiles[k]);
    D377_LAUNCHED();
    k_scan_apply<<<(unsigned)g_ntiles[k], kScanBlock, 0, ss>>>(counts, nbk + 1, tiles, cursor);
    D377_LAUNCHED();
    k_msm_scatter<<<capped((size_t)grid_for(n, 256 * kScatterIlp) * (size_t)(wb - wa)), 256, 0, ss>>>(
        dig + (size_t)wa * n, n, (uint32_t)(wb - wa), cursor, sorted + (size_t)wa * n);
    D377_LAUNCHED();
    D377_CUDA(cudaEventRecord(ms.ev_sorted[k], ss));
  }
  D377_CUDA(cudaEventRecord(ms.ev_sort1, ss));

  // ---- point side on the engine stream ----
  D377_CUDA(cudaMemsetAsync(pb, 0xff, 2 * nthreads * 4, st));
  // 1
  if (!prepared) {
    dim3 gr(grid_for(n, kBlk));
    if (point_format == D377_PT_ELEMENT && affine)
      k_msm_normalize<128><<<(unsigned)(norm_T / kNormBlk), kNormBlk, 0, ps>>>(points, n, norm_T, ws + o_norm, aff, e.tune_gcd_inv != 0);
    else if (point_format == D377_PT_XYZ && affine)
      k_msm_normalize<96><<<(unsigned)(norm_T / kNormBlk), kNormBlk, 0, ps>>>(points, n, norm_T, ws + o_norm, aff, e.tune_gcd_inv != 0);
    else if (point_format == D377_PT_ELEMENT)
      k_msm_points<D377_PT_ELEMENT><<<gr, kBlk, 0, ps>>>(points, n, cached, flags);
    else if (point_format == D377_PT_XYZ)
      k_msm_points<D377_PT_XYZ><<<gr, kBlk, 0, ps>>>(points, n, cached, flags);
    else if (point_format == D377_PT_AFFINE)
      k_msm_points_affine<D377_PT_AFFINE><<<gr, kBlk, 0, ps>>>(points, n, aff, flags);
    else
      k_msm_points_affine<D377_PT_ENCODING><<<gr, kBlk, ISQRT_SMEM_WORDS(kBlk) * 4, ps>>>(points, n, aff, flags);
    D377_LAUNCHED();
    if (pts_prefetch) {
      D377_CUDA(cudaEventRecord(ms.ev_points_done[set], ps));
      D377_CUDA(cudaStreamWaitEvent(st, ms.ev_points_done[set], 0));
    }
  }
  for (int i = 1; i <= 4; i++) D377_CUDA(stage_mark(i, st));
  // 5: bucket accumulation, group by group as the sorted lists arrive.
  // 128 threads, ~106 registers, 4 CTAs/SM: measured flat from 4 to 7 CTAs/SM (the
  // fmaheavy pipe, not latency, is the limiter), slower at 8 (spills).
  for (int k = 0; k < ngroups; k++) {
    const int wa = wlo[k], wb = whi[k];
    const size_t nbk = (size_t)(wb - wa) * g.K;
    const uint32_t* counts = (const uint32_t*)(tw + o_counts[k]);
    D377_CUDA(cudaStreamWaitEvent(st, ms.ev_sorted[k], 0));
    D377_CUDA(cudaEventRecord(ms.ev_acc0[k], st));
    if (affine && e.tune_acc_tma)
      k_msm_accumulate_tma<<<grid_for(g_nthreads[k], kBlk), kBlk, 2 * kBlk * kTmaSlot + (kBlk / 32) * 16, st>>>(
          aff_in, sorted + (size_t)wa * n, counts, (uint32_t)nbk, L, bsum + (size_t)wa * g.K,
          part + 2 * g_tbase[k], pb + 2 * g_tbase[k], (int32_t)((size_t)wa * g.K));
    else if (affine)
      k_msm_accumulate<true><<<grid_for(g_nthreads[k], kBlk), kBlk, 0, st>>>(
          aff_in, sorted + (size_t)wa * n, counts, (uint32_t)nbk, L, bsum + (size_t)wa * g.K,
          part + 2 * g_tbase[k], pb + 2 * g_tbase[k], (int32_t)((size_t)wa * g.K));
    else
      k_msm_accumulate<false><<<grid_for(g_nthreads[k], kBlk), kBlk, 0, st>>>(
          cached, sorted + (size_t)wa * n, counts, (uint32_t)nbk, L, bsum + (size_t)wa * g.K,
          part + 2 * g_tbase[k], pb + 2 * g_tbase[k], (int32_t)((size_t)wa * g.K));
    D377_LAUNCHED();
    D377_CUDA(cudaEventRecord(ms.ev_acc[k], st));
  }
  D377_CUDA(stage_mark(5, st));
  // ---- tail: everything below runs on the tail stream ----
  D377_CUDA(cudaEventRecord(ms.ev_acc_done, st));
  if (ts != st) D377_CUDA(cudaStreamWaitEvent(ts, ms.ev_acc_done, 0));
  // 6: stitch the bucket pieces of all groups; the slot lists of the levels rotate through
  // two buffers (the first level reads the accumulation's own part / pb arrays)
  {
    const int32_t* kin = pb;
    const pt_t* pin = part;
    size_t nslots = 2 * nthreads;
    int32_t* kbuf[2] = {(int32_t*)(tw + o_pb2), (int32_t*)(tw + o_pb3)};
    pt_t* pbuf[2] = {(pt_t*)(tw + o_part2), (pt_t*)(tw + o_part3)};
    for (int lvl = 0;; lvl++) {
      size_t nt = (nslots + kSegG - 1) / kSegG;
      int32_t* kout = kbuf[lvl & 1];
      pt_t* pout = pbuf[lvl & 1];
      // work-efficient serial fold while the level is large, warp scan (log depth) below
      if (e.tune_stitch_warp && nslots <= (size_t)e.tune_stitch_warp)
        k_msm_seg_reduce_warp<<<grid_for(nt * 32, kBlk), kBlk, 0, ts>>>(kin, pin, nslots, bsum, kout, pout, nt);
      else
        k_msm_seg_reduce<<<grid_for(nt, kBlk), kBlk, 0, ts>>>(kin, pin, nslots, bsum, kout, pout, nt);
      D377_LAUNCHED();
      if (nt == 1) break;
      nslots = 2 * nt;
      kin = kout;
      pin = pout;
    }
  }
  D377_CUDA(stage_mark(6, ts));
  // 7: bucket reduction of every window
  pt_t *a = (pt_t*)(tw + o_seg_a), *p = (pt_t*)(tw + o_seg_p);
  pt_t *a2 = (pt_t*)(tw + o_seg_a2), *p2 = (pt_t*)(tw + o_seg_p2);
  {
    GroupTab gt;
    gt.ng = ngroups;
    for (int k = 0; k < ngroups; k++) {
      gt.wlo[k] = wlo[k];
      gt.whi[k] = whi[k];
      gt.offs[k] = (const uint32_t*)(tw + o_counts[k]);
    }
    k_msm_bucket_reduce_ap<<<grid_for((size_t)g.W * S, kBlk), kBlk, 0, ts>>>(bsum, gt, g, Lseg, log_lseg, S, a, p);
    D377_LAUNCHED();
  }
  D377_CUDA(cudaGetLastError());
  D377_CUDA(stage_mark(7, ts));
  // 8: weighted tree -> one sum per window, Horner over the windows, output
  for (uint32_t len = S; len > 1;) {
    const uint32_t olen = (len + kWT - 1) / kWT;
    k_wsum_tree<<<dim3(olen, (unsigned)g.W), 2 * kWT, 0, ts>>>(a, p, len, a2, p2);
    D377_LAUNCHED();
    std::swap(a, a2);
    std::swap(p, p2);
    len = olen;
  }
  rc = finish(ts, a, g.W, g.c, out_element, out_encoding);
  if (rc) return rc;
  D377_CUDA(stage_mark(8, ts));
  D377_CUDA(cudaEventRecord(ms.ev_tail_done[set], ts));
  ms.tail_used[set] = true;
  if (ts != st) ms.tail_pending = true;
  return D377_OK;
}

// Enqueue a whole MSM (no host synchronisation).  `flags` (device word, reset by the
// caller) collects bit 0 = non-canonical scalar, bit 1 = invalid encoding.  The result is
// complete on result_stream(e).
int msm_enqueue(const uint8_t* scalars, const uint8_t* points, int point_format, size_t n,
                uint8_t* out_element, uint8_t* out_encoding, uint32_t* flags, size_t chunk,
                const cudaEvent_t* chunk_ready, bool inputs_ready, const cudaEvent_t* scalars_ready_ev) {
  Engine& e = engine();
  const cudaEvent_t no_event = nullptr;
  if (n == 0) return finish(result_stream(e), nullptr, 0, 0, out_element, out_encoding);
  const int pf = point_format & ~D377_SCALARS_MONTGOMERY;
  const size_t pbytes = pf == D377_PT_ELEMENT || pf == D377_PT_BASES ? 128
                        : pf == D377_PT_ENCODING ? 32 : pf == D377_PT_XYZ ? 96 : 64;
  const size_t kMax = (size_t)1 << 26;  // keeps n * W below 2^32
  if (chunk == 0 || chunk > kMax) chunk = kMax;
  size_t nchunks = (n + chunk - 1) / chunk;
  if (nchunks > 16) { set_error("msm: too many chunks (n = %zu, chunk = %zu)", n, chunk); return D377_ERR_INVALID_ARG; }
  // inputs_ready: the inputs depend on nothing but chunk_ready[k] (or on nothing)
  auto ready_of = [&](size_t k) -> const cudaEvent_t* {
    return !inputs_ready ? nullptr : chunk_ready ? &chunk_ready[k] : &no_event;
  };
  // scalars_ready_ev[k] (optional): the scalars of chunk k alone are uploaded -- the scalar
  // side starts on it, ahead of the chunk's points
  auto sc_ready_of = [&](size_t k) -> const cudaEvent_t* {
    return inputs_ready && scalars_ready_ev ? &scalars_ready_ev[k] : ready_of(k);
  };
  if (nchunks == 1) {
    if (chunk_ready) D377_CUDA(cudaStreamWaitEvent(e.stream, chunk_ready[0], 0));
    return msm_once(scalars, points, point_format, n, out_element, out_encoding, flags, sc_ready_of(0), ready_of(0));
  }
  pt_t* partials = (pt_t*)(e.d_small + kSmallPartials);  // up to 16 chunk results
  for (size_t k = 0; k < nchunks; k++) {
    size_t lo = k * chunk, len = std::min(chunk, n - lo);
    if (chunk_ready) D377_CUDA(cudaStreamWaitEvent(e.stream, chunk_ready[k], 0));
    int rc = msm_once(scalars + 32 * lo, points + pbytes * lo, point_format, len,
                      (uint8_t*)(partials + k), nullptr, flags, sc_ready_of(k), ready_of(k));
    if (rc) return rc;
  }
  cudaStream_t rs = result_stream(e);
  pt_t* tmp = (pt_t*)(e.d_small + kSmallTmp);
  k_sum_groups<false><<<1, kBlk, 0, rs>>>(partials, 1, (uint32_t)nchunks, 32, tmp);
  D377_LAUNCHED();
  D377_CUDA(cudaGetLastError());
  return finish(rs, tmp, 1, 0, out_element, out_encoding);
}

int msm_check_flags(uint32_t flags) {
  if (flags & 1u) {
    set_error("msm: a scalar is not canonical (>= r)");
    return D377_ERR_SCALAR_RANGE;
  }
  if (flags & 2u) {
    set_error("msm: an input encoding is invalid");
    return D377_ERR_INVALID_ENCODING;
  }
  return D377_OK;
}

static int msm_args_ok(Engine& e, const uint8_t* scalars, const uint8_t* points, int point_format, size_t n) {
  if (point_format >= 0) point_format &= ~D377_SCALARS_MONTGOMERY;
  if (point_format < 0 || point_format > 4) { set_error("bad point_format %d", point_format); return D377_ERR_INVALID_ARG; }
  if (n && (!scalars || !points)) { set_error("null pointer"); return D377_ERR_INVALID_ARG; }
  if (point_format == D377_PT_BASES && n) return check_bases(e, points, n);
  return D377_OK;
}

int msm_dev(const uint8_t* scalars, const uint8_t* points, int point_format, size_t n,
            uint8_t* out_element, uint8_t* out_encoding) {
  Engine& e = engine();
  int rc = msm_args_ok(e, scalars, points, point_format, n);
  if (rc) return rc;
  uint32_t* dflags = (uint32_t*)(e.d_small + kSmallFlags);
  uint32_t* hflags = (uint32_t*)(e.h_small + kSmallFlags);
  D377_CUDA(cudaMemsetAsync(dflags, 0, 4, e.stream));
  rc = msm_enqueue(scalars, points, point_format, n, out_element, out_encoding, dflags);
  if (rc) return rc;
  // status word: the only host read-back of the pipeline
  cudaStream_t rs = result_stream(e);
  D377_CUDA(cudaMemcpyAsync(hflags, dflags, 4, cudaMemcpyDeviceToHost, rs));
  D377_CUDA(cudaStreamSynchronize(rs));
  return msm_check_flags(*hflags);
}

// Same, without the host synchronisation: the status word is sticky and d377_sync reports
// it.  Back-to-back calls overlap the tail of one MSM with the head of the next.
int msm_dev_async(const uint8_t* scalars, const uint8_t* points, int point_format, size_t n,
                  uint8_t* out_element, uint8_t* out_encoding, int flags) {
  Engine& e = engine();
  int rc = msm_args_ok(e, scalars, points, point_format, n);
  if (rc) return rc;
  e.async_status_dirty = true;
  return msm_enqueue(scalars, points, point_format, n, out_element, out_encoding,
                     (uint32_t*)(e.d_small + kSmallAsyncFlags), 0, nullptr,
                     (flags & D377_MSM_INPUTS_READY) != 0);
}

bool msm_last_mixed() { return engine().msm.last_affine; }

// d377_shutdown: the streams and events above belong to the device being released
void msm_shutdown(Engine& e) {
  MsmState& ms = e.msm;
  if (ms.ready) {
    cudaStreamSynchronize(ms.sort_stream);
    cudaStreamSynchronize(ms.tail_stream);
    cudaStreamSynchronize(ms.points_stream);
    cudaStreamDestroy(ms.points_stream);
    for (int k = 0; k < 2; k++) cudaEventDestroy(ms.ev_points_done[k]);
    cudaStreamDestroy(ms.sort_stream);
    cudaStreamDestroy(ms.tail_stream);
    ms.sort_stream = ms.tail_stream = nullptr;
    cudaEventDestroy(ms.ev_fork);
    cudaEventDestroy(ms.ev_acc_done);
    cudaEventDestroy(ms.ev_join);
    for (int k = 0; k < 2; k++) cudaEventDestroy(ms.ev_tail_done[k]);
    for (int k = 0; k < kMaxGroups; k++) {
      cudaEventDestroy(ms.ev_sorted[k]);
      cudaEventDestroy(ms.ev_acc0[k]);
      cudaEventDestroy(ms.ev_acc[k]);
    }
    cudaEventDestroy(ms.ev_sort0);
    cudaEventDestroy(ms.ev_sort1);
    for (int k = 0; k <= kMsmStages; k++) cudaEventDestroy(ms.ev_stage[k]);
  }
  for (int k = 0; k < 2; k++) {
    if (ms.tail_ws[k].p) cudaFree(ms.tail_ws[k].p);
    ms.tail_ws[k] = DevBuf();
  }
  ms = MsmState();
}

int msm_stage_info(float* ms_out, int* c, int* W, uint64_t* n) {
  MsmState& ms = engine().msm;
  if (!ms.ready || !ms.stage_valid) { set_error("no msm has run yet"); return D377_ERR_INVALID_ARG; }
  D377_CUDA(cudaEventSynchronize(ms.ev_stage[kMsmStages]));
  for (int k = 0; k < kMsmStages; k++) D377_CUDA(cudaEventElapsedTime(&ms_out[k], ms.ev_stage[k], ms.ev_stage[k + 1]));
  // the scalar side runs on its own stream, overlapped with `points` and `accumulate`:
  // its whole span (count + scan + scatter of every window group) is reported as `count`
  D377_CUDA(cudaEventElapsedTime(&ms_out[1], ms.ev_sort0, ms.ev_sort1));
  if (c) *c = ms.last_geom.c;
  if (W) *W = ms.last_geom.W;
  if (n) *n = ms.last_n;
  return D377_OK;
}

// Per window group of the most recent single-chunk MSM, in ms after its start: sorted list
// ready (sort stream), accumulation start / end (engine stream), end of the MSM.
int msm_timeline(float* out, int cap, int* ngroups) {
  MsmState& ms = engine().msm;
  if (!ms.ready || !ms.stage_valid) { set_error("no msm has run yet"); return D377_ERR_INVALID_ARG; }
  D377_CUDA(cudaEventSynchronize(ms.ev_stage[kMsmStages]));
  const int ng = ms.last_groups;
  if (ngroups) *ngroups = ng;
  for (int k = 0; k < ng && 4 * k + 3 < cap; k++) {
    D377_CUDA(cudaEventElapsedTime(&out[4 * k + 0], ms.ev_stage[0], ms.ev_sorted[k]));
    D377_CUDA(cudaEventElapsedTime(&out[4 * k + 1], ms.ev_stage[0], ms.ev_acc0[k]));
    D377_CUDA(cudaEventElapsedTime(&out[4 * k + 2], ms.ev_stage[0], ms.ev_acc[k]));
    D377_CUDA(cudaEventElapsedTime(&out[4 * k + 3], ms.ev_stage[0], ms.ev_stage[kMsmStages]));
  }
  return D377_OK;
}

}  // namespace d377
