// Inverse square root in Fq: fq_isqrt(x) == Fq::sqrt_ratio_zeta(&ONE, &x)
// (reference src/ark_curve/invsqrt.rs:75-166, spec sqrt_alg.sage:35-109), the
// only form the hot path calls (encoding.rs:57,102; elligator.rs:26).
//
// The reference evaluates w = den^(2^47-1) ... to support a general ratio;
// with num = 1 the same root falls out of one 205-bit power:
//   a  = x^((m-1)/2)                      (sliding window, 201 S + 36 M + 8)
//   z  = x * a^2 = x^m                    (element of the 2^47 subgroup <g>)
//   t' = the 47-bit value with z * g^t' = 1, found 8 bits at a time with the
//        reference's tables (Sarkar 2020): 39 S + 15 M + 6 lookups
//   x square  (t' even): 1/sqrt(x)      = a * g^e
//   otherwise (t' odd) : sqrt(zeta / x) = a * zeta^((1-m)/2) * g^e
// with e = t' + ((2^47 - t' + 1) >> 1) mod 2^47 (e = 0 for t' = 0), which makes
// the result the *same* root the reference returns (the reference's t is
// 2^47 - t' and its uv equals a * g^t'), not merely a valid one.
//
// The odd-power table of the window method lives in shared memory, one
// column per thread ([slot][limb][thread], conflict free); the same slots are
// recycled for x1..x5 of the discrete-log stage.
#pragma once
#include "fq.cuh"

#define ISQRT_SLOTS 8
#define ISQRT_SMEM_WORDS(block) (ISQRT_SLOTS * 8 * (block))

struct isqrt_smem_t {
  uint32_t* base;  // points at this thread's column
  uint32_t stride;  // blockDim.x
  D377_DI void put(int slot, const fq_t& v) const {
#pragma unroll
    for (int i = 0; i < 8; i++) base[(slot * 8 + i) * stride] = v.l[i];
  }
  D377_DI fq_t get(int slot) const {
    fq_t v;
#pragma unroll
    for (int i = 0; i < 8; i++) v.l[i] = base[(slot * 8 + i) * stride];
    return v;
  }
};

D377_DI isqrt_smem_t isqrt_smem(uint32_t* smem) {
  isqrt_smem_t s;
  s.base = smem + threadIdx.x;
  s.stride = blockDim.x;
  return s;
}

D377_DI fq_r fq_gtab(int k, uint32_t nu) {
  const uint4* p = reinterpret_cast<const uint4*>(&SQRT_GTAB[k][nu & 0xff][0]);
  uint4 lo = __ldg(p), hi = __ldg(p + 1);
  fq_r r;  // generated table, canonical
  r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
  r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
  return r;
}

// reference invsqrt.rs:113 `s_lookup[&alpha]` (HashMap) -> collision-free hash
// (keyed on the low limb of the CANONICAL Montgomery form, hence the reduction)
D377_DI uint32_t fq_slookup(const fq_t& alpha_lazy) {
  const fq_r alpha = fq_reduce(alpha_lazy);
  uint32_t h = (alpha.l[0] * SQRT_SHASH_MUL) >> (32 - SQRT_SHASH_BITS);
  return __ldg(&SQRT_SHASH[h]);
}

D377_DI fq_t fq_sqr_n(fq_t a, int n) {
#pragma unroll 1
  for (int i = 0; i < n; i++) a = fq_sqr(a);
  return a;
}

// a = x^((m-1)/2)
D377_DI fq_t fq_pow_m12(const fq_t& x, const isqrt_smem_t& sm) {
  {
    fq_t x2 = fq_sqr(x);
    fq_t cur = x;
    sm.put(0, cur);
#pragma unroll 1
    for (int j = 1; j < (1 << (POW_WIN - 1)); j++) {
      cur = fq_mul(cur, x2);
      sm.put(j, cur);
    }
  }
  fq_t acc = sm.get(POW_FIRST);
#pragma unroll 1
  for (int op = 0; op < POW_NOPS; op++) {
    int ns = POW_OPS[op][0];
    int j = POW_OPS[op][1];
    acc = fq_sqr_n(acc, ns);
    acc = fq_mul(acc, sm.get(j));
  }
  acc = fq_sqr_n(acc, POW_TAIL);
  return acc;
}

// Discrete log in the 2^47 subgroup <g> (invsqrt.rs:97-153, Sarkar 2020): for
// z in <g> returns the 47-bit t' with z * g^t' = 1, eight bits at a time with the
// reference's tables.  39 S + 15 M + 6 lookups; uses shared slots 1..5.
D377_DI uint64_t fq_dlog47(fq_t z, const isqrt_smem_t& sm) {
  // x5..x1 into slots 5..1 (invsqrt.rs:97-110 with x5 := z)
  sm.put(5, z);
  z = fq_sqr_n(z, 8); sm.put(4, z);
  z = fq_sqr_n(z, 8); sm.put(3, z);
  z = fq_sqr_n(z, 8); sm.put(2, z);
  z = fq_sqr_n(z, 8); sm.put(1, z);
  z = fq_sqr_n(z, 7);  // x0

  // invsqrt.rs:113-153
  uint64_t t = fq_slookup(z);
  fq_t al = fq_mul(sm.get(1), fq_gtab(4, (uint32_t)t));  // products stay < 1.15 q
  t += (uint64_t)fq_slookup(al) << 7;
  al = fq_mul(sm.get(2), fq_gtab(3, (uint32_t)t));
  al = fq_mul(al, fq_gtab(4, (uint32_t)(t >> 8)));
  t += (uint64_t)fq_slookup(al) << 15;
  al = fq_mul(sm.get(3), fq_gtab(2, (uint32_t)t));
  al = fq_mul(al, fq_gtab(3, (uint32_t)(t >> 8)));
  al = fq_mul(al, fq_gtab(4, (uint32_t)(t >> 16)));
  t += (uint64_t)fq_slookup(al) << 23;
  al = fq_mul(sm.get(4), fq_gtab(1, (uint32_t)t));
  al = fq_mul(al, fq_gtab(2, (uint32_t)(t >> 8)));
  al = fq_mul(al, fq_gtab(3, (uint32_t)(t >> 16)));
  al = fq_mul(al, fq_gtab(4, (uint32_t)(t >> 24)));
  t += (uint64_t)fq_slookup(al) << 31;
  al = fq_mul(sm.get(5), fq_gtab(0, (uint32_t)t));
  al = fq_mul(al, fq_gtab(1, (uint32_t)(t >> 8)));
  al = fq_mul(al, fq_gtab(2, (uint32_t)(t >> 16)));
  al = fq_mul(al, fq_gtab(3, (uint32_t)(t >> 24)));
  al = fq_mul(al, fq_gtab(4, (uint32_t)(t >> 32)));
  t += (uint64_t)fq_slookup(al) << 39;
  return t & ((1ull << 47) - 1);
}

// res * g^e for a 47-bit e: one table product per byte (invsqrt.rs:156-163).
D377_DI fq_t fq_mul_gpow(fq_t res, uint64_t e) {
#pragma unroll 1
  for (int k = 0; k < 6; k++) res = fq_mul(res, fq_gtab(k, (uint32_t)(e >> (8 * k))));
  return res;
}

// returns was_square; `out` is the reference's second return value.
D377_DI bool fq_isqrt(fq_t& out, const fq_t& x, const isqrt_smem_t& sm) {
  const bool x_zero = fq_is_zero(x);
  fq_t a = fq_pow_m12(x, sm);
  const uint64_t t = fq_dlog47(fq_mul(fq_sqr(a), x), sm);  // x^m = g^(-t)

  const bool odd = t & 1;
  const uint64_t tref = ((1ull << 47) - t) & ((1ull << 47) - 1);
  const uint64_t e = (t + ((tref + 1) >> 1)) & ((1ull << 47) - 1);

  fq_t res = fq_mul(a, fq_select(odd, fq_const(FQ_ZETA_NS), fq_one()));
  res = fq_mul_gpow(res, e);

  // invsqrt.rs:84-86: den == 0 -> (false, 0)
  out = fq_select(x_zero, fq_zero(), res);
  return !odd && !x_zero;
}

// x^(2^47 - 1) by the 2^k - 1 doubling chain (46 S + 9 M); the reference's
// `den.pow(&[(1 << 47) - 1])`, invsqrt.rs:88-89.
D377_DI fq_t fq_pow_2_47_m1(const fq_t& x) {
  fq_t a2 = fq_mul(fq_sqr(x), x);
  fq_t a4 = fq_mul(fq_sqr_n(a2, 2), a2);
  fq_t a8 = fq_mul(fq_sqr_n(a4, 4), a4);
  fq_t a16 = fq_mul(fq_sqr_n(a8, 8), a8);
  fq_t a32 = fq_mul(fq_sqr_n(a16, 16), a16);
  fq_t r = fq_mul(fq_sqr_n(a32, 8), a8);   // 2^40 - 1
  r = fq_mul(fq_sqr_n(r, 4), a4);          // 2^44 - 1
  r = fq_mul(fq_sqr_n(r, 2), a2);          // 2^46 - 1
  return fq_mul(fq_sqr(r), x);             // 2^47 - 1
}

// Fq::sqrt_ratio_zeta(num, den) for a general ratio, step by step as the reference
// computes it (invsqrt.rs:75-166), so that the very same root comes out:
// (true, sqrt(num/den)), (true, 0) if num = 0, (false, 0) if den = 0, else
// (false, sqrt(zeta*num/den)).
D377_DI bool fq_sqrt_ratio_zeta(fq_t& out, const fq_t& num, const fq_t& den, const isqrt_smem_t& sm) {
  const bool num_zero = fq_is_zero(num), den_zero = fq_is_zero(den);
  fq_t s = fq_pow_2_47_m1(den);                    // :88-89
  fq_t t = fq_mul(fq_sqr(s), den);                 // :90
  fq_t w = fq_mul(fq_pow_m12(fq_mul(num, t), sm), s);  // :91
  fq_t v = fq_mul(w, den), uv = fq_mul(w, num);    // :93-94
  // :97-153 with the reference's own x5 = uv * v, so this is the reference's t
  const uint64_t t47 = fq_dlog47(fq_mul(uv, v), sm);
  const bool odd = t47 & 1;                        // q0' & 1
  fq_t res = fq_mul(uv, fq_select(odd, fq_const(FQ_ZETA_NS), fq_one()));  // :156 nonsquare_lookup
  res = fq_mul_gpow(res, (t47 + 1) >> 1);          // :155-163
  out = fq_select(num_zero || den_zero, fq_zero(), res);   // :81-86
  return num_zero || (!den_zero && !odd);
}

// 1 / x = x^(q-2), plain MSB-first square-and-multiply (252 S + ~125 M); used once per
// batch by the Montgomery-trick normalisations and by the table builders.  0 -> 0.
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

// 1 / x by the binary extended Euclidean algorithm (variable time, like every other entry
// point of this library).  ~380 halvings + ~200 subtractions of 8-limb integers: no
// multiplications at all, so the ~25 k ALU instructions neither load the multiply pipe nor
// take the 380 dependent Montgomery products of fq_inv; for the lone warp that inverts a
// CTA's product in the Montgomery-trick normalisations the latency is ~4x lower.
// Montgomery domain: the input is X = a R; starting the cofactor at R^2 instead of 1 makes
// the result R^2 X^-1 = a^-1 R directly.  0 -> 0.
D377_DI bool u256_is_one(const uint32_t (&u)[8]) {
  return ((u[0] ^ 1u) | u[1] | u[2] | u[3] | u[4] | u[5] | u[6] | u[7]) == 0;
}
D377_DI void u256_shr1(uint32_t (&u)[8]) {
#pragma unroll
  for (int i = 0; i < 7; i++) u[i] = __funnelshift_r(u[i], u[i + 1], 1);
  u[7] >>= 1;
}
// r = a - b, returns the borrow (1 if a < b)
D377_DI uint32_t u256_sub(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t (&b)[8]) {
  uint32_t bw;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(bw)
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
        "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
  return bw & 1u;
}
// x += q & mask
D377_DI void u256_add_q_masked(uint32_t (&x)[8], uint32_t mask) {
  asm("add.cc.u32 %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, %12;\n\t"
      "addc.cc.u32 %5, %5, %13;\n\t"
      "addc.cc.u32 %6, %6, %14;\n\t"
      "addc.u32 %7, %7, %15;"
      : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7])
      : "r"(Q0 & mask), "r"(Q1 & mask), "r"(Q2 & mask), "r"(Q3 & mask), "r"(Q4 & mask),
        "r"(Q5 & mask), "r"(Q6 & mask), "r"(Q7 & mask));
}
// x in [0, q) -> x / 2 mod q
D377_DI void u256_half_mod(uint32_t (&x)[8]) {
  u256_add_q_masked(x, 0u - (x[0] & 1u));   // < 2q < 2^254: no carry out
  u256_shr1(x);
}
// x = x - y mod q for x, y in [0, q)
D377_DI void u256_sub_mod(uint32_t (&x)[8], const uint32_t (&y)[8]) {
  uint32_t bw = u256_sub(x, x, y);
  u256_add_q_masked(x, 0u - bw);
}

D377_DI fq_t fq_inv_vartime(const fq_t& x) {
  const fq_r xr = fq_reduce(x);
  uint32_t u[8], v[8] = {Q0, Q1, Q2, Q3, Q4, Q5, Q6, Q7}, x1[8], x2[8];
  uint32_t nz = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    u[i] = xr.l[i];
    nz |= xr.l[i];
    x1[i] = FQ_R2[i];
    x2[i] = 0;
  }
  if (nz == 0) return fq_zero();
  // (u, v) stay coprime to each other's odd parts and one of them reaches 1 for every
  // x in [1, q); the step bound only guards against inputs outside the type's contract
  // (u = 0 would otherwise halve forever).
  int guard = 0;
#pragma unroll 1
  while (!u256_is_one(u) && !u256_is_one(v) && ++guard < 1024) {
    uint32_t unz = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) unz |= u[i];
    if (unz == 0) break;
#pragma unroll 1
    while (!(u[0] & 1u)) {
      u256_shr1(u);
      u256_half_mod(x1);
    }
#pragma unroll 1
    while (!(v[0] & 1u)) {
      u256_shr1(v);
      u256_half_mod(x2);
    }
    uint32_t d[8];
    if (u256_sub(d, u, v) == 0) {   // u >= v
#pragma unroll
      for (int i = 0; i < 8; i++) u[i] = d[i];
      u256_sub_mod(x1, x2);
    } else {
      u256_sub(v, v, u);
      u256_sub_mod(x2, x1);
    }
  }
  const bool from_u = u256_is_one(u);
  fq_r r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = from_u ? x1[i] : x2[i];
  return r;
}
