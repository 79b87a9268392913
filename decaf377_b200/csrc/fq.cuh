// Fq = GF(q), q = decaf377 base field (BLS12-377 scalar field), 253 bits.
//
// Replaces the reference's L1/L2 field layer for the batch path:
//   src/fields/fq/u32/fiat.rs:162 (fq_mul), :1360 (fq_square), :2555 (fq_add),
//   :2646 (fq_sub), :2725 (fq_opp), :2800 (fq_from_montgomery),
//   :3584 (fq_to_montgomery); src/sign.rs:3-23.
//
// Representation: one element per thread, 8 x 32-bit limbs in registers,
// Montgomery form with R = 2^256 (byte-identical to both reference backends,
// fq/u32/wrapper.rs:93-104).
//
// LAZY REDUCTION WITH STATICALLY CHECKED BOUNDS.  q < 2^253, so eight limbs hold
// any value below 13.71 q.  The reference (fiat) fully reduces after every
// operation; here the conditional subtractions are dropped and every value carries
// an upper bound in its *type*: fqb<B> holds an integer < B/1000 * q that is
// congruent to the element it stands for.
//   mul / sqr   a < A q, b < B q  ->  < (1 + A B q/R) q,   q/R = 0.07293
//   add         A + B
//   sub         a - b + K q with K = ceil(B)  ->  A + K
//   cond. sub   fq_csub<K>: subtract K q if >= K q
// Every operation static_asserts that its result stays below the 8-limb capacity, so
// an arithmetic overflow is a compile error, not a data-dependent bug.  fq_t =
// fqb<2000> (< 2q) is the storage class: everything kept in memory or passed between
// functions is < 2q; fq_reduce() gives the canonical representative (< q) where the
// ABI, a comparison or a table lookup needs it.  On B200 this is worth 10-15 % of
// the IMAD.WIDE issue rate (tools/ub_field.cu): the multiply pipe is the limiter
// and the dropped SEL/IADD3 chains competed with it for issue slots.
//
// Multiplication is a word-serial Montgomery product whose 32x32->64 partial
// products are issued as IMAD.WIDE.U32 with predicate carry chains: every
// `mad.lo.cc / madc.hi.cc` pair on the same operands below is fused by ptxas
// into one `IMAD.WIDE.U32[.X] Rd, P, Ra, Rb, Rc[, P]`.  Products at even limb
// positions and at odd limb positions are accumulated in two separate
// 8-limb accumulators so that each carry chain only ever touches aligned
// 64-bit register pairs.
#pragma once
#include <cstdint>

#define D377_CONST static __device__ __constant__ const
#define D377_TABLE static __device__ const
#include "constants.inc"

#define D377_DI __device__ __forceinline__

// ---- bounds (milli-q) ------------------------------------------------------
constexpr int FQ_CAP = 13700;  // 2^256 / q = 13.712: every value must stay below this
constexpr int FQ_RAWB = 13712; // an arbitrary 256-bit string (wire input before range checks)
__host__ __device__ constexpr int fq_bd_mul(int A, int B) {
  // (1 + A B q/R) q with q/R = 0.0729278 rounded up to 0.07293
  return 1000 + (int)(((long long)A * (long long)B * 7293LL + 99999999LL) / 100000000LL);
}
__host__ __device__ constexpr int fq_bd_k(int B) { return (B + 999) / 1000; }  // smallest K with K q >= bound
__host__ __device__ constexpr int fq_bd_max(int A, int B) { return A > B ? A : B; }

template <int B>
struct fqb {
  uint32_t l[8];
  fqb() = default;
  // widening is implicit, narrowing does not compile
  template <int A>
  D377_DI fqb(const fqb<A>& o) {
    static_assert(A <= B, "fq bound can only widen; reduce (fq_fold / fq_reduce) first");
#pragma unroll
    for (int i = 0; i < 8; i++) l[i] = o.l[i];
  }
};
using fq_t = fqb<2000>;      // storage class: < 2q
using fq_r = fqb<1000>;      // canonical representative: < q
using fq_raw_t = fqb<FQ_RAWB>;

// Re-type without a check: only where the bound is known for a reason the types cannot
// express (a value read back from memory, a proof in the comment next to the call).
template <int B, int A>
D377_DI fqb<B> fq_assume(const fqb<A>& a) {
  fqb<B> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = a.l[i];
  return r;
}

// q limbs as literals so that ptxas can fold them into immediates.
#define Q0 0x00000001u
#define Q1 0x0a118000u
#define Q2 0xd0000001u
#define Q3 0x59aa76feu
#define Q4 0x5c37b001u
#define Q5 0x60b44d1eu
#define Q6 0x9a2ca556u
#define Q7 0x12ab655eu

// limb i of K * q (K <= 13), evaluated at compile time
__host__ __device__ constexpr uint32_t fq_kq_limb(int K, int i) {
  const uint32_t Q[8] = {Q0, Q1, Q2, Q3, Q4, Q5, Q6, Q7};
  unsigned long long carry = 0;
  uint32_t r = 0;
  for (int j = 0; j <= i; j++) {
    unsigned long long t = (unsigned long long)Q[j] * (unsigned)K + carry;
    r = (uint32_t)t;
    carry = t >> 32;
  }
  return r;
}

D377_DI fq_r fq_const(const uint32_t (&c)[8]) {
  fq_r r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = c[i];
  return r;
}

D377_DI fq_r fq_zero() {
  fq_r r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = 0;
  return r;
}

D377_DI fq_r fq_one() { return fq_const(FQ_ONE); }

// r = c ? a : b
template <int A, int B>
D377_DI fqb<fq_bd_max(A, B)> fq_select(bool c, const fqb<A>& a, const fqb<B>& b) {
  fqb<fq_bd_max(A, B)> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = c ? a.l[i] : b.l[i];
  return r;
}

// t - K q, returns the borrow mask (0xffffffff if t < K q).
template <int K>
D377_DI uint32_t fq_sub_kq_raw(uint32_t (&d)[8], const uint32_t (&t)[8]) {
  constexpr uint32_t k0 = fq_kq_limb(K, 0), k1 = fq_kq_limb(K, 1), k2 = fq_kq_limb(K, 2),
                     k3 = fq_kq_limb(K, 3), k4 = fq_kq_limb(K, 4), k5 = fq_kq_limb(K, 5),
                     k6 = fq_kq_limb(K, 6), k7 = fq_kq_limb(K, 7);
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
      : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]),
        "=r"(d[7]), "=r"(bw)
      : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]),
        "n"(k0), "n"(k1), "n"(k2), "n"(k3), "n"(k4), "n"(k5), "n"(k6), "n"(k7));
  return bw;
}

// Conditional subtraction of K q: a < A q  ->  < max(K, A - K) q.
template <int K, int A>
D377_DI fqb<fq_bd_max(1000 * K, A - 1000 * K)> fq_csub(const fqb<A>& a) {
  static_assert(A <= FQ_RAWB, "bound");
  uint32_t d[8];
  uint32_t bw = fq_sub_kq_raw<K>(d, a.l);
  fqb<fq_bd_max(1000 * K, A - 1000 * K)> r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = bw ? a.l[i] : d[i];
  return r;
}

// Bring any bound back to the storage class (< 2q): ceil((A - 2000) / 2000)
// conditional subtractions of 2q.
template <int A>
D377_DI fq_t fq_fold(const fqb<A>& a) {
  if constexpr (A <= 2000) {
    return a;
  } else {
    return fq_fold(fq_csub<2>(a));
  }
}

// Canonical representative (< q).
template <int A>
D377_DI fq_r fq_reduce(const fqb<A>& a) {
  if constexpr (A <= 1000) {
    return a;
  } else {
    return fq_csub<1>(fq_fold(a));
  }
}

template <int A>
D377_DI bool fq_is_zero(const fqb<A>& a) {
  fq_r c = fq_reduce(a);
  uint32_t o = c.l[0];
#pragma unroll
  for (int i = 1; i < 8; i++) o |= c.l[i];
  return o == 0;
}

template <int A, int B>
D377_DI bool fq_eq(const fqb<A>& a, const fqb<B>& b) {
  fq_r x = fq_reduce(a), y = fq_reduce(b);
  uint32_t o = x.l[0] ^ y.l[0];
#pragma unroll
  for (int i = 1; i < 8; i++) o |= x.l[i] ^ y.l[i];
  return o == 0;
}

// fiat.rs:2555 (fq_add), without the final conditional subtraction
template <int A, int B>
D377_DI fqb<A + B> fq_add(const fqb<A>& a, const fqb<B>& b) {
  static_assert(A + B <= FQ_CAP, "fq_add would overflow 8 limbs: fold an operand first");
  fqb<A + B> t;
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(t.l[0]), "=r"(t.l[1]), "=r"(t.l[2]), "=r"(t.l[3]), "=r"(t.l[4]), "=r"(t.l[5]),
        "=r"(t.l[6]), "=r"(t.l[7])
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]),
        "r"(a.l[6]), "r"(a.l[7]), "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]),
        "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
  return t;
}

// fiat.rs:2646 (fq_sub): a - b + K q, K = ceil(bound of b); never negative, no select.
template <int A, int B>
D377_DI fqb<A + 1000 * fq_bd_k(B)> fq_sub(const fqb<A>& a, const fqb<B>& b) {
  constexpr int K = fq_bd_k(B);
  static_assert(A + 1000 * K <= FQ_CAP, "fq_sub would overflow 8 limbs: fold an operand first");
  constexpr uint32_t k0 = fq_kq_limb(K, 0), k1 = fq_kq_limb(K, 1), k2 = fq_kq_limb(K, 2),
                     k3 = fq_kq_limb(K, 3), k4 = fq_kq_limb(K, 4), k5 = fq_kq_limb(K, 5),
                     k6 = fq_kq_limb(K, 6), k7 = fq_kq_limb(K, 7);
  fqb<A + 1000 * K> t;
  // (a - b) mod 2^256, then + K q: the true value is in [0, 2^256), so the wrap-around of
  // the first chain is undone by the second.
  asm("sub.cc.u32 %0, %8, %16;\n\t"
      "subc.cc.u32 %1, %9, %17;\n\t"
      "subc.cc.u32 %2, %10, %18;\n\t"
      "subc.cc.u32 %3, %11, %19;\n\t"
      "subc.cc.u32 %4, %12, %20;\n\t"
      "subc.cc.u32 %5, %13, %21;\n\t"
      "subc.cc.u32 %6, %14, %22;\n\t"
      "subc.u32 %7, %15, %23;"
      : "=r"(t.l[0]), "=r"(t.l[1]), "=r"(t.l[2]), "=r"(t.l[3]), "=r"(t.l[4]), "=r"(t.l[5]),
        "=r"(t.l[6]), "=r"(t.l[7])
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]),
        "r"(a.l[6]), "r"(a.l[7]), "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]),
        "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
  asm("add.cc.u32 %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, %12;\n\t"
      "addc.cc.u32 %5, %5, %13;\n\t"
      "addc.cc.u32 %6, %6, %14;\n\t"
      "addc.u32 %7, %7, %15;"
      : "+r"(t.l[0]), "+r"(t.l[1]), "+r"(t.l[2]), "+r"(t.l[3]), "+r"(t.l[4]), "+r"(t.l[5]),
        "+r"(t.l[6]), "+r"(t.l[7])
      : "n"(k0), "n"(k1), "n"(k2), "n"(k3), "n"(k4), "n"(k5), "n"(k6), "n"(k7));
  return t;
}

// fiat.rs:2725 (fq_opp): K q - a.  a = 0 gives exactly K q, hence the +1 in the bound.
template <int A>
D377_DI fqb<1000 * fq_bd_k(A) + 1> fq_neg(const fqb<A>& a) {
  constexpr int K = fq_bd_k(A) < 1 ? 1 : fq_bd_k(A);
  constexpr uint32_t k0 = fq_kq_limb(K, 0), k1 = fq_kq_limb(K, 1), k2 = fq_kq_limb(K, 2),
                     k3 = fq_kq_limb(K, 3), k4 = fq_kq_limb(K, 4), k5 = fq_kq_limb(K, 5),
                     k6 = fq_kq_limb(K, 6), k7 = fq_kq_limb(K, 7);
  fqb<1000 * fq_bd_k(A) + 1> t;
  asm("sub.cc.u32 %0, %8, %16;\n\t"
      "subc.cc.u32 %1, %9, %17;\n\t"
      "subc.cc.u32 %2, %10, %18;\n\t"
      "subc.cc.u32 %3, %11, %19;\n\t"
      "subc.cc.u32 %4, %12, %20;\n\t"
      "subc.cc.u32 %5, %13, %21;\n\t"
      "subc.cc.u32 %6, %14, %22;\n\t"
      "subc.u32 %7, %15, %23;"
      : "=r"(t.l[0]), "=r"(t.l[1]), "=r"(t.l[2]), "=r"(t.l[3]), "=r"(t.l[4]), "=r"(t.l[5]),
        "=r"(t.l[6]), "=r"(t.l[7])
      : "n"(k0), "n"(k1), "n"(k2), "n"(k3), "n"(k4), "n"(k5), "n"(k6), "n"(k7), "r"(a.l[0]),
        "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]));
  return t;
}

template <int A>
D377_DI fqb<2 * A> fq_dbl(const fqb<A>& a) { return fq_add(a, a); }

// ---------------------------------------------------------------------------
// Montgomery product core.
//
// Accumulator invariant between rows: T = E + O * 2^32 where E = ev[0..8)
// sits at limb positions 0..7 and O = od[0..8) at positions 1..8.
//
// No-overflow conditions with lazily reduced operands (a < A q is the operand whose limbs
// multiply every b_i): between rows T < a + q, and T * 2^32 must fit the 9-limb (ev, od)
// pair, i.e. A + 1 <= 13.7; the result is < q + a b / R.
//
// Cost accounting (IMAD.WIDE.U32 issues at 32 lanes/clk/SM on sm_100, half the
// rate of a 32-bit IMAD, so the count of wide multiplies is the cost):
//   fq_mul  64 (operand rows) + 56 (reduction rows; the q0 = 1 column is an add) = 120
//   fq_sqr  28 (off-diagonal) + 8 (diagonal) + 56                                 =  92
// The modulus limbs are read from a *mutable* __constant__ array (see fq_mod_t::z
// for why the row factor must not look like a plain negation to ptxas).
// ---------------------------------------------------------------------------
static __device__ __constant__ uint32_t FQ_QRT[8] = {Q0, Q1, Q2, Q3, Q4, Q5, Q6, Q7};

struct fq_mod_t {
  uint32_t q1, q2, q3, q4, q5, q6, q7;
  // q0 - 1: zero at run time, but not a compile-time constant.  The Montgomery
  // factor is formed as m = z - t0; with a literal zero ptxas folds the negation
  // into the multiplies and, IMAD.WIDE having no negated-operand form, falls back
  // to IMAD.X + IMAD.HI.U32.X for the whole row.
  uint32_t z;
};
D377_DI fq_mod_t fq_mod() {
  fq_mod_t q;
  q.q1 = FQ_QRT[1]; q.q2 = FQ_QRT[2]; q.q3 = FQ_QRT[3]; q.q4 = FQ_QRT[4];
  q.q5 = FQ_QRT[5]; q.q6 = FQ_QRT[6]; q.q7 = FQ_QRT[7];
  q.z = FQ_QRT[0] - 1u;
  return q;
}

// acc(4 aligned 64-bit lanes) += {x0,x2,x4,x6} * y; the carry-out is added to `top` (the
// limb above acc[7]) inside the same chain: one IADD3.X instead of a captured carry.
#define D377_CMAD4(acc, x0, x2, x4, x6, y, top)                                             \
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"                                                   \
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"                                                  \
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"                                                 \
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"                                                 \
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"                                                 \
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"                                                 \
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"                                                 \
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"                                                 \
      "addc.u32 %8, %8, 0;"                                                                  \
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),  \
        "+r"(acc[6]), "+r"(acc[7]), "+r"(top)                                                \
      : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y))

// One Montgomery reduction row on (ev, od): adds m*q with m = -ev[0] (q = 1 mod 2^32
// so -q^-1 = 0xffffffff, fiat.rs:223) making ev[0] zero.  `cy` (0/1) is a pending
// carry of weight 2^32 that enters the odd chain.
template <bool kCy>
D377_DI void fq_redc_row(uint32_t (&ev)[8], uint32_t (&od)[8], const fq_mod_t& q, uint32_t cy) {
  uint32_t m = q.z - ev[0];
  // odd positions: q1,q3,q5,q7; cannot overflow (O*2^32 <= T < 2^288)
  if (kCy) {
    asm("add.cc.u32 %8, %8, 0xffffffff;\n\t"
        "madc.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.u32 %7, %12, %13, %7;"
        : "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]),
          "+r"(od[6]), "+r"(od[7]), "+r"(cy)
        : "r"(q.q1), "r"(q.q3), "r"(q.q5), "r"(q.q7), "r"(m));
  } else {
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(od[0]), "+r"(od[1]), "+r"(od[2]), "+r"(od[3]), "+r"(od[4]), "+r"(od[5]),
          "+r"(od[6]), "+r"(od[7])
        : "r"(q.q1), "r"(q.q3), "r"(q.q5), "r"(q.q7), "r"(m));
  }
  // even positions: q0 = 1 (ev[0] + m = 0 with carry ev[0] != 0), then q2,q4,q6
  asm("add.cc.u32 %0, %0, %12;\n\t"
      "addc.cc.u32 %1, %1, 0;\n\t"
      "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
      "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
      "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
      "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
      "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
      "madc.hi.cc.u32 %7, %11, %12, %7;\n\t"
      "addc.u32 %8, %8, 0;"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]),
        "+r"(ev[6]), "+r"(ev[7]), "+r"(od[7])
      : "r"(q.q2), "r"(q.q4), "r"(q.q6), "r"(m));
}

// Shift (ev, od) right by one limb (ev[0] == 0 on entry) and add a*b_i.
// After the shift the old odd accumulator is aligned to even positions; the
// old even accumulator moves down one 64-bit lane and becomes the odd one,
// except for its limb 1, which lands on position 0 and is folded into od[0]
// -- the carry of that fold has weight 2^32, i.e. it is exactly the carry-in
// of the new odd chain.
D377_DI void fq_mul_row_shift(uint32_t (&ev)[8], uint32_t (&od)[8], const uint32_t (&al)[8], uint32_t bi) {
  uint32_t nod[8];
  asm("add.cc.u32 %0, %0, %9;\n\t"
      "madc.lo.cc.u32 %1, %16, %20, %10;\n\t"
      "madc.hi.cc.u32 %2, %16, %20, %11;\n\t"
      "madc.lo.cc.u32 %3, %17, %20, %12;\n\t"
      "madc.hi.cc.u32 %4, %17, %20, %13;\n\t"
      "madc.lo.cc.u32 %5, %18, %20, %14;\n\t"
      "madc.hi.cc.u32 %6, %18, %20, %15;\n\t"
      "madc.lo.cc.u32 %7, %19, %20, 0;\n\t"
      "madc.hi.u32 %8, %19, %20, 0;"
      : "+r"(od[0]), "=&r"(nod[0]), "=&r"(nod[1]), "=&r"(nod[2]), "=&r"(nod[3]), "=&r"(nod[4]),
        "=&r"(nod[5]), "=&r"(nod[6]), "=&r"(nod[7])
      : "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
        "r"(al[1]), "r"(al[3]), "r"(al[5]), "r"(al[7]), "r"(bi));
  D377_CMAD4(od, al[0], al[2], al[4], al[6], bi, nod[7]);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    ev[i] = od[i];
    od[i] = nod[i];
  }
}

// Shift without a product row: returns the pending carry for the next odd chain.
D377_DI uint32_t fq_shift_only(uint32_t (&ev)[8], uint32_t (&od)[8]) {
  uint32_t cy;
  asm("add.cc.u32 %0, %0, %2;\n\t"
      "addc.u32 %1, 0, 0;"
      : "+r"(od[0]), "=r"(cy)
      : "r"(ev[1]));
  uint32_t t2 = ev[2], t3 = ev[3], t4 = ev[4], t5 = ev[5], t6 = ev[6], t7 = ev[7];
#pragma unroll
  for (int i = 0; i < 8; i++) ev[i] = od[i];
  od[0] = t2; od[1] = t3; od[2] = t4; od[3] = t5; od[4] = t6; od[5] = t7; od[6] = 0; od[7] = 0;
  return cy;
}

// Collapse T = E + O*2^32 after the last reduction row (one more limb shift) and
// optionally add an 8-limb `hi` (the upper half of a 512-bit product).  No final
// subtraction: the caller's type carries the bound.
D377_DI void fq_mont_collapse(uint32_t (&t)[8], uint32_t (&ev)[8], uint32_t (&od)[8]) {
  asm("add.cc.u32 %0, %8, %15;\n\t"
      "addc.cc.u32 %1, %9, %16;\n\t"
      "addc.cc.u32 %2, %10, %17;\n\t"
      "addc.cc.u32 %3, %11, %18;\n\t"
      "addc.cc.u32 %4, %12, %19;\n\t"
      "addc.cc.u32 %5, %13, %20;\n\t"
      "addc.cc.u32 %6, %14, %21;\n\t"
      "addc.u32 %7, 0, %22;"
      : "=&r"(t[0]), "=&r"(t[1]), "=&r"(t[2]), "=&r"(t[3]), "=&r"(t[4]), "=&r"(t[5]),
        "=&r"(t[6]), "=&r"(t[7])
      : "r"(ev[1]), "r"(ev[2]), "r"(ev[3]), "r"(ev[4]), "r"(ev[5]), "r"(ev[6]), "r"(ev[7]),
        "r"(od[0]), "r"(od[1]), "r"(od[2]), "r"(od[3]), "r"(od[4]), "r"(od[5]), "r"(od[6]),
        "r"(od[7]));
}

D377_DI void fq_add_hi(uint32_t (&t)[8], const uint32_t (&hi)[8]) {
  asm("add.cc.u32 %0, %0, %8;\n\t"
      "addc.cc.u32 %1, %1, %9;\n\t"
      "addc.cc.u32 %2, %2, %10;\n\t"
      "addc.cc.u32 %3, %3, %11;\n\t"
      "addc.cc.u32 %4, %4, %12;\n\t"
      "addc.cc.u32 %5, %5, %13;\n\t"
      "addc.cc.u32 %6, %6, %14;\n\t"
      "addc.u32 %7, %7, %15;"
      : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]),
        "+r"(t[7])
      : "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]),
        "r"(hi[7]));
}

// fiat.rs:162 (fq_mul): r = a * b / R mod q, r < q + a b / R.  Row-interleaved (CIOS) form:
// 64 + 56 wide multiplies.
template <int A, int B>
D377_DI fqb<fq_bd_mul(A, B)> fq_mul_cios(const fqb<A>& a, const fqb<B>& b) {
  static_assert(A + 1000 <= FQ_CAP, "fq_mul: first operand too large for the row accumulator");
  static_assert(fq_bd_mul(A, B) <= FQ_CAP, "fq_mul result would overflow 8 limbs");
  const fq_mod_t q = fq_mod();
  uint32_t ev[8], od[8];
  // row 0: plain products
  asm("mul.lo.u32 %0, %8, %12;\n\t"
      "mul.hi.u32 %1, %8, %12;\n\t"
      "mul.lo.u32 %2, %9, %12;\n\t"
      "mul.hi.u32 %3, %9, %12;\n\t"
      "mul.lo.u32 %4, %10, %12;\n\t"
      "mul.hi.u32 %5, %10, %12;\n\t"
      "mul.lo.u32 %6, %11, %12;\n\t"
      "mul.hi.u32 %7, %11, %12;"
      : "=&r"(ev[0]), "=&r"(ev[1]), "=&r"(ev[2]), "=&r"(ev[3]), "=&r"(ev[4]), "=&r"(ev[5]),
        "=&r"(ev[6]), "=&r"(ev[7])
      : "r"(a.l[0]), "r"(a.l[2]), "r"(a.l[4]), "r"(a.l[6]), "r"(b.l[0]));
  asm("mul.lo.u32 %0, %8, %12;\n\t"
      "mul.hi.u32 %1, %8, %12;\n\t"
      "mul.lo.u32 %2, %9, %12;\n\t"
      "mul.hi.u32 %3, %9, %12;\n\t"
      "mul.lo.u32 %4, %10, %12;\n\t"
      "mul.hi.u32 %5, %10, %12;\n\t"
      "mul.lo.u32 %6, %11, %12;\n\t"
      "mul.hi.u32 %7, %11, %12;"
      : "=&r"(od[0]), "=&r"(od[1]), "=&r"(od[2]), "=&r"(od[3]), "=&r"(od[4]), "=&r"(od[5]),
        "=&r"(od[6]), "=&r"(od[7])
      : "r"(a.l[1]), "r"(a.l[3]), "r"(a.l[5]), "r"(a.l[7]), "r"(b.l[0]));
  fq_redc_row<false>(ev, od, q, 0u);
#pragma unroll
  for (int i = 1; i < 8; i++) {
    fq_mul_row_shift(ev, od, a.l, b.l[i]);
    fq_redc_row<false>(ev, od, q, 0u);
  }
  fqb<fq_bd_mul(A, B)> r;
  fq_mont_collapse(r.l, ev, od);
  return r;
}

// a * K for a small integer constant K (the curve constants 2d = 6042, 4d, ...): in
// Montgomery form that is the plain integer product, so no Montgomery reduction is needed,
// only a 9-limb product (8 wide multiplies) brought back below 3q by one quotient estimate:
//   t = a K < 2^267,  x = t >> 235 (< 2^32),  h = x / (floor(q / 2^235) + 1) <= t / q,
//   r = t - h q  with  t / q - h < 1.3  (the estimate loses < 0.2 to the truncations of x
//   and q, 1 to the floor),  so 0 <= r < 2.3 q.
// 8 + 8 wide multiplies and one division by a constant instead of the 120 of fq_mul.
template <uint32_t K, int A>
D377_DI fqb<3000> fq_mul_small(const fqb<A>& a) {
  static_assert((long long)A * K <= 28000LL * 1000LL, "fq_mul_small: a K must stay below 2^267");
  constexpr uint32_t QS = (Q7 >> 11) + 1u;   // floor(q / 2^235) + 1
  uint32_t t[9];
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)a.l[i] * K;
    t[i] = (uint32_t)c;
    c >>= 32;
  }
  t[8] = (uint32_t)c;
  const uint32_t x = (t[8] << 21) | (t[7] >> 11);
  const uint32_t h = x / QS;
  const uint32_t Q[8] = {Q0, Q1, Q2, Q3, Q4, Q5, Q6, Q7};
  uint32_t m[8];
  c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)Q[i] * h;
    m[i] = (uint32_t)c;
    c >>= 32;
  }
  fqb<3000> r;
  asm("sub.cc.u32 %0, %8, %16;\n\t"
      "subc.cc.u32 %1, %9, %17;\n\t"
      "subc.cc.u32 %2, %10, %18;\n\t"
      "subc.cc.u32 %3, %11, %19;\n\t"
      "subc.cc.u32 %4, %12, %20;\n\t"
      "subc.cc.u32 %5, %13, %21;\n\t"
      "subc.cc.u32 %6, %14, %22;\n\t"
      "subc.u32 %7, %15, %23;"
      : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]),
        "=r"(r.l[6]), "=r"(r.l[7])
      : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]),
        "r"(m[0]), "r"(m[1]), "r"(m[2]), "r"(m[3]), "r"(m[4]), "r"(m[5]), "r"(m[6]), "r"(m[7]));
  return r;
}

// Montgomery reduction of a 512-bit value t[0..16): t / R mod q.  Eight reduction
// rows run on the low half only (their result is <= q); the high half is added at
// the end: r <= q + t / R.
D377_DI void fq_redc16(uint32_t (&r)[8], const uint32_t (&t)[16]) {
  const fq_mod_t q = fq_mod();
  uint32_t ev[8], od[8], hi[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    ev[i] = t[i];
    od[i] = 0;
    hi[i] = t[8 + i];
  }
  uint32_t cy = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if (i == 0) fq_redc_row<false>(ev, od, q, 0u); else fq_redc_row<true>(ev, od, q, cy);
    if (i < 7) cy = fq_shift_only(ev, od);
  }
  fq_mont_collapse(r, ev, od);
  fq_add_hi(r, hi);
}

// ---- Karatsuba multiplication: 48 + 56 = 104 wide multiplies instead of 120 ------------
// (An experiment that did not pay; see fq_mul below.)  The wide multiply is the scarce
// resource (31 / clk / SM against >= 64 for IADD3), so one level of Karatsuba on the 8 x 8
// limb product -- three 4 x 4 products (16 wide multiplies each) and ~95 additions instead of
// 64 wide multiplies -- would trade the plentiful instruction for the scarce one.  With
// a = a_lo + 2^128 a_hi, b likewise:
//   P0 = a_lo b_lo,  P2 = a_hi b_hi,  Pm = (a_lo + a_hi)(b_lo + b_hi),  P1 = Pm - P0 - P2,
//   T = P0 + 2^128 P1 + 2^256 P2,   r = fq_redc16(T).
// The sums a_lo + a_hi, b_lo + b_hi carry a 129th bit each; those are folded in by masked
// additions, so the middle product is a 4 x 4 product as well.

// p (8 limbs) = a (4 limbs) * b (4 limbs).  Products landing on even and on odd limb
// positions go to two accumulators (aligned 64-bit pairs, IMAD.WIDE carry chains); every
// carry-out lands on a limb no earlier row has touched, so nothing ripples.
D377_DI void fq_mul4x4(uint32_t (&p)[8], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                       uint32_t b1, uint32_t b2, uint32_t b3) {
  uint32_t e0, e1, e2, e3, e4, e5, e6, e7, o1, o2, o3, o4, o5, o6, o7;
  // row b0: plain products
  asm("mul.lo.u32 %0, %4, %6;\n\t mul.hi.u32 %1, %4, %6;\n\t"
      "mul.lo.u32 %2, %5, %6;\n\t mul.hi.u32 %3, %5, %6;"
      : "=&r"(e0), "=&r"(e1), "=&r"(e2), "=&r"(e3)
      : "r"(a0), "r"(a2), "r"(b0));
  asm("mul.lo.u32 %0, %4, %6;\n\t mul.hi.u32 %1, %4, %6;\n\t"
      "mul.lo.u32 %2, %5, %6;\n\t mul.hi.u32 %3, %5, %6;"
      : "=&r"(o1), "=&r"(o2), "=&r"(o3), "=&r"(o4)
      : "r"(a1), "r"(a3), "r"(b0));
  // row b1: a0, a2 -> odd positions 1, 3 (carry -> o5); a1, a3 -> even positions 2, 4
  asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
      "madc.lo.cc.u32 %2, %6, %7, %2;\n\t madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
      "addc.u32 %4, 0, 0;"
      : "+r"(o1), "+r"(o2), "+r"(o3), "+r"(o4), "=r"(o5)
      : "r"(a0), "r"(a2), "r"(b1));
  asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\t madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
      "madc.lo.cc.u32 %2, %5, %6, 0;\n\t madc.hi.u32 %3, %5, %6, 0;"
      : "+r"(e2), "+r"(e3), "=&r"(e4), "=&r"(e5)
      : "r"(a1), "r"(a3), "r"(b1));
  // row b2: a0, a2 -> even positions 2, 4 (carry -> e6); a1, a3 -> odd positions 3, 5
  asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
      "madc.lo.cc.u32 %2, %6, %7, %2;\n\t madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
      "addc.u32 %4, 0, 0;"
      : "+r"(e2), "+r"(e3), "+r"(e4), "+r"(e5), "=r"(e6)
      : "r"(a0), "r"(a2), "r"(b2));
  asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\t madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
      "madc.lo.cc.u32 %2, %5, %6, %2;\n\t madc.hi.u32 %3, %5, %6, 0;"
      : "+r"(o3), "+r"(o4), "+r"(o5), "=&r"(o6)
      : "r"(a1), "r"(a3), "r"(b2));
  // row b3: a0, a2 -> odd positions 3, 5 (carry -> o7); a1, a3 -> even positions 4, 6
  asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
      "madc.lo.cc.u32 %2, %6, %7, %2;\n\t madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
      "addc.u32 %4, 0, 0;"
      : "+r"(o3), "+r"(o4), "+r"(o5), "+r"(o6), "=r"(o7)
      : "r"(a0), "r"(a2), "r"(b3));
  asm("mad.lo.cc.u32 %0, %4, %6, %0;\n\t madc.hi.cc.u32 %1, %4, %6, %1;\n\t"
      "madc.lo.cc.u32 %2, %5, %6, %2;\n\t madc.hi.u32 %3, %5, %6, 0;"
      : "+r"(e4), "+r"(e5), "+r"(e6), "=&r"(e7)
      : "r"(a1), "r"(a3), "r"(b3));
  // p = E + O * 2^32
  p[0] = e0;
  asm("add.cc.u32 %0, %7, %14;\n\t"
      "addc.cc.u32 %1, %8, %15;\n\t"
      "addc.cc.u32 %2, %9, %16;\n\t"
      "addc.cc.u32 %3, %10, %17;\n\t"
      "addc.cc.u32 %4, %11, %18;\n\t"
      "addc.cc.u32 %5, %12, %19;\n\t"
      "addc.u32 %6, %13, %20;"
      : "=&r"(p[1]), "=&r"(p[2]), "=&r"(p[3]), "=&r"(p[4]), "=&r"(p[5]), "=&r"(p[6]), "=&r"(p[7])
      : "r"(e1), "r"(e2), "r"(e3), "r"(e4), "r"(e5), "r"(e6), "r"(e7), "r"(o1), "r"(o2), "r"(o3), "r"(o4),
        "r"(o5), "r"(o6), "r"(o7));
}

template <int A, int B>
D377_DI fqb<fq_bd_mul(A, B)> fq_mul_kara(const fqb<A>& a, const fqb<B>& b) {
  static_assert(fq_bd_mul(A, B) <= FQ_CAP, "fq_mul result would overflow 8 limbs");
  uint32_t p0[8], p2[8], pm[8], pm8, sa[4], sb[4], ca, cb;
  fq_mul4x4(p0, a.l[0], a.l[1], a.l[2], a.l[3], b.l[0], b.l[1], b.l[2], b.l[3]);
  fq_mul4x4(p2, a.l[4], a.l[5], a.l[6], a.l[7], b.l[4], b.l[5], b.l[6], b.l[7]);
  asm("add.cc.u32 %0, %5, %9;\n\t addc.cc.u32 %1, %6, %10;\n\t addc.cc.u32 %2, %7, %11;\n\t"
      "addc.cc.u32 %3, %8, %12;\n\t addc.u32 %4, 0, 0;"
      : "=&r"(sa[0]), "=&r"(sa[1]), "=&r"(sa[2]), "=&r"(sa[3]), "=r"(ca)
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]));
  asm("add.cc.u32 %0, %5, %9;\n\t addc.cc.u32 %1, %6, %10;\n\t addc.cc.u32 %2, %7, %11;\n\t"
      "addc.cc.u32 %3, %8, %12;\n\t addc.u32 %4, 0, 0;"
      : "=&r"(sb[0]), "=&r"(sb[1]), "=&r"(sb[2]), "=&r"(sb[3]), "=r"(cb)
      : "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
  fq_mul4x4(pm, sa[0], sa[1], sa[2], sa[3], sb[0], sb[1], sb[2], sb[3]);
  // the 129th bits: Pm += 2^128 (ca * sb + cb * sa) + 2^256 ca cb
  const uint32_t ma = 0u - ca, mb = 0u - cb;
  asm("add.cc.u32 %0, %0, %5;\n\t addc.cc.u32 %1, %1, %6;\n\t addc.cc.u32 %2, %2, %7;\n\t"
      "addc.cc.u32 %3, %3, %8;\n\t addc.u32 %4, 0, 0;"
      : "+r"(pm[4]), "+r"(pm[5]), "+r"(pm[6]), "+r"(pm[7]), "=r"(pm8)
      : "r"(sb[0] & ma), "r"(sb[1] & ma), "r"(sb[2] & ma), "r"(sb[3] & ma));
  asm("add.cc.u32 %0, %0, %5;\n\t addc.cc.u32 %1, %1, %6;\n\t addc.cc.u32 %2, %2, %7;\n\t"
      "addc.cc.u32 %3, %3, %8;\n\t addc.u32 %4, %4, %9;"
      : "+r"(pm[4]), "+r"(pm[5]), "+r"(pm[6]), "+r"(pm[7]), "+r"(pm8)
      : "r"(sa[0] & mb), "r"(sa[1] & mb), "r"(sa[2] & mb), "r"(sa[3] & mb), "r"(ca & cb));
  // P1 = Pm - P0 - P2 (9 limbs, never negative)
  asm("sub.cc.u32 %0, %0, %9;\n\t subc.cc.u32 %1, %1, %10;\n\t subc.cc.u32 %2, %2, %11;\n\t"
      "subc.cc.u32 %3, %3, %12;\n\t subc.cc.u32 %4, %4, %13;\n\t subc.cc.u32 %5, %5, %14;\n\t"
      "subc.cc.u32 %6, %6, %15;\n\t subc.cc.u32 %7, %7, %16;\n\t subc.u32 %8, %8, 0;"
      : "+r"(pm[0]), "+r"(pm[1]), "+r"(pm[2]), "+r"(pm[3]), "+r"(pm[4]), "+r"(pm[5]), "+r"(pm[6]),
        "+r"(pm[7]), "+r"(pm8)
      : "r"(p0[0]), "r"(p0[1]), "r"(p0[2]), "r"(p0[3]), "r"(p0[4]), "r"(p0[5]), "r"(p0[6]), "r"(p0[7]));
  asm("sub.cc.u32 %0, %0, %9;\n\t subc.cc.u32 %1, %1, %10;\n\t subc.cc.u32 %2, %2, %11;\n\t"
      "subc.cc.u32 %3, %3, %12;\n\t subc.cc.u32 %4, %4, %13;\n\t subc.cc.u32 %5, %5, %14;\n\t"
      "subc.cc.u32 %6, %6, %15;\n\t subc.cc.u32 %7, %7, %16;\n\t subc.u32 %8, %8, 0;"
      : "+r"(pm[0]), "+r"(pm[1]), "+r"(pm[2]), "+r"(pm[3]), "+r"(pm[4]), "+r"(pm[5]), "+r"(pm[6]),
        "+r"(pm[7]), "+r"(pm8)
      : "r"(p2[0]), "r"(p2[1]), "r"(p2[2]), "r"(p2[3]), "r"(p2[4]), "r"(p2[5]), "r"(p2[6]), "r"(p2[7]));
  // T = P0 + 2^128 P1 + 2^256 P2
  uint32_t t[16];
  t[0] = p0[0]; t[1] = p0[1]; t[2] = p0[2]; t[3] = p0[3];
  asm("add.cc.u32 %0, %12, %24;\n\t"
      "addc.cc.u32 %1, %13, %25;\n\t"
      "addc.cc.u32 %2, %14, %26;\n\t"
      "addc.cc.u32 %3, %15, %27;\n\t"
      "addc.cc.u32 %4, %16, %28;\n\t"
      "addc.cc.u32 %5, %17, %29;\n\t"
      "addc.cc.u32 %6, %18, %30;\n\t"
      "addc.cc.u32 %7, %19, %31;\n\t"
      "addc.cc.u32 %8, %20, %32;\n\t"
      "addc.cc.u32 %9, %21, 0;\n\t"
      "addc.cc.u32 %10, %22, 0;\n\t"
      "addc.u32 %11, %23, 0;"
      : "=&r"(t[4]), "=&r"(t[5]), "=&r"(t[6]), "=&r"(t[7]), "=&r"(t[8]), "=&r"(t[9]), "=&r"(t[10]),
        "=&r"(t[11]), "=&r"(t[12]), "=&r"(t[13]), "=&r"(t[14]), "=&r"(t[15])
      : "r"(p0[4]), "r"(p0[5]), "r"(p0[6]), "r"(p0[7]), "r"(p2[0]), "r"(p2[1]), "r"(p2[2]), "r"(p2[3]),
        "r"(p2[4]), "r"(p2[5]), "r"(p2[6]), "r"(p2[7]), "r"(pm[0]), "r"(pm[1]), "r"(pm[2]), "r"(pm[3]),
        "r"(pm[4]), "r"(pm[5]), "r"(pm[6]), "r"(pm[7]), "r"(pm8));
  fqb<fq_bd_mul(A, B)> r;
  fq_redc16(r.l, t);
  return r;
}

// The multiplication every formula uses.  Karatsuba was built, is bit-exact, and LOSES on
// B200 (tools/ub_field.cu, profiles/r2_ub_field_karatsuba.txt): per multiplication a dependent
// chain takes the same time (8 611 against 8 690 G wide-multiply-equivalents / s), two
// interleaved chains 5 % longer, a 7M bucket addition 7.5 % longer (7 951 against 8 594) --
// the ~95 extra additions and the 15 more live registers cost the issue slots and the
// occupancy that the 16 saved wide multiplies free on the pipe.  So the row-interleaved form
// stays; -DD377_MUL_KARATSUBA builds the other one.
template <int A, int B>
D377_DI fqb<fq_bd_mul(A, B)> fq_mul(const fqb<A>& a, const fqb<B>& b) {
#ifdef D377_MUL_KARATSUBA
  return fq_mul_kara(a, b);
#else
  return fq_mul_cios(a, b);
#endif
}

// fiat.rs:1360 (fq_square): 28 off-diagonal products accumulated in an even-
// and an odd-aligned 16-limb accumulator (every carry-out lands on a limb no
// earlier row has touched, so no ripple), doubled, plus the 8 squares, then
// fq_redc16.  r <= q + x^2 / R.
template <int A>
D377_DI fqb<fq_bd_mul(A, A)> fq_sqr(const fqb<A>& x) {
  static_assert(fq_bd_mul(A, A) <= FQ_CAP, "fq_sqr result would overflow 8 limbs");
  const uint32_t a0 = x.l[0], a1 = x.l[1], a2 = x.l[2], a3 = x.l[3], a4 = x.l[4], a5 = x.l[5],
                 a6 = x.l[6], a7 = x.l[7];
  uint32_t e2, e3, e4, e5, e6, e7, e8, e9, e10, e11, e12, e13;   // position p
  uint32_t o0, o1, o2, o3, o4, o5, o6, o7, o8, o9, o10, o11, o12, o13;  // position p + 1
  // row 0: a0 * a1..a7
  asm("mul.lo.u32 %0, %8, %9;\n\t mul.hi.u32 %1, %8, %9;\n\t"
      "mul.lo.u32 %2, %8, %10;\n\t mul.hi.u32 %3, %8, %10;\n\t"
      "mul.lo.u32 %4, %8, %11;\n\t mul.hi.u32 %5, %8, %11;\n\t"
      "mul.lo.u32 %6, %8, %12;\n\t mul.hi.u32 %7, %8, %12;"
      : "=&r"(o0), "=&r"(o1), "=&r"(o2), "=&r"(o3), "=&r"(o4), "=&r"(o5), "=&r"(o6), "=&r"(o7)
      : "r"(a0), "r"(a1), "r"(a3), "r"(a5), "r"(a7));
  asm("mul.lo.u32 %0, %6, %7;\n\t mul.hi.u32 %1, %6, %7;\n\t"
      "mul.lo.u32 %2, %6, %8;\n\t mul.hi.u32 %3, %6, %8;\n\t"
      "mul.lo.u32 %4, %6, %9;\n\t mul.hi.u32 %5, %6, %9;"
      : "=&r"(e2), "=&r"(e3), "=&r"(e4), "=&r"(e5), "=&r"(e6), "=&r"(e7)
      : "r"(a0), "r"(a2), "r"(a4), "r"(a6));
  // row 1: a1 * a2..a7 -> positions 3..8
  asm("mad.lo.cc.u32 %0, %7, %8, %0;\n\t madc.hi.cc.u32 %1, %7, %8, %1;\n\t"
      "madc.lo.cc.u32 %2, %7, %9, %2;\n\t madc.hi.cc.u32 %3, %7, %9, %3;\n\t"
      "madc.lo.cc.u32 %4, %7, %10, %4;\n\t madc.hi.cc.u32 %5, %7, %10, %5;\n\t"
      "addc.u32 %6, 0, 0;"
      : "+r"(o2), "+r"(o3), "+r"(o4), "+r"(o5), "+r"(o6), "+r"(o7), "=r"(o8)
      : "r"(a1), "r"(a2), "r"(a4), "r"(a6));
  asm("mad.lo.cc.u32 %0, %6, %7, %0;\n\t madc.hi.cc.u32 %1, %6, %7, %1;\n\t"
      "madc.lo.cc.u32 %2, %6, %8, %2;\n\t madc.hi.cc.u32 %3, %6, %8, %3;\n\t"
      "madc.lo.cc.u32 %4, %6, %9, 0;\n\t madc.hi.u32 %5, %6, %9, 0;"
      : "+r"(e4), "+r"(e5), "+r"(e6), "+r"(e7), "=&r"(e8), "=&r"(e9)
      : "r"(a1), "r"(a3), "r"(a5), "r"(a7));
  // row 2: a2 * a3..a7 -> positions 5..9
  asm("mad.lo.cc.u32 %0, %6, %7, %0;\n\t madc.hi.cc.u32 %1, %6, %7, %1;\n\t"
      "madc.lo.cc.u32 %2, %6, %8, %2;\n\t madc.hi.cc.u32 %3, %6, %8, %3;\n\t"
      "madc.lo.cc.u32 %4, %6, %9, %4;\n\t madc.hi.u32 %5, %6, %9, 0;"
      : "+r"(o4), "+r"(o5), "+r"(o6), "+r"(o7), "+r"(o8), "=&r"(o9)
      : "r"(a2), "r"(a3), "r"(a5), "r"(a7));
  asm("mad.lo.cc.u32 %0, %5, %6, %0;\n\t madc.hi.cc.u32 %1, %5, %6, %1;\n\t"
      "madc.lo.cc.u32 %2, %5, %7, %2;\n\t madc.hi.cc.u32 %3, %5, %7, %3;\n\t"
      "addc.u32 %4, 0, 0;"
      : "+r"(e6), "+r"(e7), "+r"(e8), "+r"(e9), "=r"(e10)
      : "r"(a2), "r"(a4), "r"(a6));
  // row 3: a3 * a4..a7 -> positions 7..10
  asm("mad.lo.cc.u32 %0, %5, %6, %0;\n\t madc.hi.cc.u32 %1, %5, %6, %1;\n\t"
      "madc.lo.cc.u32 %2, %5, %7, %2;\n\t madc.hi.cc.u32 %3, %5, %7, %3;\n\t"
      "addc.u32 %4, 0, 0;"
      : "+r"(o6), "+r"(o7), "+r"(o8), "+r"(o9), "=r"(o10)
      : "r"(a3), "r"(a4), "r"(a6));
  asm("mad.lo.cc.u32 %0, %4, %5, %0;\n\t madc.hi.cc.u32 %1, %4, %5, %1;\n\t"
      "madc.lo.cc.u32 %2, %4, %6, %2;\n\t madc.hi.u32 %3, %4, %6, 0;"
      : "+r"(e8), "+r"(e9), "+r"(e10), "=&r"(e11)
      : "r"(a3), "r"(a5), "r"(a7));
  // row 4: a4 * a5..a7 -> positions 9..11
  asm("mad.lo.cc.u32 %0, %4, %5, %0;\n\t madc.hi.cc.u32 %1, %4, %5, %1;\n\t"
      "madc.lo.cc.u32 %2, %4, %6, %2;\n\t madc.hi.u32 %3, %4, %6, 0;"
      : "+r"(o8), "+r"(o9), "+r"(o10), "=&r"(o11)
      : "r"(a4), "r"(a5), "r"(a7));
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32 %2, 0, 0;"
      : "+r"(e10), "+r"(e11), "=r"(e12)
      : "r"(a4), "r"(a6));
  // row 5: a5 * a6, a7 -> positions 11, 12
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32 %2, 0, 0;"
      : "+r"(o10), "+r"(o11), "=r"(o12)
      : "r"(a5), "r"(a6));
  asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\t madc.hi.u32 %1, %2, %3, 0;"
      : "+r"(e12), "=&r"(e13)
      : "r"(a5), "r"(a7));
  // row 6: a6 * a7 -> position 13
  asm("mad.lo.cc.u32 %0, %2, %3, %0;\n\t madc.hi.u32 %1, %2, %3, 0;"
      : "+r"(o12), "=&r"(o13)
      : "r"(a6), "r"(a7));
  // S = E + O * 2^32 (limbs 1..15), then T = 2 S + sum a_i^2 2^(64 i)
  uint32_t s[16];
  s[0] = 0;
  s[1] = o0;
  asm("add.cc.u32 %0, %14, %26;\n\t"
      "addc.cc.u32 %1, %15, %27;\n\t"
      "addc.cc.u32 %2, %16, %28;\n\t"
      "addc.cc.u32 %3, %17, %29;\n\t"
      "addc.cc.u32 %4, %18, %30;\n\t"
      "addc.cc.u32 %5, %19, %31;\n\t"
      "addc.cc.u32 %6, %20, %32;\n\t"
      "addc.cc.u32 %7, %21, %33;\n\t"
      "addc.cc.u32 %8, %22, %34;\n\t"
      "addc.cc.u32 %9, %23, %35;\n\t"
      "addc.cc.u32 %10, %24, %36;\n\t"
      "addc.cc.u32 %11, %25, %37;\n\t"
      "addc.cc.u32 %12, %38, 0;\n\t"
      "addc.u32 %13, 0, 0;"
      : "=&r"(s[2]), "=&r"(s[3]), "=&r"(s[4]), "=&r"(s[5]), "=&r"(s[6]), "=&r"(s[7]), "=&r"(s[8]),
        "=&r"(s[9]), "=&r"(s[10]), "=&r"(s[11]), "=&r"(s[12]), "=&r"(s[13]), "=&r"(s[14]),
        "=&r"(s[15])
      : "r"(e2), "r"(e3), "r"(e4), "r"(e5), "r"(e6), "r"(e7), "r"(e8), "r"(e9), "r"(e10), "r"(e11),
        "r"(e12), "r"(e13), "r"(o1), "r"(o2), "r"(o3), "r"(o4), "r"(o5), "r"(o6), "r"(o7), "r"(o8),
        "r"(o9), "r"(o10), "r"(o11), "r"(o12), "r"(o13));
  uint32_t t[16];
  t[0] = 0;
#pragma unroll
  for (int p = 15; p >= 1; p--) t[p] = __funnelshift_l(s[p - 1], s[p], 1);
  asm("mad.lo.cc.u32 %0, %16, %16, %0;\n\t madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
      "madc.lo.cc.u32 %2, %17, %17, %2;\n\t madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
      "madc.lo.cc.u32 %4, %18, %18, %4;\n\t madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
      "madc.lo.cc.u32 %6, %19, %19, %6;\n\t madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
      "madc.lo.cc.u32 %8, %20, %20, %8;\n\t madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
      "madc.lo.cc.u32 %10, %21, %21, %10;\n\t madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
      "madc.lo.cc.u32 %12, %22, %22, %12;\n\t madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
      "madc.lo.cc.u32 %14, %23, %23, %14;\n\t madc.hi.u32 %15, %23, %23, %15;"
      : "+r"(t[0]), "+r"(t[1]), "+r"(t[2]), "+r"(t[3]), "+r"(t[4]), "+r"(t[5]), "+r"(t[6]),
        "+r"(t[7]), "+r"(t[8]), "+r"(t[9]), "+r"(t[10]), "+r"(t[11]), "+r"(t[12]), "+r"(t[13]),
        "+r"(t[14]), "+r"(t[15])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7));
  fqb<fq_bd_mul(A, A)> r;
  fq_redc16(r.l, t);
  return r;
}

// fiat.rs:2800 (fq_from_montgomery): a / R mod q, i.e. the canonical value (< q) of
// any 256-bit a: the eight rows leave (a + M q) / R <= q, one conditional subtraction.
template <int A>
D377_DI fq_r fq_from_mont(const fqb<A>& a) {
  const fq_mod_t q = fq_mod();
  uint32_t ev[8], od[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    ev[i] = a.l[i];
    od[i] = 0;
  }
  uint32_t cy = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if (i == 0) fq_redc_row<false>(ev, od, q, 0u); else fq_redc_row<true>(ev, od, q, cy);
    if (i < 7) cy = fq_shift_only(ev, od);
  }
  fqb<2000> t;
  fq_mont_collapse(t.l, ev, od);
  return fq_csub<1>(t);
}

// fiat.rs:3584 (fq_to_montgomery) for ANY 256-bit string a (canonical or not):
// R2 is the row multiplicand (T < R2 + q between rows whatever a is) and the result is
// < q + R2 * a / R < q + R2 < 2q.  This is Fq::from_le_bytes_mod_order on 32 bytes
// (fields/fq.rs:90-102) as well.
D377_DI fq_t fq_to_mont(const fq_raw_t& a) {
  // the generic bound (1 + 13.712 * 0.07293 = 2.00002) is a hair above 2q only because
  // 0.07293 is rounded up; R2 < q makes the true bound < 2q.
  return fq_assume<2000>(fq_mul(fq_const(FQ_R2), a));
}

// sign.rs:19-23: parity of the canonical value.
template <int A>
D377_DI bool fq_is_negative(const fqb<A>& a) { return fq_from_mont(a).l[0] & 1u; }

// sign.rs:10-16
template <int A>
D377_DI fqb<fq_bd_max(A, 1000 * fq_bd_k(A) + 1)> fq_abs(const fqb<A>& a) {
  bool neg = fq_is_negative(a);
  return fq_select(neg, fq_neg(a), a);
}

// true iff the 8 raw limbs are < q (canonical), fq.rs:108-115.
D377_DI bool fq_raw_is_canonical(const fq_raw_t& a) {
  uint32_t d[8];
  return fq_sub_kq_raw<1>(d, a.l) != 0;
}

// ---- 32-byte vectorised global I/O ---------------------------------------
// Memory holds storage-class values (< 2q): wire inputs are canonical Montgomery (< q) by
// the ABI contract, internal workspaces are written by fq_store below.
// sm_100 moves 32 bytes per thread with ONE instruction (LDG.E.256 / STG.E.256); every Fq
// on this ABI is 32-byte aligned (buffers from cudaMalloc / torch are at least 256-byte
// aligned and all record sizes -- 32, 64, 96, 128 -- are multiples of 32).  The stores and the
// streaming loads of the MSM use the 256-bit forms below.  The general-purpose load stays at
// two 128-bit loads: with `ld.global.nc.v8.b32` here, ptxas 12.9 crashes (SIGSEGV) on the
// codec and scalar-multiplication translation units.
D377_DI fq_t fq_load(const void* p) {
  const uint4* v = reinterpret_cast<const uint4*>(p);
  uint4 lo = __ldg(v), hi = __ldg(v + 1);
  fq_t r;
  r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
  r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
  return r;
}

// data this kernel itself wrote earlier (the non-coherent path above must not see it)
D377_DI fq_t fq_load_rw(const void* p) {
  fq_t r;
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]),
                 "=r"(r.l[6]), "=r"(r.l[7])
               : "l"(p)
               : "memory");
  return r;
}

// Same for data that is read once (the gathered bucket operands of an MSM): marked
// evict-first in L2 so that a concurrent kernel's working set survives the stream.  The
// 256-bit form carries the eviction priority in the instruction (LDG.E.EFL2.256): no policy
// register to keep alive across the loop.
D377_DI fq_t fq_load_stream(const void* p) {
  fq_t r;
  asm volatile("ld.global.nc.L2::evict_first.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]),
                 "=r"(r.l[6]), "=r"(r.l[7])
               : "l"(p));
  return r;
}

// 32 arbitrary bytes (encodings, scalars, Fq inputs before reduction)
D377_DI fq_raw_t fq_load_raw(const void* p) { return fq_assume<FQ_RAWB>(fq_load(p)); }

// Montgomery limbs handed in by a CALLER (the ABI contract says canonical, < q, but the
// buffer is untrusted): any 256-bit string is brought below 2q here, so that the static
// bounds of everything downstream hold whatever the bytes are (a value >= 2q would
// otherwise overflow the 8-limb accumulators silently, and a multiple of q could spin
// fq_inv_vartime).  The string is read as an integer mod q, i.e. a non-canonical
// representative of a field element is accepted and means that element.  Canonical input
// (the only kind the reference can produce) takes the first branch: one compare on the top
// limb (2q = 0x2556cabd...).
D377_DI fq_t fq_load_wire(const void* p) {
  const fq_raw_t r = fq_load_raw(p);
  if (r.l[7] < 0x2556cabcu) return fq_assume<2000>(r);
  return fq_csub<2>(fq_csub<4>(fq_csub<8>(r)));
}

// internal workspaces: lazily reduced
D377_DI void fq_store(void* p, const fq_t& a) {
  asm volatile("st.global.v8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};"
               :
               : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]),
                 "r"(a.l[7]), "l"(p)
               : "memory");
}

// ABI outputs: canonical Montgomery form, byte-identical to the reference's Fq
template <int A>
D377_DI void fq_store_canon(void* p, const fqb<A>& a) { fq_store(p, fq_reduce(a)); }

// ---- Fq::from_le_bytes_mod_order for any input length (fields/fq.rs:90-102) ----------
// up to 32 bytes at any alignment, zero-padded to a 256-bit little-endian value
D377_DI fq_raw_t fq_load_bytes(const uint8_t* p, size_t len) {
  fq_raw_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = 0;
#pragma unroll 1
  for (size_t i = 0; i < len; i++) r.l[i >> 2] |= (uint32_t)p[i] << (8 * (i & 3));
  return r;
}

// The reference folds the 32-byte chunks from the most significant one down with
// acc = acc * 2^256 + chunk (FIELD_SIZE_POWER_OF_TWO = 2^256 mod q).  In Montgomery form a
// multiplication by 2^256 = R is the Montgomery product with R^2, which is also what brings
// a raw chunk into Montgomery form: one fq_mul per chunk and per step.
D377_DI fq_t fq_from_le_bytes_wide(const uint8_t* p, size_t width) {
  const size_t nchunks = (width + 31) / 32;
  fq_t acc = fq_zero();
#pragma unroll 1
  for (size_t c = nchunks; c-- > 0;) {
    const size_t off = 32 * c, len = width - off < 32 ? width - off : 32;
    const bool vec = len == 32 && ((reinterpret_cast<uintptr_t>(p + off) & 15) == 0);
    const fq_raw_t x = vec ? fq_load_raw(p + off) : fq_load_bytes(p + off, len);
    const fq_t xm = fq_to_mont(x);
    if (c + 1 == nchunks) acc = xm;
    else acc = fq_fold(fq_add(fq_mul(fq_const(FQ_R2), acc), xm));
  }
  return acc;
}
