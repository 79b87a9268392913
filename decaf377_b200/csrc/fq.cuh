// Fq = GF(q), q = decaf377 base field (BLS12-377 scalar field), 253 bits.
//
// Replaces the reference's L1/L2 field layer for the batch path:
//   src/fields/fq/u32/fiat.rs:162 (fq_mul), :1360 (fq_square), :2555 (fq_add),
//   :2646 (fq_sub), :2725 (fq_opp), :2800 (fq_from_montgomery),
//   :3584 (fq_to_montgomery); src/sign.rs:3-23.
//
// Representation: one element per thread, 8 x 32-bit limbs in registers,
// Montgomery form with R = 2^256 (byte-identical to both reference backends,
// fq/u32/wrapper.rs:93-104), always fully reduced (< q).
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

struct fq_t {
  uint32_t l[8];
};

// q limbs as literals so that ptxas can fold them into immediates.
#define Q0 0x00000001u
#define Q1 0x0a118000u
#define Q2 0xd0000001u
#define Q3 0x59aa76feu
#define Q4 0x5c37b001u
#define Q5 0x60b44d1eu
#define Q6 0x9a2ca556u
#define Q7 0x12ab655eu

D377_DI fq_t fq_const(const uint32_t (&c)[8]) {
  fq_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = c[i];
  return r;
}

D377_DI fq_t fq_zero() {
  fq_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = 0;
  return r;
}

D377_DI fq_t fq_one() { return fq_const(FQ_ONE); }

D377_DI bool fq_is_zero(const fq_t& a) {
  uint32_t o = a.l[0];
#pragma unroll
  for (int i = 1; i < 8; i++) o |= a.l[i];
  return o == 0;
}

D377_DI bool fq_eq(const fq_t& a, const fq_t& b) {
  uint32_t o = a.l[0] ^ b.l[0];
#pragma unroll
  for (int i = 1; i < 8; i++) o |= a.l[i] ^ b.l[i];
  return o == 0;
}

// r = c ? a : b
D377_DI fq_t fq_select(bool c, const fq_t& a, const fq_t& b) {
  fq_t r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = c ? a.l[i] : b.l[i];
  return r;
}

// t - q, returns borrow (1 if t < q).
D377_DI uint32_t fq_sub_mod_raw(uint32_t (&d)[8], const uint32_t (&t)[8]) {
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
        "n"(Q0), "n"(Q1), "n"(Q2), "n"(Q3), "n"(Q4), "n"(Q5), "n"(Q6), "n"(Q7));
  return bw;  // 0xffffffff when borrow
}

// Final conditional subtraction: t in [0, 2q) -> [0, q).
D377_DI void fq_reduce_once(fq_t& r, const uint32_t (&t)[8]) {
  uint32_t d[8];
  uint32_t bw = fq_sub_mod_raw(d, t);
#pragma unroll
  for (int i = 0; i < 8; i++) r.l[i] = bw ? t[i] : d[i];
}

// fiat.rs:2555 (fq_add)
D377_DI fq_t fq_add(const fq_t& a, const fq_t& b) {
  uint32_t t[8];
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]),
        "=r"(t[7])
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]),
        "r"(a.l[6]), "r"(a.l[7]), "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]),
        "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
  fq_t r;
  fq_reduce_once(r, t);  // a + b < 2q < 2^254: no carry out of limb 7
  return r;
}

// fiat.rs:2646 (fq_sub)
D377_DI fq_t fq_sub(const fq_t& a, const fq_t& b) {
  uint32_t t[8], bw;
  asm("sub.cc.u32 %0, %9, %17;\n\t"
      "subc.cc.u32 %1, %10, %18;\n\t"
      "subc.cc.u32 %2, %11, %19;\n\t"
      "subc.cc.u32 %3, %12, %20;\n\t"
      "subc.cc.u32 %4, %13, %21;\n\t"
      "subc.cc.u32 %5, %14, %22;\n\t"
      "subc.cc.u32 %6, %15, %23;\n\t"
      "subc.cc.u32 %7, %16, %24;\n\t"
      "subc.u32 %8, 0, 0;"
      : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]),
        "=r"(t[7]), "=r"(bw)
      : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]),
        "r"(a.l[6]), "r"(a.l[7]), "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]),
        "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
  // add back q & mask
  fq_t r;
  asm("add.cc.u32 %0, %8, %16;\n\t"
      "addc.cc.u32 %1, %9, %17;\n\t"
      "addc.cc.u32 %2, %10, %18;\n\t"
      "addc.cc.u32 %3, %11, %19;\n\t"
      "addc.cc.u32 %4, %12, %20;\n\t"
      "addc.cc.u32 %5, %13, %21;\n\t"
      "addc.cc.u32 %6, %14, %22;\n\t"
      "addc.u32 %7, %15, %23;"
      : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]),
        "=r"(r.l[6]), "=r"(r.l[7])
      : "r"(t[0]), "r"(t[1]), "r"(t[2]), "r"(t[3]), "r"(t[4]), "r"(t[5]), "r"(t[6]), "r"(t[7]),
        "r"(bw & Q0), "r"(bw & Q1), "r"(bw & Q2), "r"(bw & Q3), "r"(bw & Q4), "r"(bw & Q5),
        "r"(bw & Q6), "r"(bw & Q7));
  return r;
}

// fiat.rs:2725 (fq_opp)
D377_DI fq_t fq_neg(const fq_t& a) { return fq_sub(fq_zero(), a); }

D377_DI fq_t fq_dbl(const fq_t& a) { return fq_add(a, a); }

// ---------------------------------------------------------------------------
// Montgomery product core.
//
// Accumulator invariant between rows: T = E + O * 2^32 where E = ev[0..8)
// sits at limb positions 0..7 and O = od[0..8) at positions 1..8.
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

// acc(4 aligned 64-bit lanes) += {x0,x2,x4,x6} * y, returns carry-out.
#define D377_CMAD4(acc, x0, x2, x4, x6, y, cout)                                             \
  asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"                                                   \
      "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"                                                  \
      "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"                                                 \
      "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"                                                 \
      "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"                                                 \
      "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"                                                 \
      "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"                                                 \
      "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"                                                 \
      "addc.u32 %8, 0, 0;"                                                                   \
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]),  \
        "+r"(acc[6]), "+r"(acc[7]), "=r"(cout)                                               \
      : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(y))

// One Montgomery reduction row on (ev, od): adds m*q with m = -ev[0] (q = 1 mod 2^32
// so -q^-1 = 0xffffffff, fiat.rs:223) making ev[0] zero.  `cy` (0/1) is a pending
// carry of weight 2^32 that enters the odd chain.
template <bool kCy>
D377_DI void fq_redc_row(uint32_t (&ev)[8], uint32_t (&od)[8], const fq_mod_t& q, uint32_t cy) {
  uint32_t m = q.z - ev[0];
  uint32_t c;
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
      "addc.u32 %8, 0, 0;"
      : "+r"(ev[0]), "+r"(ev[1]), "+r"(ev[2]), "+r"(ev[3]), "+r"(ev[4]), "+r"(ev[5]),
        "+r"(ev[6]), "+r"(ev[7]), "=r"(c)
      : "r"(q.q2), "r"(q.q4), "r"(q.q6), "r"(m));
  od[7] += c;
}

// Shift (ev, od) right by one limb (ev[0] == 0 on entry) and add a*b_i.
// After the shift the old odd accumulator is aligned to even positions; the
// old even accumulator moves down one 64-bit lane and becomes the odd one,
// except for its limb 1, which lands on position 0 and is folded into od[0]
// -- the carry of that fold has weight 2^32, i.e. it is exactly the carry-in
// of the new odd chain.
D377_DI void fq_mul_row_shift(uint32_t (&ev)[8], uint32_t (&od)[8], const fq_t& a, uint32_t bi) {
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
        "r"(a.l[1]), "r"(a.l[3]), "r"(a.l[5]), "r"(a.l[7]), "r"(bi));
  uint32_t c;
  D377_CMAD4(od, a.l[0], a.l[2], a.l[4], a.l[6], bi, c);
  nod[7] += c;
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

// Collapse T = E + O*2^32 after the last reduction row (one more limb shift),
// optionally add an 8-limb `hi` (the upper half of a 512-bit product), and do the
// final conditional subtraction.  Result < 2q before it in both uses.
D377_DI fq_t fq_mont_finish(uint32_t (&ev)[8], uint32_t (&od)[8]) {
  uint32_t t[8];
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
  fq_t r;
  fq_reduce_once(r, t);
  return r;
}

D377_DI fq_t fq_mont_finish_hi(uint32_t (&ev)[8], uint32_t (&od)[8], const uint32_t (&hi)[8]) {
  uint32_t t[8];
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
  fq_t r;
  fq_reduce_once(r, t);
  return r;
}

// fiat.rs:162 (fq_mul): r = a * b / R mod q
D377_DI fq_t fq_mul(const fq_t& a, const fq_t& b) {
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
    fq_mul_row_shift(ev, od, a, b.l[i]);
    fq_redc_row<false>(ev, od, q, 0u);
  }
  return fq_mont_finish(ev, od);
}

// Montgomery reduction of a 512-bit value t[0..16): t / R mod q.  Eight reduction
// rows run on the low half only (the result of those is <= q); the high half is
// added at the end (t < q^2 makes the total < 2q).
D377_DI fq_t fq_redc16(const uint32_t (&t)[16]) {
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
  return fq_mont_finish_hi(ev, od, hi);
}

// fiat.rs:1360 (fq_square): 28 off-diagonal products accumulated in an even-
// and an odd-aligned 16-limb accumulator (every carry-out lands on a limb no
// earlier row has touched, so no ripple), doubled, plus the 8 squares, then
// fq_redc16.
D377_DI fq_t fq_sqr(const fq_t& x) {
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
  return fq_redc16(t);
}

// fiat.rs:2800 (fq_from_montgomery): a / R mod q, i.e. the canonical value.
D377_DI fq_t fq_from_mont(const fq_t& a) {
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
  return fq_mont_finish(ev, od);
}

// fiat.rs:3584 (fq_to_montgomery)
D377_DI fq_t fq_to_mont(const fq_t& a) { return fq_mul(a, fq_const(FQ_R2)); }

// sign.rs:19-23: parity of the canonical value.
D377_DI bool fq_is_negative(const fq_t& a) { return fq_from_mont(a).l[0] & 1u; }

// sign.rs:10-16
D377_DI fq_t fq_abs(const fq_t& a) {
  bool neg = fq_is_negative(a);
  fq_t n = fq_neg(a);
  return fq_select(neg, n, a);
}

// true iff the 8 raw limbs are < q (canonical), fq.rs:108-115.
D377_DI bool fq_raw_is_canonical(const fq_t& a) {
  uint32_t d[8];
  return fq_sub_mod_raw(d, a.l) != 0;
}

// ---- 32-byte vectorised global I/O ---------------------------------------
D377_DI fq_t fq_load(const void* p) {
  const uint4* v = reinterpret_cast<const uint4*>(p);
  uint4 lo = __ldg(v), hi = __ldg(v + 1);
  fq_t r;
  r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
  r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
  return r;
}

D377_DI void fq_store(void* p, const fq_t& a) {
  uint4* v = reinterpret_cast<uint4*>(p);
  v[0] = make_uint4(a.l[0], a.l[1], a.l[2], a.l[3]);
  v[1] = make_uint4(a.l[4], a.l[5], a.l[6], a.l[7]);
}
