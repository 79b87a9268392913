/* tests/abi_smoke.c -- the C ABI of include/decaf377_b200.h driven from a plain C host.
 *
 * Compiled with gcc and linked against decaf377_b200/libdecaf377_b200.so by
 * tests/test_gpu_round2.py::test_c_host_program_drives_the_abi.  It makes the same calls, in
 * the same order, as the Rust shim of rust/src/gpu.rs (init -> batch codec -> Elligator ->
 * scalar mul -> fixed base -> group ops -> MSM in all its forms -> shutdown) and checks them
 * against vectors the reference itself holds: the encodings of i*G, i = 0..15
 * (reference tests/encoding.rs:61-78), identity = 00..00 and generator = 08 00..00
 * (tests/encoding.rs:19-52), plus algebraic identities that need no second implementation.
 * No oracle, no Python: this is the boundary as a non-Python host sees it. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "decaf377_b200.h"

static int failures = 0;
#define CHECK(cond)                                                              \
  do {                                                                           \
    if (!(cond)) {                                                               \
      printf("FAIL %s:%d: %s  [%s]\n", __FILE__, __LINE__, #cond, d377_last_error()); \
      failures++;                                                                \
    }                                                                            \
  } while (0)
#define OK(call) CHECK((call) == D377_OK)

static const uint8_t KAT[16][32] = {
  {0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00},
  {0x08, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00},
  {0xb2, 0xec, 0xf9, 0xb9, 0x08, 0x2d, 0x63, 0x06, 0x53, 0x8b, 0xe7, 0x3b, 0x0d, 0x6e, 0xe7, 0x41, 0x14, 0x1f, 0x32, 0x22, 0x15, 0x2d, 0xa7, 0x86, 0x85, 0xd6, 0x59, 0x6e, 0xfc, 0x8c, 0x15, 0x06},
  {0x2e, 0xbd, 0x42, 0xdd, 0x3a, 0x23, 0x07, 0x08, 0x3c, 0x83, 0x4e, 0x79, 0xfb, 0x9e, 0x78, 0x7e, 0x35, 0x2d, 0xd3, 0x3e, 0x0d, 0x71, 0x9f, 0x86, 0xae, 0x4a, 0xdb, 0x02, 0xfe, 0x38, 0x24, 0x09},
  {0x6a, 0xcd, 0x32, 0x7d, 0x70, 0xf9, 0x58, 0x8f, 0xac, 0x37, 0x3d, 0x16, 0x5f, 0x4d, 0x9d, 0x53, 0x00, 0x51, 0x02, 0x74, 0xdf, 0xfd, 0xfd, 0xf2, 0xbf, 0x09, 0x55, 0xac, 0xd7, 0x8d, 0xa5, 0x0d},
  {0x46, 0x0f, 0x91, 0x3e, 0x51, 0x64, 0x41, 0xc2, 0x86, 0xd9, 0x5d, 0xd3, 0x0b, 0x0a, 0x2d, 0x2b, 0xf1, 0x42, 0x64, 0xf3, 0x25, 0x52, 0x8b, 0x06, 0x45, 0x5d, 0x7c, 0xb9, 0x3b, 0xa1, 0x3a, 0x0b},
  {0xec, 0x87, 0x98, 0xbc, 0xbb, 0x3b, 0xf2, 0x93, 0x29, 0x54, 0x9d, 0x76, 0x9f, 0x89, 0xcf, 0x79, 0x93, 0xe1, 0x5e, 0x2c, 0x68, 0xec, 0x7a, 0xa2, 0xa9, 0x56, 0xed, 0xf5, 0xec, 0x62, 0xae, 0x07},
  {0x48, 0xb0, 0x1e, 0x51, 0x3d, 0xd3, 0x7d, 0x94, 0xc3, 0xb4, 0x89, 0x40, 0xdc, 0x13, 0x3b, 0x92, 0xcc, 0xba, 0x7f, 0x54, 0x6e, 0x99, 0xd3, 0xfc, 0x2e, 0x60, 0x2d, 0x28, 0x4f, 0x60, 0x9f, 0x00},
  {0xa4, 0xe8, 0x5d, 0xdd, 0xd1, 0x9c, 0x80, 0xec, 0xf5, 0xef, 0x10, 0xb9, 0xd2, 0x7b, 0x66, 0x26, 0xac, 0x1a, 0x4f, 0x90, 0xbd, 0x10, 0xd2, 0x63, 0xc7, 0x17, 0xec, 0xce, 0x4d, 0xa6, 0x57, 0x0a},
  {0x1a, 0x8f, 0xea, 0x8c, 0xbf, 0xbc, 0x91, 0x23, 0x6d, 0x8c, 0x79, 0x24, 0xe3, 0xe7, 0xe6, 0x17, 0xf9, 0xdd, 0x54, 0x4b, 0x71, 0x0e, 0xe8, 0x38, 0x27, 0x73, 0x7f, 0xe8, 0xdc, 0x63, 0xae, 0x00},
  {0x0a, 0x0f, 0x86, 0xea, 0xac, 0x0c, 0x1a, 0xf3, 0x0e, 0xb1, 0x38, 0x46, 0x7c, 0x49, 0x38, 0x1e, 0xdb, 0x28, 0x08, 0x90, 0x4c, 0x81, 0xa4, 0xb8, 0x1d, 0x2b, 0x02, 0xa2, 0xd7, 0x81, 0x60, 0x06},
  {0x58, 0x81, 0x25, 0xa8, 0xf4, 0xe2, 0xba, 0xb8, 0xd1, 0x6a, 0xff, 0xc4, 0xca, 0x60, 0xc5, 0xf6, 0x4b, 0x50, 0xd3, 0x8d, 0x2b, 0xb0, 0x53, 0x14, 0x80, 0x21, 0x63, 0x1f, 0x72, 0xe9, 0x9b, 0x06},
  {0xf4, 0x3f, 0x4c, 0xef, 0xbe, 0x73, 0x26, 0xea, 0xab, 0x15, 0x84, 0x72, 0x2b, 0x1b, 0x48, 0x60, 0xde, 0x55, 0x4b, 0x23, 0xa1, 0x44, 0x90, 0xa0, 0x3f, 0x3f, 0xd6, 0x3a, 0x08, 0x9a, 0xdd, 0x0b},
  {0x76, 0xc7, 0x39, 0xa3, 0x3f, 0xfd, 0x15, 0xcf, 0x65, 0x54, 0xa8, 0xe7, 0x05, 0xdc, 0x57, 0x3f, 0x26, 0x49, 0x0b, 0x64, 0xde, 0x0c, 0x5b, 0xd4, 0xe4, 0xac, 0x75, 0xed, 0x5a, 0xf8, 0xe6, 0x0b},
  {0x20, 0x01, 0x36, 0x95, 0x2d, 0x18, 0xd3, 0xf6, 0xc7, 0x03, 0x47, 0x03, 0x2b, 0xa3, 0xfe, 0xf4, 0xf6, 0x0c, 0x24, 0x0d, 0x70, 0x6b, 0xe2, 0x95, 0x0b, 0x4f, 0x42, 0xf1, 0xa7, 0x08, 0x77, 0x05},
  {0xbc, 0xb0, 0xf9, 0x22, 0xdf, 0x1c, 0x7a, 0xa9, 0x57, 0x93, 0x94, 0x02, 0x01, 0x87, 0xa2, 0xe1, 0x9e, 0x2d, 0x80, 0x73, 0x45, 0x2c, 0x6a, 0xb9, 0xb0, 0xc4, 0xb0, 0x52, 0xaa, 0x50, 0xf5, 0x05},
};

enum { N = 16, M = 4096 };

int main(void) {
  OK(d377_init(0));
  CHECK(d377_get_device() == 0);

  /* ---- Encoding::vartime_decompress / Element::vartime_compress ---- */
  static uint8_t el[N * 128], enc[N * 32], ok[N];
  OK(d377_batch_decompress(&KAT[0][0], N, el, ok));
  for (int i = 0; i < N; i++) CHECK(ok[i] == 1);
  OK(d377_batch_compress(el, N, enc));
  CHECK(memcmp(enc, KAT, sizeof KAT) == 0);
  {
    /* AffinePoint wire form of the same round trip (ark_curve/serialize.rs:8-46) */
    static uint8_t xy[N * 64], enc2[N * 32], ok2[N];
    OK(d377_batch_decompress_fmt(&KAT[0][0], N, D377_PT_AFFINE, xy, ok2));
    for (int i = 0; i < N; i++) {
      CHECK(ok2[i] == 1);
      CHECK(memcmp(xy + 64 * i, el + 128 * i, 64) == 0);
    }
    OK(d377_batch_compress_fmt(xy, D377_PT_AFFINE, N, enc2));
    CHECK(memcmp(enc2, KAT, sizeof KAT) == 0);
    CHECK(d377_batch_compress_fmt(xy, D377_PT_ENCODING, N, enc2) == D377_ERR_INVALID_ARG);
  }
  {
    uint8_t bad[2 * 32] = {0}, out[2 * 128], okb[2];
    bad[0] = 1;          /* s = 1: invalid (tests/encoding.rs:28-52) */
    bad[32 + 31] = 0x80; /* top bit set */
    OK(d377_batch_decompress(bad, 2, out, okb));
    CHECK(okb[0] == 0 && okb[1] == 0);
  }
  {
    uint8_t oc[N];
    OK(d377_batch_on_curve(el, N, 1, oc));
    for (int i = 0; i < N; i++) CHECK(oc[i] == 1);
  }

  /* ---- GENERATOR * s with the tables: s = 0..15 gives the KATs ---- */
  static uint8_t sc[N * 32], fb[N * 32], fbel[N * 128], eq[N];
  memset(sc, 0, sizeof sc);
  for (int i = 0; i < N; i++) sc[32 * i] = (uint8_t)i;
  OK(d377_fixed_base_mul(sc, N, fb, D377_OUT_ENCODING));
  CHECK(memcmp(fb, KAT, sizeof KAT) == 0);
  OK(d377_fixed_base_mul(sc, N, fbel, D377_OUT_ELEMENT));
  OK(d377_batch_element_eq(fbel, el, N, eq));
  for (int i = 0; i < N; i++) CHECK(eq[i] == 1);

  /* ---- &Element * &Fr: i * G from the generator's encoding ---- */
  static uint8_t gens[N * 32], sm[N * 32];
  for (int i = 0; i < N; i++) memcpy(gens + 32 * i, KAT[1], 32);
  OK(d377_batch_scalar_mul(gens, D377_PT_ENCODING, sc, N, sm, D377_OUT_ENCODING, ok));
  CHECK(memcmp(sm, KAT, sizeof KAT) == 0);

  /* ---- group operations: (a + b) - b == a, -(-a) == a, 2a == a + a ---- */
  static uint8_t t1[N * 128], t2[N * 128], rev[N * 128];
  for (int i = 0; i < N; i++) memcpy(rev + 128 * i, el + 128 * (N - 1 - i), 128);
  OK(d377_batch_add(el, rev, N, t1));
  OK(d377_batch_compress(t1, N, enc));
  for (int i = 0; i < N; i++) CHECK(memcmp(enc + 32 * i, KAT[15], 32) == 0); /* i + (15 - i) */
  OK(d377_batch_sub(t1, rev, N, t2));
  OK(d377_batch_element_eq(t2, el, N, eq));
  for (int i = 0; i < N; i++) CHECK(eq[i] == 1);
  OK(d377_batch_neg(el, N, t1));
  OK(d377_batch_neg(t1, N, t2));
  CHECK(memcmp(t2, el, sizeof el) == 0);
  OK(d377_batch_double(el, 8, t1));
  OK(d377_batch_compress(t1, 8, enc));
  for (int i = 0; i < 8; i++) CHECK(memcmp(enc + 32 * i, KAT[2 * i], 32) == 0);

  /* ---- Elligator: the fused encoding output equals compress of the Element output;
   *      64-byte inputs with a zero upper half equal the 32-byte inputs ---- */
  static uint8_t r32[M * 32], r64[M * 64], e1[M * 32], e2[M * 32], pts[M * 128];
  uint32_t x = 0x377u;
  for (size_t i = 0; i < sizeof r32; i++) {
    x = x * 1664525u + 1013904223u;
    r32[i] = (uint8_t)(x >> 24);
  }
  memset(r64, 0, sizeof r64);
  for (int i = 0; i < M; i++) memcpy(r64 + 64 * i, r32 + 32 * i, 32);
  OK(d377_batch_encode_to_curve(r32, M, pts, D377_OUT_ELEMENT));
  OK(d377_batch_compress(pts, M, e1));
  OK(d377_batch_encode_to_curve(r32, M, e2, D377_OUT_ENCODING));
  CHECK(memcmp(e1, e2, sizeof e1) == 0);
  OK(d377_batch_encode_to_curve_wide(r64, 64, M, e2, D377_OUT_ENCODING));
  CHECK(memcmp(e1, e2, sizeof e1) == 0);
  OK(d377_batch_hash_to_curve(r32, r32 + 32 * (M / 2), M / 2, e1, D377_OUT_ENCODING));
  OK(d377_batch_hash_to_curve(r32, r32 + 32 * (M / 2), M / 2, pts, D377_OUT_ELEMENT));
  OK(d377_batch_compress(pts, M / 2, e2));
  CHECK(memcmp(e1, e2, 32 * (M / 2)) == 0);

  /* ---- Element::vartime_multiscalar_mul ----
   * points i*G (i = 0..15), scalars all one: sum = 120 * G = fixed_base(120) */
  uint8_t ones[N * 32] = {0}, s120[32] = {120}, want[32], got_el[128], got[32];
  for (int i = 0; i < N; i++) ones[32 * i] = 1;
  OK(d377_fixed_base_mul(s120, 1, want, D377_OUT_ENCODING));
  OK(d377_msm(ones, el, D377_PT_ELEMENT, N, got_el, got));
  CHECK(memcmp(got, want, 32) == 0);
  OK(d377_msm(ones, &KAT[0][0], D377_PT_ENCODING, N, NULL, got));
  CHECK(memcmp(got, want, 32) == 0);
  /* scalars i, points G: sum_i i * G = 120 * G as well */
  static uint8_t gel[N * 128];
  for (int i = 0; i < N; i++) memcpy(gel + 128 * i, el + 128, 128);
  OK(d377_msm(sc, gel, D377_PT_ELEMENT, N, NULL, got));
  CHECK(memcmp(got, want, 32) == 0);
  /* empty MSM is the identity (Element::default()) */
  OK(d377_msm(NULL, NULL, D377_PT_ELEMENT, 0, NULL, got));
  CHECK(memcmp(got, KAT[0], 32) == 0);
  /* a scalar >= r is refused (Fr::from_bytes_checked) */
  {
    uint8_t big[N * 32];
    memcpy(big, sc, sizeof big);
    memset(big + 32 * 3, 0xff, 32);
    CHECK(d377_msm(big, gel, D377_PT_ELEMENT, N, NULL, got) == D377_ERR_SCALAR_RANGE);
  }
  /* AffinePoint bases (batch_convert_to_mul_base) and prepared bases */
  static uint8_t aff[N * 64];
  uint8_t* bases = NULL;
  OK(d377_batch_normalize(gel, N, aff));
  OK(d377_msm(sc, aff, D377_PT_AFFINE, N, NULL, got));
  CHECK(memcmp(got, want, 32) == 0);
  OK(d377_msm_bases_create(gel, D377_PT_ELEMENT, N, &bases));
  OK(d377_msm(sc, bases, D377_PT_BASES, N, NULL, got));
  CHECK(memcmp(got, want, 32) == 0);
  CHECK(d377_msm(sc, bases + 128, D377_PT_BASES, 4, NULL, got) == D377_ERR_INVALID_ARG);
  /* pipelined form: two MSMs in flight */
  OK(d377_msm_submit(sc, gel, D377_PT_ELEMENT, N, 0));
  OK(d377_msm_submit(sc, bases, D377_PT_BASES, N, 1));
  OK(d377_msm_wait(0, NULL, got));
  CHECK(memcmp(got, want, 32) == 0);
  OK(d377_msm_wait(1, NULL, got));
  CHECK(memcmp(got, want, 32) == 0);
  OK(d377_msm_bases_destroy(bases));
  CHECK(d377_msm_bases_destroy(bases) == D377_ERR_INVALID_ARG);

  /* many small MSMs in one call: pairs (i, G), segments [0, 8) and [8, 16): 28 G and 92 G */
  {
    uint32_t offs[3] = {0, 8, 16};
    uint8_t outs[2 * 32], k2[2 * 32] = {0}, w2[2 * 32], okm[2];
    k2[0] = 28;
    k2[32] = 92;
    OK(d377_fixed_base_mul(k2, 2, w2, D377_OUT_ENCODING));
    OK(d377_batch_msm(sc, gel, D377_PT_ELEMENT, offs, 2, outs, D377_OUT_ENCODING, okm));
    CHECK(memcmp(outs, w2, sizeof outs) == 0 && okm[0] == 1 && okm[1] == 1);
  }

  /* a larger MSM against the sum of the per-element products:
   * sum_i s_i * P_i == element_sum(batch_scalar_mul) */
  {
    static uint8_t s2[M * 32], prod[M * 128], sum_enc[32];
    memcpy(s2, r32, sizeof s2);
    for (int i = 0; i < M; i++) s2[32 * i + 31] &= 0x03;
    OK(d377_batch_encode_to_curve(r32, M, pts, D377_OUT_ELEMENT));
    OK(d377_batch_scalar_mul(pts, D377_PT_ELEMENT, s2, M, prod, D377_OUT_ELEMENT, NULL));
    OK(d377_element_sum(prod, M, NULL, sum_enc));
    OK(d377_msm(s2, pts, D377_PT_ELEMENT, M, NULL, got));
    CHECK(memcmp(got, sum_enc, 32) == 0);
    /* the same MSM over every GPU of the process (here: the one engine) */
    int dev0 = 0;
    OK(d377_init_multi(&dev0, 1));
    OK(d377_msm_multi(s2, pts, D377_PT_ELEMENT, M, 1, NULL, got));
    CHECK(memcmp(got, sum_enc, 32) == 0);
  }

  CHECK(d377_launch_count() > 0);
  OK(d377_shutdown());
  CHECK(d377_batch_compress(el, 1, enc) == D377_ERR_NOT_INITIALISED);
  if (failures) {
    printf("abi_smoke: %d check(s) failed\n", failures);
    return 1;
  }
  printf("abi_smoke: all checks passed\n");
  return 0;
}
