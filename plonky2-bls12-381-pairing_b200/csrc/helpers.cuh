// helpers.cuh -- batched witness-generation helpers: the native computations the
// reference's circuit generators run per witness -- inverse, square root with a prescribed sign,
// Legendre symbol / is_square, exponentiation:
//   Fq   inverse / div    /root/reference/src/fields/fq_target.rs:316-343 (x.inverse()), :405-440
//   Fq   sqrt with sgn0   /root/reference/src/fields/fq_target.rs:316-343 (FqSqrtGenerator::run_once)
//   Fq   legendre, is_square, pow   fq_target.rs:243-280 ; pow_fq /root/reference/src/fields/helpers.rs:176-195
//   Fq2  inverse / sqrt with sgn0   /root/reference/src/fields/fq2_target.rs:320-352, :373-410 ; sgn0_fq2 helpers.rs:169-174
//   Fq6 / Fq12 inverse    /root/reference/src/fields/fq6_target.rs:384-418 ; fq12_target.rs:340-374 (tower.cuh f6_inv / f12_inv)
// One element per thread, registers only (no slot arena), external format in and out.  Both square
// roots of a non-zero element have opposite sgn0, so the result is unique whatever algorithm ark's
// `sqrt()` uses internally.  Requires the 13 x 32-bit format (fp32.cuh).
#pragma once
#include "tower.cuh"

#if B381_FMT == 32
namespace b381 {

enum HelperErr { HERR_NOT_CANONICAL = 1, HERR_ZERO_DIVISION = 2, HERR_NOT_SQUARE = 4 };

struct ExpTab {
  uint32_t pm2[12];      // p - 2          (inverse)
  uint32_t pp1d4[12];    // (p + 1) / 4    (square root, p = 3 mod 4)
  uint32_t pm1d2[12];    // (p - 1) / 2    (Legendre symbol)
};
#if defined(__CUDACC__)
__constant__ ExpTab g_et = {B381_EXP_PM2, B381_EXP_PP1D4, B381_EXP_PM1D2};
#else
static const ExpTab g_et = {B381_EXP_PM2, B381_EXP_PP1D4, B381_EXP_PM1D2};
#endif

B381_DEV B381_INL void fp_one(Fp& r) { fp_const(r, g_ct.one); }

// r = a^e, e = nwords x 32-bit little-endian words shared by the whole batch (uniform control flow);
// plain left-to-right square-and-multiply starting at the leading one.  e = 0 gives 1.
B381_DEV B381_INL void fp_pow_words(Fp& r, const Fp& a, const uint32_t* e, int nwords) {
  Fp x;
  fp_one(x);
  bool started = false;
  for (int i = nwords * 32 - 1; i >= 0; i--) {
    const bool bit = (e[i >> 5] >> (i & 31)) & 1u;
    if (started) { Fp t; fp_mul(t, x, x); x = t; }
    if (bit) {
      if (started) { Fp t; fp_mul(t, x, a); x = t; } else { x = a; started = true; }
    }
  }
  r = x;
}

// canonical NON-Montgomery integer of a stored value (for parity / sgn0): a~ * 1 / R' = a
B381_DEV B381_INL void fp_plain(Fp& r, const Fp& a) {
  Fp one;
  fp_zero(one);
  one.l[0] = 1;
  B381_TB(one.mag = 1e-30; one.lb = 0;)
  fp_mul(r, a, one);
  fp_canon_small(r);
}

B381_DEV B381_INL bool fp_is_zero_any(const Fp& a) { Fp t = a; fp_canon(t); return fp_is_zero_canon(t); }
B381_DEV B381_INL bool fp_equal_any(const Fp& a, const Fp& b) {
  Fp t;
  fp_sub(t, a, b);
  fp_canon(t);
  return fp_is_zero_canon(t);
}
// -a as a non-negative value (a below 128 p)
B381_DEV B381_INL void fp_neg_nn(Fp& r, const Fp& a) { fp_neg(r, a); fp_add_p128(r, r); }

// sgn0_fq (helpers.rs:159-167): parity of the canonical integer
B381_DEV B381_INL bool fp_sgn0(const Fp& a) { Fp t; fp_plain(t, a); return (t.l[0] & 1u) != 0; }

// square root in Fp (p = 3 mod 4): s = a^((p+1)/4); returns false if a is not a square
B381_DEV B381_INL bool fp_sqrt_any(Fp& s, const Fp& a) {
  fp_pow_words(s, a, g_et.pp1d4, 12);
  Fp q;
  fp_mul(q, s, s);
  return fp_equal_any(q, a);
}

// ---- per-element programs: external words in, external words out, error bits returned --------------
B381_DEV B381_INL int prog_fp_inv(const uint32_t* a, uint32_t* out) {
  uint32_t w[12], wo[12];
  for (int j = 0; j < 12; j++) w[j] = a[j];
  Fp x, r;
  int err = fp_from_ext(x, w) ? 0 : HERR_NOT_CANONICAL;
  uint32_t nz = 0;
  for (int j = 0; j < 12; j++) nz |= w[j];
  if (nz == 0) err |= HERR_ZERO_DIVISION;            // ark: inverse() of zero is None, the reference unwraps
  fp_pow_words(r, x, g_et.pm2, 12);
  fp_to_ext(wo, r);
  for (int j = 0; j < 12; j++) out[j] = wo[j];
  return err;
}

// out = a^e with the reference's pow_fq semantics (helpers.rs:176-195): the NAF ladder starts from
// res = a, so an all-zero exponent returns a (not 1)
B381_DEV B381_INL int prog_fp_pow(const uint32_t* a, const uint32_t* e, int nwords, uint32_t* out) {
  uint32_t w[12], wo[12];
  for (int j = 0; j < 12; j++) w[j] = a[j];
  Fp x, r;
  int err = fp_from_ext(x, w) ? 0 : HERR_NOT_CANONICAL;
  uint32_t nz = 0;
  for (int j = 0; j < nwords; j++) nz |= e[j];
  if (nz == 0) r = x; else fp_pow_words(r, x, e, nwords);
  fp_to_ext(wo, r);
  for (int j = 0; j < 12; j++) out[j] = wo[j];
  return err;
}

// legendre(a) == 1 (fq_target.rs:269-280): zero is NOT a square by that definition
B381_DEV B381_INL int prog_fp_is_square(const uint32_t* a, uint8_t* out) {
  uint32_t w[12];
  for (int j = 0; j < 12; j++) w[j] = a[j];
  Fp x, l, one;
  int err = fp_from_ext(x, w) ? 0 : HERR_NOT_CANONICAL;
  fp_pow_words(l, x, g_et.pm1d2, 12);
  fp_one(one);
  *out = fp_equal_any(l, one) ? 1 : 0;
  return err;
}

// FqSqrtGenerator::run_once (fq_target.rs:316-343): sqrt(x), negated if its sgn0 differs from sgn
B381_DEV B381_INL int prog_fp_sqrt(const uint32_t* a, int sgn, uint32_t* out) {
  uint32_t w[12], wo[12];
  for (int j = 0; j < 12; j++) w[j] = a[j];
  Fp x, s;
  int err = fp_from_ext(x, w) ? 0 : HERR_NOT_CANONICAL;
  if (!fp_sqrt_any(s, x)) err |= HERR_NOT_SQUARE;   // the reference: x.sqrt().unwrap() panics
  const bool have = fp_sgn0(s);
  if (have != (sgn != 0)) {
    if (fp_is_zero_any(s)) err |= HERR_NOT_SQUARE;  // sqrt = 0 cannot have sgn0 = 1 (the reference's assert_eq fails)
    Fp t;
    fp_neg_nn(t, s);
    s = t;
  }
  fp_to_ext(wo, s);
  for (int j = 0; j < 12; j++) out[j] = wo[j];
  return err;
}

B381_DEV B381_INL int prog_fp2_inv(const uint32_t* a, uint32_t* out) {
  uint32_t w0[12], w1[12], wo[12];
  for (int j = 0; j < 12; j++) { w0[j] = a[j]; w1[j] = a[12 + j]; }
  Fp a0, a1, n, ni, r0, r1, t;
  int err = (fp_from_ext(a0, w0) && fp_from_ext(a1, w1)) ? 0 : HERR_NOT_CANONICAL;
  uint32_t nz = 0;
  for (int j = 0; j < 12; j++) nz |= w0[j] | w1[j];
  if (nz == 0) err |= HERR_ZERO_DIVISION;
  Acc T;
  acc_mul(T, a0, a0);
  acc_mac(T, a1, a1);
  acc_redc(n, T);                                    // norm a0^2 + a1^2
  fp_pow_words(ni, n, g_et.pm2, 12);
  fp_mul(r0, a0, ni);
  fp_neg_nn(t, a1);
  fp_mul(r1, t, ni);
  fp_to_ext(wo, r0);
  for (int j = 0; j < 12; j++) out[j] = wo[j];
  fp_to_ext(wo, r1);
  for (int j = 0; j < 12; j++) out[12 + j] = wo[j];
  return err;
}

B381_DEV B381_INL void f2_sqr_fp(Fp& r0, Fp& r1, const Fp& a0, const Fp& a1) { f2_sqr_reg(r0, r1, a0, a1); }

// Fq2 square root by the norm method, then the sign rule of sgn0_fq2 (helpers.rs:169-174):
//   a1 = 0: sqrt(a0) or u sqrt(-a0);  else n = a0^2 + a1^2, s = sqrt(n), d = (a0 +- s)/2 (the one that
//   is a square), c0 = sqrt(d), c1 = a1 / (2 c0).  The result is verified by squaring.
B381_DEV B381_INL int prog_fp2_sqrt(const uint32_t* a, int sgn, uint32_t* out) {
  uint32_t w0[12], w1[12], wo[12];
  for (int j = 0; j < 12; j++) { w0[j] = a[j]; w1[j] = a[12 + j]; }
  Fp a0, a1, c0, c1;
  int err = (fp_from_ext(a0, w0) && fp_from_ext(a1, w1)) ? 0 : HERR_NOT_CANONICAL;
  uint32_t nz1 = 0;
  for (int j = 0; j < 12; j++) nz1 |= w1[j];
  if (nz1 == 0) {
    Fp s, m;
    if (fp_sqrt_any(s, a0)) { c0 = s; fp_zero(c1); }
    else { fp_neg_nn(m, a0); fp_sqrt_any(s, m); fp_zero(c0); c1 = s; }      // -1 is a non-residue: -a0 is a square
  } else {
    Fp n, s, d, t, h, ti;
    Acc T;
    acc_mul(T, a0, a0);
    acc_mac(T, a1, a1);
    acc_redc(n, T);
    if (!fp_sqrt_any(s, n)) err |= HERR_NOT_SQUARE;
    fp_add(d, a0, s);
    fp_half(h, d);                                   // (a0 + s) / 2
    if (!fp_sqrt_any(t, h)) {
      fp_neg_nn(d, s);
      fp_add(d, d, a0);
      fp_half(h, d);                                 // (a0 - s) / 2
      fp_sqrt_any(t, h);
    }
    c0 = t;
    fp_dbl(d, t);
    fp_pow_words(ti, d, g_et.pm2, 12);               // 1 / (2 c0)
    fp_mul(c1, a1, ti);
  }
  {                                                  // verify (c0 + c1 u)^2 == a
    Fp q0, q1;
    f2_sqr_fp(q0, q1, c0, c1);
    if (!(fp_equal_any(q0, a0) && fp_equal_any(q1, a1))) err |= HERR_NOT_SQUARE;
  }
  const bool z0 = fp_is_zero_any(c0);
  const bool have = fp_sgn0(c0) || (z0 && fp_sgn0(c1));
  if (have != (sgn != 0)) {
    if (z0 && fp_is_zero_any(c1)) err |= HERR_NOT_SQUARE;
    Fp t;
    fp_neg_nn(t, c0); c0 = t;
    fp_neg_nn(t, c1); c1 = t;
  }
  fp_to_ext(wo, c0);
  for (int j = 0; j < 12; j++) out[j] = wo[j];
  fp_to_ext(wo, c1);
  for (int j = 0; j < 12; j++) out[12 + j] = wo[j];
  return err;
}

// Fq2 is_square: the norm is a square in Fq (a^((p^2-1)/2) = norm^((p-1)/2)); zero is not
B381_DEV B381_INL int prog_fp2_is_square(const uint32_t* a, uint8_t* out) {
  uint32_t w0[12], w1[12];
  for (int j = 0; j < 12; j++) { w0[j] = a[j]; w1[j] = a[12 + j]; }
  Fp a0, a1, n, l, one;
  int err = (fp_from_ext(a0, w0) && fp_from_ext(a1, w1)) ? 0 : HERR_NOT_CANONICAL;
  Acc T;
  acc_mul(T, a0, a0);
  acc_mac(T, a1, a1);
  acc_redc(n, T);
  fp_pow_words(l, n, g_et.pm1d2, 12);
  fp_one(one);
  *out = fp_equal_any(l, one) ? 1 : 0;
  return err;
}

}  // namespace b381
#endif  // B381_FMT == 32
