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

namespace b381 {

enum HelperErr { HERR_NOT_CANONICAL = 1, HERR_ZERO_DIVISION = 2, HERR_NOT_SQUARE = 4 };

struct ExpTab {
  uint32_t pm2[12];      // p - 2          (inverse)
  uint32_t pp1d4[12];    // (p + 1) / 4    (square root, p = 3 mod 4)
  uint32_t pm1d2[12];    // (p - 1) / 2    (Legendre symbol)
};
#if defined(__CUDACC__)
static __constant__ ExpTab g_et = {B381_EXP_PM2, B381_EXP_PP1D4, B381_EXP_PM1D2};
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
  B381_SETRANGE(one, 0.0, 1e-30);
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
  fp_inv_safegcd(r, x);
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
  fp_inv_safegcd(ni, n);
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
    fp_inv_safegcd(ti, d);               // 1 / (2 c0)
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

// ---------------------------------------------------------------------------------------------------
// Wire formats
// ---------------------------------------------------------------------------------------------------
namespace b381 {

// (a) the reference's own witness format: the canonical (non-Montgomery) integer as 12 x 32-bit
// little-endian digits -- `let value_b: BigUint = value.into(); value_b.to_u32_digits()` padded to 12,
// /root/reference/src/fields/fq_target.rs:300-313; the inverse is from_biguint_to_fq, helpers.rs:154-157.
B381_DEV B381_INL int prog_fp_to_digits(const uint32_t* a, uint32_t* out) {
  uint32_t w[12];
  for (int j = 0; j < 12; j++) w[j] = a[j];
  Fp x, r;
  int err = fp_from_ext(x, w) ? 0 : HERR_NOT_CANONICAL;
  fp_plain(r, x);
  for (int j = 0; j < 12; j++) out[j] = r.l[j];
  return err;
}

B381_DEV B381_INL void fp_from_plain(Fp& r, const Fp& x) {   // plain integer X < 2^384 -> X 2^416 mod p
  const uint32_t r2[NL] = B381_R2;
  Fp c;
  fp_set(c, r2);
  fp_mul(r, x, c);
}

B381_DEV B381_INL int prog_fp_from_digits(const uint32_t* d, uint32_t* out) {
  uint32_t w[12], wo[12];
  for (int j = 0; j < 12; j++) w[j] = d[j];
  Fp x, r;
  fp_unpack32(x, w);
  int err = fp_below_p(x) ? 0 : HERR_NOT_CANONICAL;   // Fq::from_bigint(..).unwrap() panics when >= p
  fp_from_plain(r, x);
  fp_to_ext(wo, r);
  for (int j = 0; j < 12; j++) out[j] = wo[j];
  return err;
}

// Fq12Target::set_witness (fq12_target.rs:408-416): MyFq12 order (helpers.rs:39-41) x 12 digits each.
// in: Fq12 in tower order (144 words, Montgomery); out: 144 digits, coefficient k of the w-basis at 12 k.
B381_DEV B381_INL int prog_fp12_to_witness(const uint32_t* f, uint32_t* out) {
  int err = 0;
  for (int pos = 0; pos < 12; pos++) {
    const int k = pos / 6, q = pos % 6, i = q & 1, j = q >> 1;     // w-basis slot -> c_i.c_j.c_k
    err |= prog_fp_to_digits(f + 12 * ((i * 3 + j) * 2 + k), out + 12 * pos);
  }
  return err;
}

// (b) ZCash / IETF point encodings as used by ark-bls12-381 0.4 and zkcrypto (big-endian x; top bits of
// byte 0: 0x80 compressed, 0x40 infinity, 0x20 y lexicographically largest).  No subgroup check.
enum { HERR_BAD_ENCODING = 8 };

B381_DEV B381_INL void be48_to_words(uint32_t (&w)[12], const uint8_t* b, uint8_t mask0) {
  for (int j = 0; j < 12; j++) {
    const uint8_t* q = b + 44 - 4 * j;
    uint32_t b0 = q[0];
    if (j == 11) b0 &= mask0;
    w[j] = (b0 << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
  }
}
B381_DEV B381_INL void words_to_be48(uint8_t* b, const uint32_t (&w)[12]) {
  for (int j = 0; j < 12; j++) {
    uint8_t* q = b + 44 - 4 * j;
    q[0] = (uint8_t)(w[j] >> 24); q[1] = (uint8_t)(w[j] >> 16); q[2] = (uint8_t)(w[j] >> 8); q[3] = (uint8_t)w[j];
  }
}
// plain canonical words of a stored value
B381_DEV B381_INL void fp_plain_words(uint32_t (&w)[12], const Fp& a) {
  Fp t;
  fp_plain(t, a);
  for (int j = 0; j < 12; j++) w[j] = t.l[j];
}
// y > (p - 1) / 2 as plain integers
B381_DEV B381_INL bool words_lex_largest(const uint32_t (&w)[12]) {
  const uint32_t h[12] = B381_PM1D2_WORDS;
  for (int j = 11; j >= 0; j--) {
    if (w[j] > h[j]) return true;
    if (w[j] < h[j]) return false;
  }
  return false;
}
B381_DEV B381_INL bool words_zero(const uint32_t (&w)[12]) {
  uint32_t o = 0;
  for (int j = 0; j < 12; j++) o |= w[j];
  return o == 0;
}
B381_DEV B381_INL bool bytes_zero(const uint8_t* b, int n) {
  uint32_t o = 0;
  for (int j = 0; j < n; j++) o |= b[j];
  return o == 0;
}
// internal value of the small constant 4
B381_DEV B381_INL void fp_four(Fp& r) {
  Fp x;
  fp_zero(x);
  x.l[0] = 4;
  B381_SETRANGE(x, 0.0, 1e-30);
  fp_from_plain(r, x);
}

// G1: bytes -> affine (x, y) in the C-ABI layout (24 words, Montgomery) + infinity flag
B381_DEV B381_INL int prog_g1_deserialize(const uint8_t* in, int compressed, uint32_t* g1, uint8_t* inf) {
  const uint8_t f = in[0];
  const bool fc = (f & 0x80) != 0, fi = (f & 0x40) != 0, fs = (f & 0x20) != 0;
  int err = 0;
  if (fc != (compressed != 0)) err |= HERR_BAD_ENCODING;
  const int len = compressed ? 48 : 96;
  uint32_t wx[12], wy[12], wo[12];
  be48_to_words(wx, in, 0x1f);
  if (fi) {                                          // infinity: every other bit must be clear
    if (fs || !words_zero(wx) || !bytes_zero(in + 48, len - 48)) err |= HERR_BAD_ENCODING;
    for (int j = 0; j < 24; j++) g1[j] = 0;
    *inf = 1;
    return err;
  }
  *inf = 0;
  Fp xp, x, y, t, rhs, b4;
  fp_unpack32(xp, wx);
  if (!fp_below_p(xp)) err |= HERR_NOT_CANONICAL;
  fp_from_plain(x, xp);
  fp_mul(t, x, x);
  fp_mul(rhs, t, x);
  fp_four(b4);
  fp_add(rhs, rhs, b4);                              // x^3 + 4
  if (compressed) {
    if (!fp_sqrt_any(y, rhs)) err |= HERR_NOT_SQUARE;          // x is not on the curve
    uint32_t py[12];
    fp_plain_words(py, y);
    if (words_lex_largest(py) != fs) { fp_neg_nn(t, y); y = t; }
  } else {
    if (fs) err |= HERR_BAD_ENCODING;
    Fp yp;
    be48_to_words(wy, in + 48, 0xff);
    fp_unpack32(yp, wy);
    if (!fp_below_p(yp)) err |= HERR_NOT_CANONICAL;
    fp_from_plain(y, yp);
    fp_mul(t, y, y);
    if (!fp_equal_any(t, rhs)) err |= HERR_NOT_SQUARE;         // not on the curve
  }
  fp_to_ext(wo, x);
  for (int j = 0; j < 12; j++) g1[j] = wo[j];
  fp_to_ext(wo, y);
  for (int j = 0; j < 12; j++) g1[12 + j] = wo[j];
  return err;
}

B381_DEV B381_INL int prog_g1_serialize(const uint32_t* g1, int inf, int compressed, uint8_t* out) {
  const int len = compressed ? 48 : 96;
  if (inf & 1) {
    for (int j = 0; j < len; j++) out[j] = 0;
    out[0] = (uint8_t)((compressed ? 0x80 : 0) | 0x40);
    return 0;
  }
  uint32_t w[12], px[12], py[12];
  Fp x, y;
  int err = 0;
  for (int j = 0; j < 12; j++) w[j] = g1[j];
  if (!fp_from_ext(x, w)) err |= HERR_NOT_CANONICAL;
  for (int j = 0; j < 12; j++) w[j] = g1[12 + j];
  if (!fp_from_ext(y, w)) err |= HERR_NOT_CANONICAL;
  fp_plain_words(px, x);
  fp_plain_words(py, y);
  words_to_be48(out, px);
  if (compressed) out[0] |= (uint8_t)(0x80 | (words_lex_largest(py) ? 0x20 : 0));
  else words_to_be48(out + 48, py);
  return err;
}

// Fq2 lexicographic order (zkcrypto Fp2::lexicographically_largest): c1 first, then c0
B381_DEV B381_INL bool f2_lex_largest(const Fp& c0, const Fp& c1) {
  uint32_t p0[12], p1[12];
  fp_plain_words(p0, c0);
  fp_plain_words(p1, c1);
  return words_lex_largest(p1) || (words_zero(p1) && words_lex_largest(p0));
}

// register-level Fq2 square root (any root); false if a is not a square
B381_DEV B381_INL bool f2_sqrt_any(Fp& c0, Fp& c1, const Fp& a0, const Fp& a1) {
  bool ok = true;
  if (fp_is_zero_any(a1)) {
    Fp s, m;
    if (fp_sqrt_any(s, a0)) { c0 = s; fp_zero(c1); }
    else { fp_neg_nn(m, a0); ok = fp_sqrt_any(s, m); fp_zero(c0); c1 = s; }
  } else {
    Fp n, s, d, t, h, ti;
    Acc T;
    acc_mul(T, a0, a0);
    acc_mac(T, a1, a1);
    acc_redc(n, T);
    ok = fp_sqrt_any(s, n);
    fp_add(d, a0, s);
    fp_half(h, d);
    if (!fp_sqrt_any(t, h)) {
      fp_neg_nn(d, s);
      fp_add(d, d, a0);
      fp_half(h, d);
      fp_sqrt_any(t, h);
    }
    c0 = t;
    fp_dbl(d, t);
    fp_inv_safegcd(ti, d);
    fp_mul(c1, a1, ti);
  }
  Fp q0, q1;
  f2_sqr_reg(q0, q1, c0, c1);
  return ok && fp_equal_any(q0, a0) && fp_equal_any(q1, a1);
}

// G2: x = x.c1 || x.c0 (flags in the first byte), twist curve y^2 = x^3 + 4 (1 + u)
B381_DEV B381_INL int prog_g2_deserialize(const uint8_t* in, int compressed, uint32_t* g2, uint8_t* inf) {
  const uint8_t f = in[0];
  const bool fc = (f & 0x80) != 0, fi = (f & 0x40) != 0, fs = (f & 0x20) != 0;
  int err = 0;
  if (fc != (compressed != 0)) err |= HERR_BAD_ENCODING;
  const int len = compressed ? 96 : 192;
  uint32_t w1[12], w0[12], wo[12];
  be48_to_words(w1, in, 0x1f);
  be48_to_words(w0, in + 48, 0xff);
  if (fi) {
    if (fs || !words_zero(w1) || !bytes_zero(in + 48, len - 48)) err |= HERR_BAD_ENCODING;
    for (int j = 0; j < 48; j++) g2[j] = 0;
    *inf = 1;
    return err;
  }
  *inf = 0;
  Fp p0, p1, x0, x1, y0, y1, s0, s1, r0, r1, b4;
  fp_unpack32(p0, w0); fp_unpack32(p1, w1);
  if (!fp_below_p(p0) || !fp_below_p(p1)) err |= HERR_NOT_CANONICAL;
  fp_from_plain(x0, p0); fp_from_plain(x1, p1);
  f2_sqr_reg(s0, s1, x0, x1);
  f2_mul_reg(r0, r1, s0, s1, x0, x1);
  fp_four(b4);
  fp_add(r0, r0, b4); fp_add(r1, r1, b4);            // x^3 + 4 + 4u
  if (compressed) {
    if (!f2_sqrt_any(y0, y1, r0, r1)) err |= HERR_NOT_SQUARE;
    if (f2_lex_largest(y0, y1) != fs) { Fp t; fp_neg_nn(t, y0); y0 = t; fp_neg_nn(t, y1); y1 = t; }
  } else {
    if (fs) err |= HERR_BAD_ENCODING;
    uint32_t v1[12], v0[12];
    Fp q0, q1, t0, t1;
    be48_to_words(v1, in + 96, 0xff);
    be48_to_words(v0, in + 144, 0xff);
    fp_unpack32(q0, v0); fp_unpack32(q1, v1);
    if (!fp_below_p(q0) || !fp_below_p(q1)) err |= HERR_NOT_CANONICAL;
    fp_from_plain(y0, q0); fp_from_plain(y1, q1);
    f2_sqr_reg(t0, t1, y0, y1);
    if (!(fp_equal_any(t0, r0) && fp_equal_any(t1, r1))) err |= HERR_NOT_SQUARE;
  }
  fp_to_ext(wo, x0); for (int j = 0; j < 12; j++) g2[j] = wo[j];
  fp_to_ext(wo, x1); for (int j = 0; j < 12; j++) g2[12 + j] = wo[j];
  fp_to_ext(wo, y0); for (int j = 0; j < 12; j++) g2[24 + j] = wo[j];
  fp_to_ext(wo, y1); for (int j = 0; j < 12; j++) g2[36 + j] = wo[j];
  return err;
}

B381_DEV B381_INL int prog_g2_serialize(const uint32_t* g2, int inf, int compressed, uint8_t* out) {
  const int len = compressed ? 96 : 192;
  if (inf & 1) {
    for (int j = 0; j < len; j++) out[j] = 0;
    out[0] = (uint8_t)((compressed ? 0x80 : 0) | 0x40);
    return 0;
  }
  int err = 0;
  Fp c[4];
  uint32_t w[12], pw[12];
  for (int k = 0; k < 4; k++) {
    for (int j = 0; j < 12; j++) w[j] = g2[12 * k + j];
    if (!fp_from_ext(c[k], w)) err |= HERR_NOT_CANONICAL;
  }
  fp_plain_words(pw, c[1]); words_to_be48(out, pw);          // x.c1
  fp_plain_words(pw, c[0]); words_to_be48(out + 48, pw);     // x.c0
  if (compressed) out[0] |= (uint8_t)(0x80 | (f2_lex_largest(c[2], c[3]) ? 0x20 : 0));
  else {
    fp_plain_words(pw, c[3]); words_to_be48(out + 96, pw);
    fp_plain_words(pw, c[2]); words_to_be48(out + 144, pw);
  }
  return err;
}

}  // namespace b381
