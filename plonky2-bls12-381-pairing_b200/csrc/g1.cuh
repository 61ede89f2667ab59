// g1.cuh -- G1 over the BASE field, in registers: Jacobian doubling / mixed addition / full
// addition, double-and-add ladders, the GLV endomorphism (x, y) -> (BETA x, y), the endomorphism subgroup test,
// cofactor clearing, and the building blocks of the bucket MSM.
//
// The group law is the ark-ec 0.4 short-Weierstrass Jacobian arithmetic the reference's native loop uses for
// `R + R` / `R + Q` (/root/reference/src/miller_loop_native_optimized.rs:93,98), restated for a = 0 over Fq:
// dbl-2009-l, madd-2007-bl, add-2007-bl.  Subgroup membership and cofactor clearing follow ark-bls12-381 0.4
// `curves/g1.rs` (eprint 2021/1130 section 6; eprint 2019/403 section 5), restated in the test oracle
// (g1_in_subgroup_fast, g1_clear_cofactor) and pinned there by the RFC 9380 effective cofactor.
// Round 1 ran G1 embedded in Fq2 through the slot arena (three times the multiplications plus the memory
// traffic); here a point is 36 words of registers and a mixed addition is 11 Fp multiplications.
//
// Value discipline: every coordinate is non-negative and below 2^384, so all products are the 12-word kind
// (144 + 156 IMAD.WIDE); differences go through the weak reduction (output below 1.02 p).
#pragma once
#include "helpers.cuh"

namespace b381 {

struct G1J { Fp x, y, z; };          // Jacobian; z = 0 (as a residue): identity
struct G1A { Fp x, y; };             // affine, internal format

struct G1Const { limb_t beta[NL]; limb_t four[NL]; };
#if defined(__CUDACC__)
static __constant__ G1Const g_g1c = {B381_G1_BETA, B381_FOUR};
#else
static const G1Const g_g1c = {B381_G1_BETA, B381_FOUR};
#endif

// group-level functions: inlined on the device (a point lives in registers), ordinary functions in the host
// simulation (always_inline there multiplies the compile time of the test build by ten)
#if defined(__CUDACC__)
#define B381_G1FN B381_DEV B381_INL
#else
#define B381_G1FN __attribute__((noinline))
#endif

B381_DEV B381_INL void r1_mul(Fp& r, const Fp& a, const Fp& b) { Fp t; fp_mul12(t, a, b); r = t; }
B381_DEV B381_INL void r1_sqr(Fp& r, const Fp& a) { Fp t; fp_mul12(t, a, a); r = t; }
B381_DEV B381_INL void r1_sub(Fp& r, const Fp& a, const Fp& b) { Fp t; fp_sub(t, a, b); fp_wreduce(t); r = t; }
// a value below 1.02 p (weak-reduced) is zero mod p iff it is 0 or p
B381_DEV B381_INL bool r1_is_zero_wr(const Fp& a) {
  uint32_t o0 = 0, op = 0;
#pragma unroll
  for (int k = 0; k < NL; k++) { o0 |= a.l[k]; op |= a.l[k] ^ pword(k); }
  return o0 == 0 || op == 0;
}
B381_DEV B381_INL bool r1_is_zero(const Fp& a) { Fp t = a; fp_wreduce(t); return r1_is_zero_wr(t); }

B381_DEV B381_INL void g1_set_identity(G1J& r) {
  fp_const(r.x, g_ct.one); fp_const(r.y, g_ct.one); fp_zero(r.z);      // (1, 1, 0) as ark-ec 0.4
}
B381_DEV B381_INL bool g1_is_identity(const G1J& p) { return r1_is_zero(p.z); }

// dbl-2009-l (a = 0): 2 M + 5 S
B381_G1FN void g1_double(G1J& r, const G1J& p) {
  Fp A, B, C, D, E, F, t;
  r1_sqr(A, p.x);
  r1_sqr(B, p.y);
  r1_sqr(C, B);
  fp_add(t, p.x, B);
  r1_sqr(t, t);
  fp_sub(t, t, A); fp_sub(t, t, C); fp_dbl(t, t);
  fp_wreduce(t); D = t;                               // D = 2 ((X + B)^2 - A - C)
  fp_dbl(E, A); fp_add(E, E, A);                      // E = 3 A
  r1_sqr(F, E);
  Fp z3;
  r1_mul(z3, p.y, p.z); fp_dbl(z3, z3);               // Z3 = 2 Y Z  (before X / Y are overwritten: r may alias p)
  fp_dbl(t, D);
  r1_sub(r.x, F, t);                                  // X3 = F - 2 D
  r1_sub(t, D, r.x);
  r1_mul(t, E, t);
  fp_dbl(C, C); fp_dbl(C, C); fp_dbl(C, C);           // 8 C
  r1_sub(r.y, t, C);                                  // Y3 = E (D - X3) - 8 C
  r.z = z3;
}

// madd-2007-bl: r = p + q, q affine and NOT the identity: 7 M + 4 S.  Handles p = identity, p = q (doubling) and
// p = -q (identity) like ark-ec's add_assign.
B381_G1FN void g1_add_mixed(G1J& r, const G1J& p, const G1A& q) {
  if (g1_is_identity(p)) { r.x = q.x; r.y = q.y; fp_const(r.z, g_ct.one); return; }
  Fp Z1Z1, U2, S2, H, HH, I, J, rr, V, t;
  r1_sqr(Z1Z1, p.z);
  r1_mul(U2, q.x, Z1Z1);
  r1_mul(S2, q.y, p.z); r1_mul(S2, S2, Z1Z1);
  r1_sub(H, U2, p.x);
  r1_sub(rr, S2, p.y);
  if (r1_is_zero_wr(H)) {
    if (r1_is_zero_wr(rr)) { G1J d = p; g1_double(r, d); return; }
    g1_set_identity(r);
    return;
  }
  fp_dbl(rr, rr);                                     // r = 2 (S2 - Y1)
  r1_sqr(HH, H);
  fp_dbl(I, HH); fp_dbl(I, I);                        // I = 4 HH
  r1_mul(J, H, I);
  r1_mul(V, p.x, I);
  Fp x3, y3, z3;
  r1_sqr(t, rr);
  fp_sub(t, t, J); fp_sub(t, t, V); fp_sub(t, t, V);
  fp_wreduce(t); x3 = t;                              // X3 = r^2 - J - 2 V
  r1_sub(t, V, x3);
  r1_mul(t, rr, t);
  r1_mul(J, p.y, J); fp_dbl(J, J);
  r1_sub(y3, t, J);                                   // Y3 = r (V - X3) - 2 Y1 J
  fp_add(t, p.z, H);
  r1_sqr(t, t);
  fp_sub(t, t, Z1Z1); fp_sub(t, t, HH);
  fp_wreduce(t); z3 = t;                              // Z3 = (Z1 + H)^2 - Z1Z1 - HH
  r.x = x3; r.y = y3; r.z = z3;
}

// add-2007-bl: r = p + q, both Jacobian: 11 M + 5 S
B381_G1FN void g1_add(G1J& r, const G1J& p, const G1J& q) {
  if (g1_is_identity(p)) { r = q; return; }
  if (g1_is_identity(q)) { r = p; return; }
  Fp Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t;
  r1_sqr(Z1Z1, p.z);
  r1_sqr(Z2Z2, q.z);
  r1_mul(U1, p.x, Z2Z2);
  r1_mul(U2, q.x, Z1Z1);
  r1_mul(S1, p.y, q.z); r1_mul(S1, S1, Z2Z2);
  r1_mul(S2, q.y, p.z); r1_mul(S2, S2, Z1Z1);
  r1_sub(H, U2, U1);
  r1_sub(rr, S2, S1);
  if (r1_is_zero_wr(H)) {
    if (r1_is_zero_wr(rr)) { G1J d = p; g1_double(r, d); return; }
    g1_set_identity(r);
    return;
  }
  fp_dbl(rr, rr);
  fp_dbl(I, H); r1_sqr(I, I);                         // I = (2 H)^2
  r1_mul(J, H, I);
  r1_mul(V, U1, I);
  Fp x3, y3, z3;
  r1_sqr(t, rr);
  fp_sub(t, t, J); fp_sub(t, t, V); fp_sub(t, t, V);
  fp_wreduce(t); x3 = t;
  r1_sub(t, V, x3);
  r1_mul(t, rr, t);
  r1_mul(J, S1, J); fp_dbl(J, J);
  r1_sub(y3, t, J);
  fp_add(t, p.z, q.z);
  r1_sqr(t, t);
  fp_sub(t, t, Z1Z1); fp_sub(t, t, Z2Z2);
  fp_wreduce(t);
  r1_mul(z3, t, H);                                   // Z3 = ((Z1 + Z2)^2 - Z1Z1 - Z2Z2) H
  r.x = x3; r.y = y3; r.z = z3;
}

B381_DEV B381_INL void g1_neg(G1J& r, const G1J& p) { r.x = p.x; r.z = p.z; Fp t; fp_neg(t, p.y); fp_wreduce(t); r.y = t; }

// r = [|x|] p, |x| = 0xd201000000010000 (public: uniform control flow; 63 doublings + 5 additions)
B381_G1FN void g1_mul_x_abs(G1J& r, const G1A& p) {
  G1J acc;
  acc.x = p.x; acc.y = p.y; fp_const(acc.z, g_ct.one);
  const uint64_t xabs = B381_X_ABS;
  for (int b = 62; b >= 0; b--) {
    g1_double(acc, acc);
    if ((xabs >> b) & 1) g1_add_mixed(acc, acc, p);
  }
  r = acc;
}

// r = [|x|] p for a JACOBIAN base (5 full additions): the second ladder of the subgroup test, no inversion in between
B381_G1FN void g1_mul_x_abs_jac(G1J& r, const G1J& p) {
  G1J acc = p;
  const uint64_t xabs = B381_X_ABS;
  for (int b = 62; b >= 0; b--) {
    g1_double(acc, acc);
    if ((xabs >> b) & 1) g1_add(acc, acc, p);
  }
  r = acc;
}

// Jacobian -> affine (x / z^2, y / z^3); returns false for the identity
B381_G1FN bool g1_to_affine(G1A& r, const G1J& p) {
  if (g1_is_identity(p)) return false;
  Fp zi, zi2;
  fp_inv_safegcd(zi, p.z);
  r1_sqr(zi2, zi);
  r1_mul(r.x, p.x, zi2);
  r1_mul(zi2, zi2, zi);
  r1_mul(r.y, p.y, zi2);
  return true;
}

// external affine (24 words) -> internal; returns error bits
B381_G1FN int g1_load_ext(G1A& r, const uint32_t* pt) {
  uint32_t w0[12], w1[12];
  for (int j = 0; j < 12; j++) { w0[j] = pt[j]; w1[j] = pt[12 + j]; }
  const bool ok = fp_from_ext(r.x, w0) & fp_from_ext(r.y, w1);
  return ok ? 0 : HERR_NOT_CANONICAL;
}
B381_G1FN void g1_store_ext(uint32_t* out, const G1A& p) {
  uint32_t w[12];
  fp_to_ext(w, p.x);
  for (int j = 0; j < 12; j++) out[j] = w[j];
  fp_to_ext(w, p.y);
  for (int j = 0; j < 12; j++) out[12 + j] = w[j];
}
// store a Jacobian point as affine external words + identity flag
B381_G1FN void g1_store_jac_ext(uint32_t* out, uint8_t* out_inf, const G1J& p) {
  G1A a;
  if (!g1_to_affine(a, p)) {
    for (int j = 0; j < 24; j++) out[j] = 0;
    *out_inf = 1;
    return;
  }
  g1_store_ext(out, a);
  *out_inf = 0;
}

// is (X, Y, Z) equal to the affine point (ax, ay)?  X == ax Z^2, Y == ay Z^3
B381_G1FN bool g1_jac_equals_affine(const G1J& p, const Fp& ax, const Fp& ay) {
  if (g1_is_identity(p)) return false;
  Fp z2, t, u;
  r1_sqr(z2, p.z);
  r1_mul(t, ax, z2);
  r1_sub(u, t, p.x);
  if (!r1_is_zero_wr(u)) return false;
  r1_mul(z2, z2, p.z);
  r1_mul(t, ay, z2);
  r1_sub(u, t, p.y);
  return r1_is_zero_wr(u);
}

// ---- programs (one point per thread, external words in / out, error bits returned) ---------------------------
// ark g1.rs is_in_correct_subgroup_assuming_on_curve: endomorphism(P) == -[X^2] P, early-out [X] P == P
B381_G1FN int prog_g1_in_subgroup(const uint32_t* pt, int inf, uint8_t* out) {
  if (inf & 1) { *out = 1; return 0; }
  G1A p;
  int err = g1_load_ext(p, pt);
  G1J xp, x2p;
  g1_mul_x_abs(xp, p);
  if (g1_jac_equals_affine(xp, p.x, p.y)) { *out = 0; return err; }
  g1_mul_x_abs_jac(x2p, xp);                         // ([X] P = identity gives the identity, never equal to an affine point)
  Fp bx, ny, beta;
  fp_const(beta, g_g1c.beta);
  r1_mul(bx, p.x, beta);
  fp_neg(ny, p.y); fp_wreduce(ny);                   // -[X^2] P == (BETA x, y)  <=>  [X^2] P == (BETA x, -y)
  *out = g1_jac_equals_affine(x2p, bx, ny) ? 1 : 0;
  return err;
}

// ark g1.rs clear_cofactor: [1 - x] P = [|x|] P + P
B381_G1FN int prog_g1_clear_cofactor(const uint32_t* pt, int inf, uint32_t* out, uint8_t* out_inf) {
  if (inf & 1) { for (int j = 0; j < 24; j++) out[j] = 0; *out_inf = 1; return 0; }
  G1A p;
  int err = g1_load_ext(p, pt);
  G1J xp;
  g1_mul_x_abs(xp, p);
  g1_add_mixed(xp, xp, p);
  g1_store_jac_ext(out, out_inf, xp);
  return err;
}

// [k] P, k = 256-bit scalar (8 little-endian words), left-to-right double-and-add with mixed additions.
// Per-thread scalars: divergent control flow.  Not constant time.
B381_G1FN int prog_g1_scalar_mul(const uint32_t* pt, int inf, const uint32_t* k, uint32_t* out, uint8_t* out_inf) {
  uint32_t kk[8], nz = 0;
  for (int i = 0; i < 8; i++) { kk[i] = k[i]; nz |= kk[i]; }
  if ((inf & 1) || nz == 0) { for (int j = 0; j < 24; j++) out[j] = 0; *out_inf = 1; return 0; }
  G1A p;
  int err = g1_load_ext(p, pt);
  G1J acc;
  g1_set_identity(acc);
  bool started = false;
  for (int b = 255; b >= 0; b--) {
    const bool bit = (kk[b >> 5] >> (b & 31)) & 1u;
    if (started) g1_double(acc, acc);
    if (bit) { g1_add_mixed(acc, acc, p); started = true; }
  }
  g1_store_jac_ext(out, out_inf, acc);
  return err;
}

// ---- bucket method (Pippenger) building blocks: SURVEY 8f rank 4 ----------------------------------------------
// Points are pre-converted to the internal format once (12 words per coordinate: every window re-reads them);
// bucket sums and partial results are Jacobian points of 36 words (12 per coordinate) in global memory.
constexpr int G1_RAW_AFF = 24, G1_RAW_JAC = 36;

B381_DEV B381_INL void g1_ld_raw_aff(G1A& r, const uint32_t* src) {
#pragma unroll
  for (int k = 0; k < 12; k++) { r.x.l[k] = src[k]; r.y.l[k] = src[12 + k]; }
  r.x.l[12] = 0; r.y.l[12] = 0;
  B381_SETRANGE(r.x, 0.0, 1.03); B381_SETRANGE(r.y, 0.0, 1.03);
}
B381_DEV B381_INL void g1_st_raw_aff(uint32_t* dst, const G1A& p) {
#pragma unroll
  for (int k = 0; k < 12; k++) { dst[k] = p.x.l[k]; dst[12 + k] = p.y.l[k]; }
}
B381_DEV B381_INL void g1_ld_raw_jac(G1J& r, const uint32_t* src) {
#pragma unroll
  for (int k = 0; k < 12; k++) { r.x.l[k] = src[k]; r.y.l[k] = src[12 + k]; r.z.l[k] = src[24 + k]; }
  r.x.l[12] = 0; r.y.l[12] = 0; r.z.l[12] = 0;
  B381_SETRANGE(r.x, 0.0, 2.1); B381_SETRANGE(r.y, 0.0, 2.1); B381_SETRANGE(r.z, 0.0, 2.1);
}
B381_DEV B381_INL void g1_st_raw_jac(uint32_t* dst, const G1J& p) {
  B381_CHECK(p.x.l[12] == 0 && p.y.l[12] == 0 && p.z.l[12] == 0, "raw Jacobian store: thirteenth word");
#pragma unroll
  for (int k = 0; k < 12; k++) { dst[k] = p.x.l[k]; dst[12 + k] = p.y.l[k]; dst[24 + k] = p.z.l[k]; }
}

// digit w (c bits) of a 256-bit scalar held as 8 little-endian words
B381_DEV B381_INL uint32_t msm_digit(const uint32_t* k, int w, int c) {
  const int bit = w * c;
  if (bit >= 256) return 0;
  const int word = bit >> 5, sh = bit & 31;
  uint64_t v = k[word];
  if (word + 1 < 8) v |= (uint64_t)k[word + 1] << 32;
  return (uint32_t)(v >> sh) & ((1u << c) - 1u);
}

// sum of the affine raw points pts[idx[lo .. hi)] -> Jacobian
B381_G1FN void msm_bucket_sum(G1J& acc, const uint32_t* pts_raw, const uint32_t* idx, size_t lo, size_t hi) {
  g1_set_identity(acc);
  for (size_t t = lo; t < hi; t++) {
    G1A q;
    g1_ld_raw_aff(q, pts_raw + (size_t)G1_RAW_AFF * idx[t]);
    g1_add_mixed(acc, acc, q);
  }
}

// chunk [lo, hi) of the buckets of one window (bucket d holds the points whose digit is d):
//   sum_{d in chunk} d B_d = running-sum accumulation + (lo - 1) * (sum of the chunk's buckets)
B381_G1FN void msm_chunk_weighted(G1J& out, const uint32_t* buckets_raw, uint32_t lo, uint32_t hi) {
  G1J run, acc;
  g1_set_identity(run);
  g1_set_identity(acc);
  for (uint32_t d = hi; d-- > lo;) {
    G1J b;
    g1_ld_raw_jac(b, buckets_raw + (size_t)G1_RAW_JAC * d);
    g1_add(run, run, b);
    g1_add(acc, acc, run);                           // acc = sum (d - lo + 1) B_d
  }
  // + (lo - 1) * run
  const uint32_t m = lo - 1;
  if (m != 0) {
    G1J t;
    g1_set_identity(t);
    bool started = false;
    for (int bit = 31; bit >= 0; bit--) {
      if (started) g1_double(t, t);
      if ((m >> bit) & 1u) { g1_add(t, t, run); started = true; }
    }
    g1_add(acc, acc, t);
  }
  out = acc;
}

// Horner over the window sums S_0 .. S_{W-1}: R = sum 2^(c w) S_w
B381_G1FN void msm_combine_windows(G1J& r, const uint32_t* sums_raw, int W, int c) {
  G1J acc;
  g1_ld_raw_jac(acc, sums_raw + (size_t)G1_RAW_JAC * (W - 1));
  for (int w = W - 2; w >= 0; w--) {
    for (int i = 0; i < c; i++) g1_double(acc, acc);
    G1J s;
    g1_ld_raw_jac(s, sums_raw + (size_t)G1_RAW_JAC * w);
    g1_add(acc, acc, s);
  }
  r = acc;
}

}  // namespace b381
