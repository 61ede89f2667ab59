/* b381_ref.c -- CPU restatement in C of the reference's native pairing path (ARK mode), used as
 * (1) a second, independent checker for the CUDA path at sizes the Python oracle cannot reach and
 * (2) the timed CPU baseline of bench.py (`cpu_baseline`, `--impl reference`), threaded over the batch.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY: the product never links or loads this file.
 *
 * The reference (NikolayKostadinov21/plonky2-bls12-381-pairing, /root/reference) cannot be built
 * here (no cargo/rustc, un-vendored crates), so this is a "port": it restates
 *   - ark-ff 0.4 Fp384 arithmetic: 6 x u64 little-endian limbs, Montgomery R = 2^384 (the layout the
 *     reference's types hold, src/fields/helpers.rs:8-11);
 *   - the tower formulas of src/fields_as_trees/fq{2,6,12}_target_tree.rs (cited per function);
 *   - ark-ec 0.4 Bls12::multi_miller_loop / final_exponentiation, the
 *     truth the reference defers to (src/miller_loop_native_optimized.rs:131-132,151,163).
 * It is validated against oracle/b381_oracle.py and the golden vectors by tests/test_oracle.py.
 * Parity at the pairing level is unpinned by the reference itself -- see DESIGN.md.
 *
 * Buffers: identical to include/b381.h (12 x u32 = 6 x u64 LE per Fp, Montgomery R = 2^384).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[6]; } fp;
typedef struct { fp c0, c1; } fp2;
typedef struct { fp2 c0, c1, c2; } fp6;
typedef struct { fp6 c0, c1; } fp12;

static const fp P = {{0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL,
                      0x64774b84f38512bfULL, 0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL}};
static const uint64_t N0 = 0x89f3fffcfffcfffdULL;                 /* -p^-1 mod 2^64 */
static const fp ONE = {{0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL,
                        0x77ce585370525745ULL, 0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL}};   /* R mod p */
static const uint64_t X_ABS = 0xd201000000010000ULL;             /* src/global_constants.rs:7 */

/* ---- Fp ---------------------------------------------------------------------------------------- */
static inline int fp_geq_p(const fp* a) {
  for (int i = 5; i >= 0; i--) {
    if (a->l[i] > P.l[i]) return 1;
    if (a->l[i] < P.l[i]) return 0;
  }
  return 1;
}
static inline void fp_sub_p(fp* a) {
  u128 b = 0;
  for (int i = 0; i < 6; i++) {
    u128 t = (u128)a->l[i] - P.l[i] - (uint64_t)b;
    a->l[i] = (uint64_t)t;
    b = (t >> 64) & 1;
  }
}
static inline void fp_add(fp* r, const fp* a, const fp* b) {
  u128 c = 0;
  for (int i = 0; i < 6; i++) {
    c += (u128)a->l[i] + b->l[i];
    r->l[i] = (uint64_t)c;
    c >>= 64;
  }
  if (fp_geq_p(r)) fp_sub_p(r);            /* p < 2^381, no carry out of limb 5 */
}
static inline void fp_sub(fp* r, const fp* a, const fp* b) {
  u128 bw = 0;
  for (int i = 0; i < 6; i++) {
    u128 t = (u128)a->l[i] - b->l[i] - (uint64_t)bw;
    r->l[i] = (uint64_t)t;
    bw = (t >> 64) & 1;
  }
  if (bw) {
    u128 c = 0;
    for (int i = 0; i < 6; i++) {
      c += (u128)r->l[i] + P.l[i];
      r->l[i] = (uint64_t)c;
      c >>= 64;
    }
  }
}
static inline int fp_is_zero(const fp* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3] | a->l[4] | a->l[5]) == 0; }
static inline void fp_neg(fp* r, const fp* a) {
  if (fp_is_zero(a)) { *r = *a; return; }
  fp z;
  memset(&z, 0, sizeof z);
  fp_sub(r, &z, a);
}
static inline void fp_dbl(fp* r, const fp* a) { fp_add(r, a, a); }

/* CIOS Montgomery multiplication, 6 x 64-bit limbs (ark-ff 0.4 Fp384 `mul_assign` semantics),
   written with scalar accumulators so the compiler keeps everything in registers (mulx/adcx). */
#define B381_ROW(bi)                                                                   \
  {                                                                                    \
    uint64_t c_, m_;                                                                   \
    u128 z_;                                                                           \
    z_ = (u128)a0 * (bi) + t0; t0 = (uint64_t)z_; c_ = (uint64_t)(z_ >> 64);           \
    z_ = (u128)a1 * (bi) + t1 + c_; t1 = (uint64_t)z_; c_ = (uint64_t)(z_ >> 64);      \
    z_ = (u128)a2 * (bi) + t2 + c_; t2 = (uint64_t)z_; c_ = (uint64_t)(z_ >> 64);      \
    z_ = (u128)a3 * (bi) + t3 + c_; t3 = (uint64_t)z_; c_ = (uint64_t)(z_ >> 64);      \
    z_ = (u128)a4 * (bi) + t4 + c_; t4 = (uint64_t)z_; c_ = (uint64_t)(z_ >> 64);      \
    z_ = (u128)a5 * (bi) + t5 + c_; t5 = (uint64_t)z_; c_ = (uint64_t)(z_ >> 64);      \
    t6 += c_;                    /* p < 2^381: the running value stays below 2^448 */  \
    m_ = t0 * N0;                                                                      \
    z_ = (u128)m_ * P.l[0] + t0; c_ = (uint64_t)(z_ >> 64);                            \
    z_ = (u128)m_ * P.l[1] + t1 + c_; t0 = (uint64_t)z_; c_ = (uint64_t)(z_ >> 64);    \
    z_ = (u128)m_ * P.l[2] + t2 + c_; t1 = (uint64_t)z_; c_ = (uint64_t)(z_ >> 64);    \
    z_ = (u128)m_ * P.l[3] + t3 + c_; t2 = (uint64_t)z_; c_ = (uint64_t)(z_ >> 64);    \
    z_ = (u128)m_ * P.l[4] + t4 + c_; t3 = (uint64_t)z_; c_ = (uint64_t)(z_ >> 64);    \
    z_ = (u128)m_ * P.l[5] + t5 + c_; t4 = (uint64_t)z_; c_ = (uint64_t)(z_ >> 64);    \
    t5 = t6 + c_;                                                                      \
    t6 = 0;                                                                            \
  }
static void fp_mul(fp* r, const fp* a, const fp* b) {
  const uint64_t a0 = a->l[0], a1 = a->l[1], a2 = a->l[2], a3 = a->l[3], a4 = a->l[4], a5 = a->l[5];
  const uint64_t b0 = b->l[0], b1 = b->l[1], b2 = b->l[2], b3 = b->l[3], b4 = b->l[4], b5 = b->l[5];
  uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0, t6 = 0;
  B381_ROW(b0) B381_ROW(b1) B381_ROW(b2) B381_ROW(b3) B381_ROW(b4) B381_ROW(b5)
  r->l[0] = t0; r->l[1] = t1; r->l[2] = t2; r->l[3] = t3; r->l[4] = t4; r->l[5] = t5;
  if (fp_geq_p(r)) fp_sub_p(r);
}
static inline void fp_sqr(fp* r, const fp* a) { fp_mul(r, a, a); }

/* a^(p-2) */
static void fp_inv(fp* r, const fp* a) {
  fp e = P, x = ONE;
  e.l[0] -= 2;
  for (int i = 5; i >= 0; i--)
    for (int b = 63; b >= 0; b--) {
      fp_sqr(&x, &x);
      if ((e.l[i] >> b) & 1) fp_mul(&x, &x, a);
    }
  *r = x;
}

/* ---- Fp2: src/fields_as_trees/fq2_target_tree.rs:66-142 ------------------------------------------ */
static inline void f2_add(fp2* r, const fp2* a, const fp2* b) { fp_add(&r->c0, &a->c0, &b->c0); fp_add(&r->c1, &a->c1, &b->c1); }
static inline void f2_sub(fp2* r, const fp2* a, const fp2* b) { fp_sub(&r->c0, &a->c0, &b->c0); fp_sub(&r->c1, &a->c1, &b->c1); }
static inline void f2_neg(fp2* r, const fp2* a) { fp_neg(&r->c0, &a->c0); fp_neg(&r->c1, &a->c1); }
static inline void f2_dbl(fp2* r, const fp2* a) { fp_dbl(&r->c0, &a->c0); fp_dbl(&r->c1, &a->c1); }
static void f2_mul(fp2* r, const fp2* a, const fp2* b) {           /* :97-115, Karatsuba */
  fp t0, t1, s0, s1, m;
  fp_mul(&t0, &a->c0, &b->c0);
  fp_mul(&t1, &a->c1, &b->c1);
  fp_add(&s0, &a->c0, &a->c1);
  fp_add(&s1, &b->c0, &b->c1);
  fp_mul(&m, &s0, &s1);
  fp_sub(&r->c0, &t0, &t1);
  fp_sub(&m, &m, &t0);
  fp_sub(&r->c1, &m, &t1);
}
static void f2_sqr(fp2* r, const fp2* a) {                         /* :80-91 */
  fp s, d, m;
  fp_add(&s, &a->c0, &a->c1);
  fp_sub(&d, &a->c0, &a->c1);
  fp_mul(&m, &a->c0, &a->c1);
  fp_mul(&r->c0, &s, &d);
  fp_dbl(&r->c1, &m);
}
static inline void f2_mul_fp(fp2* r, const fp2* a, const fp* s) { fp_mul(&r->c0, &a->c0, s); fp_mul(&r->c1, &a->c1, s); }
static inline void f2_mul_xi(fp2* r, const fp2* a) {               /* :137-142 */
  fp t0, t1;
  fp_sub(&t0, &a->c0, &a->c1);
  fp_add(&t1, &a->c0, &a->c1);
  r->c0 = t0; r->c1 = t1;
}
static inline void f2_conj(fp2* r, const fp2* a) { r->c0 = a->c0; fp_neg(&r->c1, &a->c1); }
static void f2_inv(fp2* r, const fp2* a) {                         /* :66-78 */
  fp n, t;
  fp_sqr(&n, &a->c0);
  fp_sqr(&t, &a->c1);
  fp_add(&n, &n, &t);
  fp_inv(&n, &n);
  fp_mul(&r->c0, &a->c0, &n);
  fp_mul(&t, &a->c1, &n);
  fp_neg(&r->c1, &t);
}

/* ---- Fp6: src/fields_as_trees/fq6_target_tree.rs:59-293 ------------------------------------------ */
static inline void f6_add(fp6* r, const fp6* a, const fp6* b) { f2_add(&r->c0, &a->c0, &b->c0); f2_add(&r->c1, &a->c1, &b->c1); f2_add(&r->c2, &a->c2, &b->c2); }
static inline void f6_sub(fp6* r, const fp6* a, const fp6* b) { f2_sub(&r->c0, &a->c0, &b->c0); f2_sub(&r->c1, &a->c1, &b->c1); f2_sub(&r->c2, &a->c2, &b->c2); }
static inline void f6_neg(fp6* r, const fp6* a) { f2_neg(&r->c0, &a->c0); f2_neg(&r->c1, &a->c1); f2_neg(&r->c2, &a->c2); }
static void f6_mul(fp6* r, const fp6* a, const fp6* b) {           /* :172-214 */
  fp2 aa, bb, cc, t, s0, s1, c0, c1, c2;
  f2_mul(&aa, &a->c0, &b->c0);
  f2_mul(&bb, &a->c1, &b->c1);
  f2_mul(&cc, &a->c2, &b->c2);
  f2_add(&s0, &a->c1, &a->c2); f2_add(&s1, &b->c1, &b->c2); f2_mul(&t, &s0, &s1);
  f2_sub(&t, &t, &bb); f2_sub(&t, &t, &cc); f2_mul_xi(&t, &t); f2_add(&c0, &t, &aa);
  f2_add(&s0, &a->c0, &a->c1); f2_add(&s1, &b->c0, &b->c1); f2_mul(&t, &s0, &s1);
  f2_sub(&t, &t, &aa); f2_sub(&t, &t, &bb); f2_mul_xi(&s0, &cc); f2_add(&c1, &t, &s0);
  f2_add(&s0, &a->c0, &a->c2); f2_add(&s1, &b->c0, &b->c2); f2_mul(&t, &s0, &s1);
  f2_sub(&t, &t, &aa); f2_add(&t, &t, &bb); f2_sub(&c2, &t, &cc);
  r->c0 = c0; r->c1 = c1; r->c2 = c2;
}
static inline void f6_sqr(fp6* r, const fp6* a) { f6_mul(r, a, a); }
static inline void f6_mul_by_v(fp6* r, const fp6* a) {             /* :219-230 */
  fp2 t;
  f2_mul_xi(&t, &a->c2);
  fp2 a0 = a->c0, a1 = a->c1;
  r->c0 = t; r->c1 = a0; r->c2 = a1;
}
static void f6_mul_by_01(fp6* r, const fp6* a, const fp2* c0, const fp2* c1) {   /* :232-259 */
  fp2 a_a, b_b, t1, t2, t3, s, u;
  f2_mul(&a_a, &a->c0, c0);
  f2_mul(&b_b, &a->c1, c1);
  f2_add(&s, &a->c1, &a->c2); f2_mul(&t1, c1, &s); f2_sub(&t1, &t1, &b_b); f2_mul_xi(&t1, &t1); f2_add(&t1, &t1, &a_a);
  f2_add(&s, c0, c1); f2_add(&u, &a->c0, &a->c1); f2_mul(&t3, &s, &u); f2_sub(&t3, &t3, &a_a); f2_sub(&t3, &t3, &b_b);
  f2_add(&s, &a->c0, &a->c2); f2_mul(&t2, c0, &s); f2_sub(&t2, &t2, &a_a); f2_add(&t2, &t2, &b_b);
  r->c0 = t1; r->c1 = t3; r->c2 = t2;
}
static void f6_mul_by_1(fp6* r, const fp6* a, const fp2* c1) {     /* :261-268 */
  fp2 t0, t1, t2;
  f2_mul(&t0, &a->c2, c1); f2_mul_xi(&t0, &t0);
  f2_mul(&t1, &a->c0, c1);
  f2_mul(&t2, &a->c1, c1);
  r->c0 = t0; r->c1 = t1; r->c2 = t2;
}
static void f6_inv(fp6* r, const fp6* a) {                         /* :59-89 */
  fp2 c0, c1, c2, t, u;
  f2_sqr(&c0, &a->c0); f2_mul(&t, &a->c1, &a->c2); f2_mul_xi(&t, &t); f2_sub(&c0, &c0, &t);
  f2_sqr(&c1, &a->c2); f2_mul_xi(&c1, &c1); f2_mul(&t, &a->c0, &a->c1); f2_sub(&c1, &c1, &t);
  f2_sqr(&c2, &a->c1); f2_mul(&t, &a->c0, &a->c2); f2_sub(&c2, &c2, &t);
  f2_mul(&t, &a->c2, &c1); f2_mul(&u, &a->c1, &c2); f2_add(&t, &t, &u); f2_mul_xi(&t, &t);
  f2_mul(&u, &a->c0, &c0); f2_add(&t, &t, &u);
  f2_inv(&t, &t);
  f2_mul(&r->c0, &t, &c0); f2_mul(&r->c1, &t, &c1); f2_mul(&r->c2, &t, &c2);
}

/* ---- Fp12: src/fields_as_trees/fq12_target_tree.rs:53-176 ---------------------------------------- */
static void f12_mul(fp12* r, const fp12* a, const fp12* b) {       /* :130-141 */
  fp6 aa, bb, s0, s1, t;
  f6_mul(&aa, &a->c0, &b->c0);
  f6_mul(&bb, &a->c1, &b->c1);
  f6_add(&s0, &a->c0, &a->c1);
  f6_add(&s1, &b->c0, &b->c1);
  f6_mul(&t, &s0, &s1);
  f6_sub(&t, &t, &aa);
  f6_sub(&r->c1, &t, &bb);
  f6_mul_by_v(&bb, &bb);
  f6_add(&r->c0, &bb, &aa);
}
static void f12_sqr(fp12* r, const fp12* a) {                      /* :143-155 */
  fp6 ab, s, u, t;
  f6_mul(&ab, &a->c0, &a->c1);
  f6_add(&s, &a->c0, &a->c1);
  f6_mul_by_v(&u, &a->c1);
  f6_add(&u, &u, &a->c0);
  f6_mul(&t, &u, &s);
  f6_sub(&t, &t, &ab);
  f6_add(&r->c1, &ab, &ab);
  f6_mul_by_v(&ab, &ab);
  f6_sub(&r->c0, &t, &ab);
}
static inline void f12_conj(fp12* r, const fp12* a) { r->c0 = a->c0; f6_neg(&r->c1, &a->c1); }
static void f12_inv(fp12* r, const fp12* a) {                      /* :77-90 */
  fp6 t, u;
  f6_sqr(&t, &a->c0);
  f6_sqr(&u, &a->c1);
  f6_mul_by_v(&u, &u);
  f6_sub(&t, &t, &u);
  f6_inv(&t, &t);
  f6_mul(&r->c0, &a->c0, &t);
  f6_mul(&u, &a->c1, &t);
  f6_neg(&r->c1, &u);
}
static void f12_mul_by_014(fp12* f, const fp2* c0, const fp2* c1, const fp2* c4) {   /* :157-176; src/miller_loop_native.rs:118-137 */
  fp6 aa, bb, s;
  fp2 o;
  f6_mul_by_01(&aa, &f->c0, c0, c1);
  f6_mul_by_1(&bb, &f->c1, c4);
  f2_add(&o, c1, c4);
  f6_add(&s, &f->c1, &f->c0);
  f6_mul_by_01(&s, &s, c0, &o);
  f6_sub(&s, &s, &aa);
  f6_sub(&f->c1, &s, &bb);
  f6_mul_by_v(&bb, &bb);
  f6_add(&f->c0, &bb, &aa);
}
static void f12_one(fp12* r) {
  memset(r, 0, sizeof *r);
  r->c0.c0.c0 = ONE;
}
static int f12_is_zero(const fp12* a) {
  const uint64_t* w = (const uint64_t*)a;
  uint64_t o = 0;
  for (int i = 0; i < 72; i++) o |= w[i];
  return o == 0;
}

/* Frobenius coefficients gamma_k[j] = xi^(j (p^k-1)/6), computed once (k = 1, 2) */
static fp2 GAMMA[3][6];
static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static void f2_pow_big(fp2* r, const fp2* a, const uint64_t* e, int nlimbs) {
  fp2 x;
  memset(&x, 0, sizeof x);
  x.c0 = ONE;
  for (int i = nlimbs - 1; i >= 0; i--)
    for (int b = 63; b >= 0; b--) {
      f2_sqr(&x, &x);
      if ((e[i] >> b) & 1) f2_mul(&x, &x, a);
    }
  *r = x;
}
static void init_gamma(void) {
  /* (p - 1)/6 as limbs; gamma_1[1] = xi^((p-1)/6); gamma_1[j] = gamma_1[1]^j;
     gamma_2[j] = gamma_1[j] * conj(gamma_1[j]) (norm: xi^(j(p^2-1)/6) = g^(p+1) = g * g^p, g^p = conj(g^...)) */
  uint64_t e[6];
  u128 rem = 0;
  fp pm1 = P;
  pm1.l[0] -= 1;
  for (int i = 5; i >= 0; i--) {
    u128 cur = (rem << 64) | pm1.l[i];
    e[i] = (uint64_t)(cur / 6);
    rem = cur % 6;
  }
  fp2 xi;
  xi.c0 = ONE; xi.c1 = ONE;
  fp2 g1;
  f2_pow_big(&g1, &xi, e, 6);
  memset(GAMMA, 0, sizeof GAMMA);
  for (int k = 1; k <= 2; k++) { GAMMA[k][0].c0 = ONE; }
  GAMMA[1][1] = g1;
  for (int j = 2; j < 6; j++) f2_mul(&GAMMA[1][j], &GAMMA[1][j - 1], &g1);
  /* xi^(j(p^2-1)/6) = (xi^(j(p-1)/6))^(p+1) = g_j^p * g_j ; g_j^p = conj(g_j) (Frobenius on Fp2) */
  for (int j = 1; j < 6; j++) {
    fp2 c;
    f2_conj(&c, &GAMMA[1][j]);
    f2_mul(&GAMMA[2][j], &c, &GAMMA[1][j]);
  }
}
static void f12_frobenius(fp12* r, const fp12* a, int k) {         /* :92-128; fq6_target_tree.rs:129-169 */
  pthread_once(&g_once, init_gamma);
  const fp2* in[6] = {&a->c0.c0, &a->c1.c0, &a->c0.c1, &a->c1.c1, &a->c0.c2, &a->c1.c2};   /* coefficient of w^t */
  fp2 out[6];
  for (int t = 0; t < 6; t++) {
    fp2 c = *in[t];
    if (k & 1) f2_conj(&c, &c);
    if (t == 0) out[t] = c; else f2_mul(&out[t], &c, &GAMMA[k][t]);
  }
  r->c0.c0 = out[0]; r->c1.c0 = out[1]; r->c0.c1 = out[2]; r->c1.c1 = out[3]; r->c0.c2 = out[4]; r->c1.c2 = out[5];
}

/* Granger-Scott cyclotomic squaring: src/fields_as_trees/miller_loop.rs:29-104 */
static void fp4_square(fp2* c0, fp2* c1, const fp2* a, const fp2* b) {
  fp2 t0, t1, s;
  f2_sqr(&t0, a);
  f2_sqr(&t1, b);
  f2_add(&s, a, b);
  f2_sqr(&s, &s);
  f2_sub(&s, &s, &t0);
  f2_sub(c1, &s, &t1);
  f2_mul_xi(&t1, &t1);
  f2_add(c0, &t1, &t0);
}
static void f12_cyclotomic_square(fp12* r, const fp12* f) {
  fp2 z0 = f->c0.c0, z4 = f->c0.c1, z3 = f->c0.c2, z2 = f->c1.c0, z1 = f->c1.c1, z5 = f->c1.c2;
  fp2 t0, t1, t2, t3, u;
  fp4_square(&t0, &t1, &z0, &z1);
  f2_sub(&u, &t0, &z0); f2_dbl(&u, &u); f2_add(&z0, &u, &t0);
  f2_add(&u, &t1, &z1); f2_dbl(&u, &u); f2_add(&z1, &u, &t1);
  fp4_square(&t0, &t1, &z2, &z3);
  fp4_square(&t2, &t3, &z4, &z5);
  f2_sub(&u, &t0, &z4); f2_dbl(&u, &u); f2_add(&z4, &u, &t0);
  f2_add(&u, &t1, &z5); f2_dbl(&u, &u); f2_add(&z5, &u, &t1);
  f2_mul_xi(&t0, &t3);
  f2_add(&u, &t0, &z2); f2_dbl(&u, &u); f2_add(&z2, &u, &t0);
  f2_sub(&u, &t2, &z3); f2_dbl(&u, &u); f2_add(&z3, &u, &t2);
  r->c0.c0 = z0; r->c0.c1 = z4; r->c0.c2 = z3; r->c1.c0 = z2; r->c1.c1 = z1; r->c1.c2 = z5;
}
static void f12_exp_by_x(fp12* r, const fp12* a) {                 /* ark Bls12::exp_by_x, x < 0 */
  fp12 t = *a;
  for (int b = 62; b >= 0; b--) {
    f12_cyclotomic_square(&t, &t);
    if ((X_ABS >> b) & 1) f12_mul(&t, &t, a);
  }
  f12_conj(r, &t);
}

/* ---- ARK Miller loop (ark-ec 0.4 models/bls12/{g2,mod}.rs) -------------------------- */
static void fp_half(fp* r, const fp* a) {                          /* a * 2^-1 */
  u128 c = 0;
  uint64_t t[7] = {a->l[0], a->l[1], a->l[2], a->l[3], a->l[4], a->l[5], 0};
  if (t[0] & 1) {
    for (int i = 0; i < 6; i++) {
      c += (u128)t[i] + P.l[i];
      t[i] = (uint64_t)c;
      c >>= 64;
    }
    t[6] = (uint64_t)c;
  }
  for (int i = 0; i < 6; i++) r->l[i] = (t[i] >> 1) | (t[i + 1] << 63);
}
static inline void f2_half(fp2* r, const fp2* a) { fp_half(&r->c0, &a->c0); fp_half(&r->c1, &a->c1); }

typedef struct { fp2 x, y, z; } g2proj;

static void ark_double_step(g2proj* R, fp2 co[3]) {
  fp2 a, b, c, e, f, g, h, i, j, e2, t;
  f2_mul(&a, &R->x, &R->y); f2_half(&a, &a);
  f2_sqr(&b, &R->y);
  f2_sqr(&c, &R->z);
  f2_dbl(&t, &c); f2_add(&t, &t, &c);                 /* 3c */
  f2_dbl(&t, &t); f2_dbl(&t, &t); f2_mul_xi(&e, &t);  /* e = (4,4) * 3c = 4 xi 3c */
  f2_dbl(&f, &e); f2_add(&f, &f, &e);
  f2_add(&g, &b, &f); f2_half(&g, &g);
  f2_add(&t, &R->y, &R->z); f2_sqr(&h, &t); f2_add(&t, &b, &c); f2_sub(&h, &h, &t);
  f2_sub(&i, &e, &b);
  f2_sqr(&j, &R->x);
  f2_sqr(&e2, &e);
  f2_sub(&t, &b, &f); f2_mul(&R->x, &a, &t);
  f2_sqr(&g, &g); f2_dbl(&t, &e2); f2_add(&t, &t, &e2); f2_sub(&R->y, &g, &t);
  f2_mul(&R->z, &b, &h);
  co[0] = i;
  f2_dbl(&t, &j); f2_add(&co[1], &t, &j);
  f2_neg(&co[2], &h);
}
static void ark_add_step(g2proj* R, const fp2* qx, const fp2* qy, fp2 co[3]) {
  fp2 th, la, c, d, e, f, g, h, t, u;
  f2_mul(&t, qy, &R->z); f2_sub(&th, &R->y, &t);
  f2_mul(&t, qx, &R->z); f2_sub(&la, &R->x, &t);
  f2_sqr(&c, &th);
  f2_sqr(&d, &la);
  f2_mul(&e, &la, &d);
  f2_mul(&f, &R->z, &c);
  f2_mul(&g, &R->x, &d);
  f2_add(&h, &e, &f); f2_dbl(&t, &g); f2_sub(&h, &h, &t);
  f2_mul(&R->x, &la, &h);
  f2_sub(&t, &g, &h); f2_mul(&t, &th, &t); f2_mul(&u, &e, &R->y); f2_sub(&R->y, &t, &u);
  f2_mul(&R->z, &R->z, &e);
  f2_mul(&t, &th, qx); f2_mul(&u, &la, qy); f2_sub(&co[0], &t, &u);
  f2_neg(&co[1], &th);
  co[2] = la;
}
static void ark_ell(fp12* f, const fp2 co[3], const fp* px, const fp* py) {
  fp2 c1, c2;
  f2_mul_fp(&c2, &co[2], py);
  f2_mul_fp(&c1, &co[1], px);
  f12_mul_by_014(f, &co[0], &c1, &c2);
}
static void ark_miller_loop(fp12* f, const fp* px, const fp* py, const fp2* qx, const fp2* qy) {
  g2proj R;
  fp2 co[3];
  R.x = *qx; R.y = *qy;
  memset(&R.z, 0, sizeof R.z);
  R.z.c0 = ONE;
  f12_one(f);
  for (int b = 62; b >= 0; b--) {
    f12_sqr(f, f);
    ark_double_step(&R, co);
    ark_ell(f, co, px, py);
    if ((X_ABS >> b) & 1) {
      ark_add_step(&R, qx, qy, co);
      ark_ell(f, co, px, py);
    }
  }
  f12_conj(f, f);
}
static int ark_final_exponentiation(fp12* out, const fp12* f) {    /* SURVEY A.5 */
  if (f12_is_zero(f)) return -1;
  fp12 f1, f2, r, y0, y1, y2;
  f12_conj(&f1, f);
  f12_inv(&f2, f);
  f12_mul(&r, &f1, &f2);
  f2 = r;
  f12_frobenius(&r, &r, 2);
  f12_mul(&r, &r, &f2);
  f12_cyclotomic_square(&y0, &r);
  f12_exp_by_x(&y1, &r);
  f12_conj(&y2, &r);
  f12_mul(&y1, &y1, &y2);
  f12_exp_by_x(&y2, &y1);
  f12_conj(&y1, &y1);
  f12_mul(&y1, &y1, &y2);
  f12_exp_by_x(&y2, &y1);
  f12_frobenius(&y1, &y1, 1);
  f12_mul(&y1, &y1, &y2);
  f12_mul(&r, &r, &y0);
  f12_exp_by_x(&y0, &y1);
  f12_exp_by_x(&y2, &y0);
  f12_frobenius(&y0, &y1, 2);
  f12_conj(&y1, &y1);
  f12_mul(&y1, &y1, &y2);
  f12_mul(&y1, &y1, &y0);
  f12_mul(out, &r, &y1);
  return 0;
}

/* ---- batch drivers (pthreads over the batch: the "rayon over the batch" stand-in) ------------------- */
enum { OP_MILLER = 0, OP_PAIRING = 1, OP_FINAL_EXP = 2, OP_FP_MUL = 3, OP_FP12_MUL = 4 };
typedef struct {
  int op;
  const uint32_t *a, *b;
  const uint8_t* inf;
  uint32_t* out;
  size_t lo, hi;
  int err;
} job_t;

static void do_pair(const uint32_t* g1, const uint32_t* g2, int inf, uint32_t* out, int with_fe, int* err) {
  fp12 f;
  if (inf & 3) {
    f12_one(&f);
  } else {
    fp px, py;
    fp2 qx, qy;
    memcpy(&px, g1, 48); memcpy(&py, g1 + 12, 48);
    memcpy(&qx, g2, 96); memcpy(&qy, g2 + 24, 96);
    ark_miller_loop(&f, &px, &py, &qx, &qy);
  }
  if (with_fe) {
    fp12 e;
    if (ark_final_exponentiation(&e, &f)) { *err = 1; memset(&e, 0, sizeof e); }
    f = e;
  }
  memcpy(out, &f, 576);
}

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  for (size_t i = j->lo; i < j->hi; i++) {
    switch (j->op) {
      case OP_MILLER: do_pair(j->a + 24 * i, j->b + 48 * i, j->inf ? j->inf[i] : 0, j->out + 144 * i, 0, &j->err); break;
      case OP_PAIRING: do_pair(j->a + 24 * i, j->b + 48 * i, j->inf ? j->inf[i] : 0, j->out + 144 * i, 1, &j->err); break;
      case OP_FINAL_EXP: {
        fp12 f, e;
        memcpy(&f, j->a + 144 * i, 576);
        if (ark_final_exponentiation(&e, &f)) { j->err = 1; memset(&e, 0, sizeof e); }
        memcpy(j->out + 144 * i, &e, 576);
      } break;
      case OP_FP_MUL: {
        fp x, y, r;
        memcpy(&x, j->a + 12 * i, 48); memcpy(&y, j->b + 12 * i, 48);
        fp_mul(&r, &x, &y);
        memcpy(j->out + 12 * i, &r, 48);
      } break;
      case OP_FP12_MUL: {
        fp12 x, y, r;
        memcpy(&x, j->a + 144 * i, 576); memcpy(&y, j->b + 144 * i, 576);
        f12_mul(&r, &x, &y);
        memcpy(j->out + 144 * i, &r, 576);
      } break;
    }
  }
  return NULL;
}

static int run_batch(int op, const uint32_t* a, const uint32_t* b, const uint8_t* inf, uint32_t* out, size_t n, int nthreads) {
  pthread_once(&g_once, init_gamma);
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > n) nthreads = (int)n;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  job_t* jobs = (job_t*)malloc(sizeof(job_t) * nthreads);
  for (int t = 0; t < nthreads; t++) {
    jobs[t].op = op; jobs[t].a = a; jobs[t].b = b; jobs[t].inf = inf; jobs[t].out = out; jobs[t].err = 0;
    jobs[t].lo = n * t / nthreads; jobs[t].hi = n * (t + 1) / nthreads;
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  int err = 0;
  for (int t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); err |= jobs[t].err; }
  free(th); free(jobs);
  return err;
}

int ref_miller_loop(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out, size_t n, int nthreads) { return run_batch(OP_MILLER, g1, g2, inf, out, n, nthreads); }
int ref_pairing(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out, size_t n, int nthreads) { return run_batch(OP_PAIRING, g1, g2, inf, out, n, nthreads); }
int ref_final_exp(const uint32_t* f, uint32_t* out, size_t n, int nthreads) { return run_batch(OP_FINAL_EXP, f, NULL, NULL, out, n, nthreads); }
int ref_fp_mul(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, int nthreads) { return run_batch(OP_FP_MUL, a, b, NULL, out, n, nthreads); }
int ref_fp12_mul(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, int nthreads) { return run_batch(OP_FP12_MUL, a, b, NULL, out, n, nthreads); }

/* multi_miller_loop: product of the Miller values (ark semantics), single-threaded product */
int ref_multi_miller_loop(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out144, size_t n, int nthreads) {
  uint32_t* tmp = (uint32_t*)malloc(n * 576);
  int err = run_batch(OP_MILLER, g1, g2, inf, tmp, n, nthreads);
  fp12 acc, x;
  f12_one(&acc);
  for (size_t i = 0; i < n; i++) {
    memcpy(&x, tmp + 144 * i, 576);
    f12_mul(&acc, &acc, &x);
  }
  memcpy(out144, &acc, 576);
  free(tmp);
  return err;
}
