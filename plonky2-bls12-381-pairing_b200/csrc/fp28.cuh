// fp28.cuh -- BLS12-381 base-field arithmetic for sm_100a, carry-free formulation.
//
// Why this format.  On B200 the integer multiplier issues IMAD / IMAD.WIDE at 64 lanes/clk/SM,
// but the carry-chained form IMAD.WIDE.U32.X (what mad.lo.cc/madc.hi.cc compile to) only at
// ~31 lanes/clk/SM (profiles/imad_probe_r01.jsonl).  So instead of 12 saturated 32-bit limbs with
// hardware carry chains, an Fp element is held as 14 SIGNED limbs of 28 bits; products are
// accumulated column-wise into signed 64-bit accumulators with plain IMAD.WIDE (no carry in, no
// carry out), and carries are resolved by shifts on the ALU pipe, which co-issues with IMAD.
//
//   value  = sum_k l[k] * 2^(28k),  "loosely normalised": l[k] in [-2^6, 2^28 + 2^6], l[13] signed
//   Montgomery radix R' = 2^420 (15 reduction rows over 14-limb operands): for ANY two 14-limb
//   operands (|v| < 2^391 ~ 1260 p) the reduced product lies in (-eps, p + eps), so additions and
//   subtractions never need a modular correction ("lazy" linear ops), only a carry pass.
//
// Everything here is plain C++ (int64 += int32*int32 compiles to IMAD.WIDE), so the identical
// source is compiled for the host by tests/ (bit-exact simulation of the device arithmetic, with
// optional worst-case bound tracking under B381_TRACK_BOUNDS).
//
// External format at the C ABI (include/b381.h): 12 x u32 little-endian limbs, Montgomery R = 2^384
// (= ark_ff Fp384 in-memory layout used by /root/reference/src/fields/helpers.rs:8-11).
#pragma once
#include <stdint.h>
#include "b381_consts.h"

#if defined(__CUDACC__)
#define B381_HD __host__ __device__
#define B381_INL __forceinline__
#else
#define B381_HD
#define B381_INL inline __attribute__((always_inline))
#endif

#ifdef B381_TRACK_BOUNDS
#include <cassert>
#include <cstdio>
#include <cstdlib>
#define B381_TB(x) x
#define B381_CHECK(cond, msg) do { if (!(cond)) { fprintf(stderr, "bound violation: %s (%s:%d)\n", msg, __FILE__, __LINE__); abort(); } } while (0)
#else
#define B381_TB(x)
#define B381_CHECK(cond, msg)
#endif

#ifdef B381_TRACK_BOUNDS
#define B381_SETRANGE(r, lo, hi) do { (r).mag = ((lo) < 0 ? -(lo) : (lo)) > ((hi) < 0 ? -(hi) : (hi)) ? ((lo) < 0 ? -(lo) : (lo)) : ((hi) < 0 ? -(hi) : (hi)); (r).lb = 1.0; (r).nonneg = (lo) >= 0; } while (0)
#else
#define B381_SETRANGE(r, lo, hi)
#endif

namespace b381 {

typedef int32_t limb_t;
constexpr int NL = B381_NL;          // 14 limbs
constexpr int W = B381_W;            // 28 bits
constexpr int32_t MASK = B381_MASK;
constexpr int NROWS = 15;            // Montgomery rows  => R' = 2^420
constexpr int NCOL = NL + NROWS;     // 29 columns (27 product columns + carries)

// one Fp element in registers
struct Fp {
  int32_t l[NL];
#ifdef B381_TRACK_BOUNDS
  double mag;   // bound on |value| / p
  double lb;    // bound on max |limb| / 2^28
  bool nonneg;  // limbs 0..12 are known to be >= 0 (the sign lives in limb 13 only)
#endif
};

// column accumulator (signed 64-bit per column)
struct Acc {
  int64_t c[NCOL];
#ifdef B381_TRACK_BOUNDS
  double cb;    // bound on max |column| / 2^56
  double mag;   // bound on |value| / p^2
#endif
};

B381_HD B381_INL constexpr int32_t plimb(int j) {
  switch (j) {
    case 0: return B381_P0;   case 1: return B381_P1;   case 2: return B381_P2;   case 3: return B381_P3;
    case 4: return B381_P4;   case 5: return B381_P5;   case 6: return B381_P6;   case 7: return B381_P7;
    case 8: return B381_P8;   case 9: return B381_P9;   case 10: return B381_P10; case 11: return B381_P11;
    case 12: return B381_P12; default: return B381_P13;
  }
}

// ---------------------------------------------------------------------------------------------
// linear operations (lazy: no modular correction)
// ---------------------------------------------------------------------------------------------
B381_HD B381_INL void fp_zero(Fp& r) {
#pragma unroll
  for (int k = 0; k < NL; k++) r.l[k] = 0;
  B381_TB(r.mag = 0; r.lb = 0; r.nonneg = true;)
}

B381_HD B381_INL void fp_add(Fp& r, const Fp& a, const Fp& b) {
#pragma unroll
  for (int k = 0; k < NL; k++) r.l[k] = a.l[k] + b.l[k];
  B381_TB(r.mag = a.mag + b.mag; r.lb = a.lb + b.lb; r.nonneg = a.nonneg && b.nonneg;)
}

B381_HD B381_INL void fp_sub(Fp& r, const Fp& a, const Fp& b) {
#pragma unroll
  for (int k = 0; k < NL; k++) r.l[k] = a.l[k] - b.l[k];
  B381_TB(r.mag = a.mag + b.mag; r.lb = a.lb + b.lb; r.nonneg = false;)
}

B381_HD B381_INL void fp_neg(Fp& r, const Fp& a) {
#pragma unroll
  for (int k = 0; k < NL; k++) r.l[k] = -a.l[k];
  B381_TB(r.mag = a.mag; r.lb = a.lb; r.nonneg = false;)
}

B381_HD B381_INL void fp_dbl(Fp& r, const Fp& a) {
#pragma unroll
  for (int k = 0; k < NL; k++) r.l[k] = a.l[k] << 1;
  B381_TB(r.mag = 2 * a.mag; r.lb = 2 * a.lb; r.nonneg = a.nonneg;)
}

// carry pass used before every store: sequential and exact, limbs 0..12 end in [0, 2^28) and the
// sign of the value lives in limb 13 (so multiplications can treat limbs 0..12 as unsigned)
B381_HD B381_INL void fp_norm(Fp& a) {
  B381_CHECK(a.lb < 7.9, "fp_norm: limb overflow (int32)");
  B381_CHECK(a.mag < 1200.0, "fp_norm: value does not fit 14 limbs");
#pragma unroll
  for (int k = 0; k < NL - 1; k++) {
    int32_t c = a.l[k] >> W;
    a.l[k] &= MASK;
    a.l[k + 1] += c;
  }
  B381_TB(a.lb = 1.0; a.nonneg = true;)
}

// exact sequential carry pass: limbs 0..12 in [0, 2^28), l[13] carries the sign
B381_HD B381_INL void fp_carry_exact(Fp& a) {
#pragma unroll
  for (int k = 0; k < NL - 1; k++) {
    int32_t c = a.l[k] >> W;
    a.l[k] &= MASK;
    a.l[k + 1] += c;
  }
  B381_TB(a.lb = 1.0; a.nonneg = true;)
}

// weak reduction: subtract q*p with q ~ floor(v/p) estimated from the top limbs, carrying in 64 bits.
// Any input that fits the limbs (|v| < 1200 p, limbs < 2^31) comes out in (-0.01p, 1.01p) with
// exactly normalised limbs.  Needed wherever a value feeds back LINEARLY into itself (the -2z term
// of the cyclotomic squaring), since lazy additions alone would double its magnitude each round.
// 14 IMAD.WIDE + 1.
B381_HD B381_INL void fp_wreduce(Fp& a) {
  B381_CHECK(a.lb < 7.9, "fp_wreduce: limb overflow (int32)");
  B381_CHECK(a.mag < 1200.0, "fp_wreduce: value does not fit 14 limbs");
  const int32_t t = a.l[NL - 1] + (a.l[NL - 2] >> W);                  // ~ floor(v / 2^364)
  const int32_t q = (int32_t)(((int64_t)t * 40323) >> 32);              // ~ floor(t / (p >> 364)), p >> 364 = 106513.9
  int64_t c = 0;
#pragma unroll
  for (int k = 0; k < NL - 1; k++) {
    c += (int64_t)a.l[k] - (int64_t)q * (int64_t)plimb(k);
    a.l[k] = (int32_t)c & MASK;
    c >>= W;
  }
  c += (int64_t)a.l[NL - 1] - (int64_t)q * (int64_t)plimb(NL - 1);
  a.l[NL - 1] = (int32_t)c;
  B381_TB(a.mag = 1.01; a.lb = 1.0; a.nonneg = true;)
}

// exact halving mod p: (v + (v odd ? p : 0)) / 2.  Equals multiplication by 2^-1 (ark-ec g2.rs
// double_in_place's mul_assign_by_fp(two_inv)).
B381_HD B381_INL void fp_half(Fp& r, const Fp& a) {
  int32_t odd = -(a.l[0] & 1);      // 0 or -1 (parity of the value = parity of limb 0)
  Fp t;
#pragma unroll
  for (int k = 0; k < NL; k++) t.l[k] = a.l[k] + (plimb(k) & odd);
#pragma unroll
  for (int k = 0; k < NL - 1; k++) r.l[k] = (t.l[k] >> 1) + ((t.l[k + 1] & 1) << (W - 1));
  r.l[NL - 1] = t.l[NL - 1] >> 1;
  B381_CHECK(a.nonneg, "fp_half: limbs must be non-negative");
  B381_TB(r.mag = (a.mag + 1) / 2; r.lb = (a.lb + 1) / 2 + 0.5; r.nonneg = true;)
}

// ---------------------------------------------------------------------------------------------
// multiply-accumulate into columns, Montgomery reduction
// ---------------------------------------------------------------------------------------------
B381_HD B381_INL void acc_zero(Acc& t) {
#pragma unroll
  for (int k = 0; k < NCOL; k++) t.c[k] = 0;
  B381_TB(t.cb = 0; t.mag = 0;)
}

// Multiply-accumulate into a 64-bit column.  Stored values keep limbs 0..12 in [0, 2^28] and carry
// the sign in limb 13 only, so 169 of the 196 products of a 14 x 14 block are UNSIGNED: they compile
// to one IMAD.WIDE.U32 each (ptxas emulates signed 32x32->64 with an unsigned multiply plus a
// correction add, which doubled the ALU traffic of an all-signed formulation); only the 27 products
// touching a top limb are signed.  ptxas lowers every 64-bit multiply-accumulate to IMAD.WIDE (RZ
// addend) plus a shared IADD3 / IADD3.X pair per two products, i.e. one multiplier-pipe and one
// ALU-pipe instruction per MAC (both pipes issue 64 lanes/clk/SM), whatever the source form.
// B381_MAC_STYLE 1 emits volatile PTX in row-major order (isolated MAC blocks then reach 59
// IMAD/clk/SM instead of 30, tools/phase_probe.cu) but raises register pressure in the full
// primitives and is not faster end to end; 0 (default) leaves the order to ptxas.
#ifndef B381_MAC_STYLE
#define B381_MAC_STYLE 0
#endif
#if defined(__CUDA_ARCH__) && B381_MAC_STYLE == 1
#define B381_MACU(c, a, b) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c) : "r"(a), "r"(b))
#define B381_MACS(c, a, b) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(c) : "r"(a), "r"(b))
#define B381_MACI(c, a, imm) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c) : "r"(a), "n"(imm))
#else
#define B381_MACU(c, a, b) (c) = (int64_t)((uint64_t)(c) + (uint64_t)(uint32_t)(a) * (uint64_t)(uint32_t)(b))
#define B381_MACS(c, a, b) (c) += (int64_t)(int32_t)(a) * (int64_t)(int32_t)(b)
#define B381_MACI(c, a, imm) (c) = (int64_t)((uint64_t)(c) + (uint64_t)(uint32_t)(a) * (uint64_t)(uint32_t)(imm))
#endif
// NVVM widens the Montgomery quotient digit m = (c * n0') & MASK to 64-bit arithmetic (mul.lo.s64 +
// and.b64) and every m * p_k to a 64 x 64-bit mul.lo.s64; ptxas then emits, per reduction MAC, an
// IMAD.WIDE.U32 plus a dead IADD3 of the (zero) high-half cross product -- 16 % of all executed
// instructions of the pairing kernel.  Passing m through an empty asm as a 32-bit register keeps it a
// 32-bit value, its zero extension is then visible to ptxas and each MAC is a single IMAD.WIDE.U32
// with immediate multiplicand and fused 64-bit accumulate (f2_sqr: 1928 -> 1352 SASS instructions).
#if defined(__CUDA_ARCH__) && !defined(B381_NO_OPAQUE_M)
#define B381_OPAQUE32(m) asm("" : "+r"(m))
#else
#define B381_OPAQUE32(m)
#endif
// one Montgomery row: t.c[i .. i+13] += m * p
#define B381_DECL_P
#define B381_ROW_P_IMM(t, i, m)                                                                          \
  B381_MACI(t.c[(i) + 0], m, B381_P0);  B381_MACI(t.c[(i) + 1], m, B381_P1);  B381_MACI(t.c[(i) + 2], m, B381_P2);   \
  B381_MACI(t.c[(i) + 3], m, B381_P3);  B381_MACI(t.c[(i) + 4], m, B381_P4);  B381_MACI(t.c[(i) + 5], m, B381_P5);   \
  B381_MACI(t.c[(i) + 6], m, B381_P6);  B381_MACI(t.c[(i) + 7], m, B381_P7);  B381_MACI(t.c[(i) + 8], m, B381_P8);   \
  B381_MACI(t.c[(i) + 9], m, B381_P9);  B381_MACI(t.c[(i) + 10], m, B381_P10); B381_MACI(t.c[(i) + 11], m, B381_P11); \
  B381_MACI(t.c[(i) + 12], m, B381_P12); B381_MACI(t.c[(i) + 13], m, B381_P13)
#define B381_ROW_P(t, i, m) B381_ROW_P_IMM(t, i, m)

// t += a * b   (196 IMAD.WIDE, no carries)
B381_HD B381_INL void acc_mac(Acc& t, const Fp& a, const Fp& b) {
#pragma unroll
  for (int i = 0; i < NL; i++)
#pragma unroll
    for (int j = 0; j < NL; j++) {
      if (i == NL - 1 || j == NL - 1) B381_MACS(t.c[i + j], a.l[i], b.l[j]);   // a top limb is signed
      else B381_MACU(t.c[i + j], a.l[i], b.l[j]);
    }
  B381_TB(t.cb += 14.0 * a.lb * b.lb; t.mag += a.mag * b.mag;)
  B381_CHECK(a.nonneg && b.nonneg, "acc_mac: operand limbs 0..12 must be non-negative");
  B381_CHECK(t.cb < 120.0, "acc_mac: column overflow");
}

// Same as acc_mac but without the conservative column bookkeeping: used only for the Karatsuba
// cross terms of a sum of products, where the bound is established algebraically by the caller
// (f2_sop in tower.cuh).
B381_HD B381_INL void acc_mac_cross(Acc& t, const Fp& a, const Fp& b) {
#pragma unroll
  for (int i = 0; i < NL; i++)
#pragma unroll
    for (int j = 0; j < NL; j++) {
      if (i == NL - 1 || j == NL - 1) B381_MACS(t.c[i + j], a.l[i], b.l[j]);
      else B381_MACU(t.c[i + j], a.l[i], b.l[j]);
    }
  B381_TB(t.mag += a.mag * b.mag;)
  B381_CHECK(a.nonneg && b.nonneg, "acc_mac_cross: operand limbs 0..12 must be non-negative");
}

B381_HD B381_INL void acc_add(Acc& r, const Acc& a, const Acc& b) {
#pragma unroll
  for (int k = 0; k < 2 * NL - 1; k++) r.c[k] = a.c[k] + b.c[k];
#pragma unroll
  for (int k = 2 * NL - 1; k < NCOL; k++) r.c[k] = 0;
  B381_TB(r.cb = a.cb + b.cb; r.mag = a.mag + b.mag;)
  B381_CHECK(r.cb < 120.0, "acc_add: column overflow");
}

B381_HD B381_INL void acc_sub(Acc& r, const Acc& a, const Acc& b) {
#pragma unroll
  for (int k = 0; k < 2 * NL - 1; k++) r.c[k] = a.c[k] - b.c[k];
#pragma unroll
  for (int k = 2 * NL - 1; k < NCOL; k++) r.c[k] = 0;
  B381_TB(r.cb = a.cb + b.cb; r.mag = a.mag + b.mag;)
  B381_CHECK(r.cb < 120.0, "acc_sub: column overflow");
}

B381_HD B381_INL void acc_neg(Acc& r, const Acc& a) {
#pragma unroll
  for (int k = 0; k < 2 * NL - 1; k++) r.c[k] = -a.c[k];
#pragma unroll
  for (int k = 2 * NL - 1; k < NCOL; k++) r.c[k] = 0;
  B381_TB(r.cb = a.cb; r.mag = a.mag;)
}

// r = t / 2^420 mod p, result in (-eps, p + eps), limbs exactly normalised (l[13] signed).
// 15 + 15*14 = 225 IMAD/IMAD.WIDE.
B381_HD B381_INL void acc_redc(Fp& r, Acc& t) {
  B381_CHECK(t.cb + 14.0 + 1.0 < 127.0, "acc_redc: column overflow");
  B381_CHECK(t.mag < 1.5e6, "acc_redc: input too large");
  B381_DECL_P;
#pragma unroll
  for (int i = 0; i < NROWS; i++) {
    uint32_t m = ((uint32_t)t.c[i] * (uint32_t)B381_N0P) & (uint32_t)MASK;
    B381_OPAQUE32(m);
    B381_ROW_P(t, i, m);
    t.c[i + 1] += t.c[i] >> W;
  }
#pragma unroll
  for (int k = 0; k < NL - 1; k++) {
    r.l[k] = (int32_t)t.c[NROWS + k] & MASK;
    t.c[NROWS + k + 1] += t.c[NROWS + k] >> W;
  }
  r.l[NL - 1] = (int32_t)t.c[NROWS + NL - 1];
  B381_TB(r.mag = 1.0 + t.mag / 1.0e11 + 1e-9; r.lb = 1.0; r.nonneg = true;)   // p / 2^420 * p^2 / p ~ 2^-39
}

// two independent reductions with their rows interleaved in source order: the latency of one row's
// m = c * n0' chain hides behind the other reduction's 14 multiply-accumulates.
B381_HD B381_INL void acc_redc2(Fp& r0, Acc& t0, Fp& r1, Acc& t1) {
  B381_CHECK(t0.cb + 14.0 + 1.0 < 127.0 && t1.cb + 14.0 + 1.0 < 127.0, "acc_redc2: column overflow");
  B381_CHECK(t0.mag < 1.5e6 && t1.mag < 1.5e6, "acc_redc2: input too large");
  B381_DECL_P;
#pragma unroll
  for (int i = 0; i < NROWS; i++) {
    uint32_t m0 = ((uint32_t)t0.c[i] * (uint32_t)B381_N0P) & (uint32_t)MASK;
    B381_OPAQUE32(m0);
    uint32_t m1 = ((uint32_t)t1.c[i] * (uint32_t)B381_N0P) & (uint32_t)MASK;
    B381_OPAQUE32(m1);
    B381_ROW_P(t0, i, m0);
    B381_ROW_P(t1, i, m1);
    t0.c[i + 1] += t0.c[i] >> W;
    t1.c[i + 1] += t1.c[i] >> W;
  }
#pragma unroll
  for (int k = 0; k < NL - 1; k++) {
    r0.l[k] = (int32_t)t0.c[NROWS + k] & MASK;
    r1.l[k] = (int32_t)t1.c[NROWS + k] & MASK;
    t0.c[NROWS + k + 1] += t0.c[NROWS + k] >> W;
    t1.c[NROWS + k + 1] += t1.c[NROWS + k] >> W;
  }
  r0.l[NL - 1] = (int32_t)t0.c[NROWS + NL - 1];
  r1.l[NL - 1] = (int32_t)t1.c[NROWS + NL - 1];
  B381_TB(r0.mag = 1.0 + t0.mag / 1.0e11 + 1e-9; r0.lb = 1.0; r1.mag = 1.0 + t1.mag / 1.0e11 + 1e-9; r1.lb = 1.0; r0.nonneg = r1.nonneg = true;)
}

// r = t / 2^384 mod p: Montgomery reduction in the EXTERNAL domain (R = 2^384), 13 full rows plus
// one 20-bit row.  Used by the element-wise Fp / Fp2 multiply entry points, which then need no
// domain conversion.  For canonical operands the result lies in (-p, 2p).
B381_HD B381_INL void acc_redc384(Fp& r, Acc& t) {
  B381_CHECK(t.cb + 14.0 + 1.0 < 127.0, "acc_redc384: column overflow");
  B381_CHECK(t.mag < 4.0, "acc_redc384: operands must be canonical");
  B381_DECL_P;
#pragma unroll
  for (int i = 0; i < NL - 1; i++) {
    uint32_t m = ((uint32_t)t.c[i] * (uint32_t)B381_N0P) & (uint32_t)MASK;
    B381_OPAQUE32(m);
    B381_ROW_P(t, i, m);
    t.c[i + 1] += t.c[i] >> W;
  }
  {
    uint32_t m = ((uint32_t)t.c[NL - 1] * (uint32_t)B381_N0P) & 0xfffffu;
    B381_OPAQUE32(m);
    B381_ROW_P(t, NL - 1, m);
  }
  int32_t d[NL + 1];
#pragma unroll
  for (int k = 0; k < NL; k++) {
    d[k] = (int32_t)t.c[NL - 1 + k] & MASK;
    t.c[NL + k] += t.c[NL - 1 + k] >> W;
  }
  d[NL] = (int32_t)t.c[2 * NL - 1];
#pragma unroll
  for (int k = 0; k < NL - 1; k++) r.l[k] = (d[k] >> 20) | ((d[k + 1] & 0xfffff) << 8);
  r.l[NL - 1] = (d[NL - 1] >> 20) + (d[NL] << 8);
  B381_TB(r.mag = 1.0 + t.mag / 9.8; r.lb = 1.0; r.nonneg = true;)
}

B381_HD B381_INL void fp_mul(Fp& r, const Fp& a, const Fp& b) {
  Acc t;
  acc_zero(t);
  acc_mac(t, a, b);
  acc_redc(r, t);
}

B381_HD B381_INL void fp_set(Fp& r, const int32_t (&v)[NL]) {
#pragma unroll
  for (int k = 0; k < NL; k++) r.l[k] = v[k];
  B381_TB(r.mag = 1.0; r.lb = 1.0; r.nonneg = true;)
}

// ---------------------------------------------------------------------------------------------
// canonical form, comparisons, external format (12 x u32, Montgomery R = 2^384)
// ---------------------------------------------------------------------------------------------
// bring a value in (-p, 2p) to [0, p), limbs exact
B381_HD B381_INL void fp_canon_small(Fp& a) {
  B381_CHECK(a.mag < 2.0, "fp_canon_small: input range");
  fp_carry_exact(a);
  int32_t neg = a.l[NL - 1] >> 31;                   // -1 if negative
#pragma unroll
  for (int k = 0; k < NL; k++) a.l[k] += plimb(k) & neg;
  fp_carry_exact(a);
  Fp t;
#pragma unroll
  for (int k = 0; k < NL; k++) t.l[k] = a.l[k] - plimb(k);
  fp_carry_exact(t);
  int32_t ge = ~(t.l[NL - 1] >> 31);                 // -1 if a >= p
#pragma unroll
  for (int k = 0; k < NL; k++) a.l[k] = (t.l[k] & ge) | (a.l[k] & ~ge);
  B381_TB(a.mag = 1.0; a.lb = 1.0; a.nonneg = true;)
}

// full reduction of any stored value to canonical [0,p): one Montgomery multiplication by R' mod p
B381_HD B381_INL void fp_canon(Fp& a) {
  const int32_t one[NL] = B381_ONE;
  Fp o;
  fp_set(o, one);
  Fp n = a;
  fp_norm(n);
  fp_mul(a, n, o);
  fp_canon_small(a);
}

B381_HD B381_INL bool fp_is_zero_canon(const Fp& a) {
  int32_t o = 0;
#pragma unroll
  for (int k = 0; k < NL; k++) o |= a.l[k];
  return o == 0;
}

B381_HD B381_INL bool fp_eq_canon(const Fp& a, const Fp& b) {
  int32_t o = 0;
#pragma unroll
  for (int k = 0; k < NL; k++) o |= a.l[k] ^ b.l[k];
  return o == 0;
}

// 12 x u32 (plain integer X < 2^384) -> 14 x 28-bit limbs
B381_HD B381_INL void fp_unpack32(Fp& r, const uint32_t (&w)[12]) {
#pragma unroll
  for (int k = 0; k < NL; k++) {
    const int bit = W * k, j = bit >> 5, s = bit & 31;
    uint32_t v = w[j] >> s;
    if (s > 32 - W && j + 1 < 12) v |= w[j + 1] << (32 - s);
    r.l[k] = (int32_t)(v & (uint32_t)MASK);
  }
  B381_TB(r.mag = 9.9; r.lb = 1.0; r.nonneg = true;)     // any 384-bit integer
}

// canonical limbs -> 12 x u32
B381_HD B381_INL void fp_pack32(uint32_t (&w)[12], const Fp& a) {
#pragma unroll
  for (int j = 0; j < 12; j++) {
    const int bit = 32 * j, k = bit / W, s = bit % W;
    uint32_t v = (uint32_t)a.l[k] >> s;
    if (k + 1 < NL) v |= (uint32_t)a.l[k + 1] << (W - s);
    if (W - s + W < 32 && k + 2 < NL) v |= (uint32_t)a.l[k + 2] << (2 * W - s);
    w[j] = v;
  }
}

// range check of an unpacked plain integer: X < p
B381_HD B381_INL bool fp_below_p(const Fp& x) {
  Fp t;
#pragma unroll
  for (int k = 0; k < NL; k++) t.l[k] = x.l[k] - plimb(k);
  B381_TB(t.mag = 11; t.lb = 2; t.nonneg = false;)
  fp_carry_exact(t);
  return (t.l[NL - 1] >> 31) != 0;
}

// X (12 x u32, Montgomery R = 2^384, canonical) -> internal.  Returns false if X >= p.
B381_HD B381_INL bool fp_from_ext(Fp& r, const uint32_t (&w)[12]) {
  Fp x;
  fp_unpack32(x, w);
  // range check X < p
  Fp t;
#pragma unroll
  for (int k = 0; k < NL; k++) t.l[k] = x.l[k] - plimb(k);
  B381_TB(t.mag = 11; t.lb = 2; t.nonneg = false;)
  fp_carry_exact(t);
  bool ok = (t.l[NL - 1] >> 31) != 0;
  const int32_t cin[NL] = B381_CIN;
  Fp c;
  fp_set(c, cin);
  fp_mul(r, x, c);
  return ok;
}

// internal -> 12 x u32, Montgomery R = 2^384, canonical
B381_HD B381_INL void fp_to_ext(uint32_t (&w)[12], const Fp& a) {
  const int32_t cout[NL] = B381_COUT;
  Fp c, n = a, o;
  fp_set(c, cout);
  fp_norm(n);
  fp_mul(o, n, c);
  fp_canon_small(o);
  fp_pack32(w, o);
}

}  // namespace b381
