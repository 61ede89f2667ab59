// fp32.cuh -- BLS12-381 base-field arithmetic for sm_100a, 13 x 32-bit two's-complement words with
// carry chains (B381_FMT == 32).  Same interface and the same lazy semantics as fp28.cuh, so
// tower.cuh / programs.cuh / kernels.cu are shared between the two formats.
//
// Why this format.  Measured on B200 (tools/imad_probe5.cu, profiles/imad_probe5_r01.jsonl): a warp-wide
// IMAD.WIDE occupies the multiplier pipe of its SM sub-partition for 4 cycles (32 thread-ops/clk/SM)
// whether or not it carries (IMAD.WIDE.U32.X: 31/clk/SM), and the pairing kernel already keeps that
// pipe 75 % busy.  What is left is the NUMBER of wide multiplies per field multiplication:
//   14 x 28-bit limbs, carry-free columns (fp28.cuh):  196 + 15*14+15 = 421 per Fp multiplication
//   13 x 32-bit words, carry chains (this file):       169 + 13*12    = 325 (+13 narrow IMAD)
//   ... with operands provably below 2^384 (acc_mul12): 144 + 13*12    = 300 = the algorithmic count
// and the accumulators shrink from 29 x 64-bit columns (58 registers) to 26 words.
//
//   value   = two's-complement integer of 13 x 32-bit words (sign in bit 415), |v| < 2^20 p by the
//             bound tracker; additions / subtractions / negations are plain carry chains (no modular
//             correction: "lazy"), products of two such values are 26-word two's-complement integers.
//             STORED values are kept non-negative and below 5 p (weak reduction / + p / + 5 p offsets),
//             so the hot path multiplies unsigned 12-word operands; the tracker (lower and upper
//             bound per value) asserts it at every multiplication and store.
//   radix   R' = 2^416: 13 Montgomery rows over the 12 words of p.  (T + M p) / 2^416 lies in
//             (T / 2^416, T / 2^416 + p]: any product of operands below 2^17 p reduces to (-eps, p + eps).
//
// Carry chains are PTX add.cc / madc.{lo,hi}.cc sequences (one asm statement per instruction, in the
// style of the public sppark / blst GPU field code; ptxas fuses each mad.lo.cc / madc.hi.cc pair into
// one IMAD.WIDE.U32(.X)).  On the host the same macros expand to uint64 arithmetic with an explicit
// carry variable, so tests/hostsim/ runs the identical algorithm bit for bit, with optional
// magnitude tracking under B381_TRACK_BOUNDS.
//
// External format at the C ABI (include/b381.h): 12 x u32 little-endian words, Montgomery R = 2^384
// (= ark_ff Fp384 in-memory layout used by /root/reference/src/fields/helpers.rs:8-11).
#pragma once
#include <stdint.h>
#include "b381_consts32.h"

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
#include <execinfo.h>
#define B381_TB(x) x
#define B381_CHECK(cond, msg) do { if (!(cond)) { fprintf(stderr, "bound violation: %s (%s:%d)\n", msg, __FILE__, __LINE__); void* bt_[24]; int n_ = backtrace(bt_, 24); backtrace_symbols_fd(bt_, n_, 2); abort(); } } while (0)
#else
#define B381_TB(x)
#define B381_CHECK(cond, msg)
#endif

// ---- carry-chain instruction macros ------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define B381_CC_DECL
#define ADD_CC(r, a, b)   asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b))
#define ADDC_CC(r, a, b)  asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b))
#define ADDC(r, a, b)     asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b))
#define SUB_CC(r, a, b)   asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b))
#define SUBC_CC(r, a, b)  asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b))
#define SUBC(r, a, b)     asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b))
#define MUL_WIDE(lo, hi, a, b) asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b))
#define MAD_LO_CC(r, a, b)   asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(r) : "r"(a), "r"(b))
#define MADC_LO_CC(r, a, b)  asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(r) : "r"(a), "r"(b))
#define MADC_HI_CC(r, a, b)  asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(r) : "r"(a), "r"(b))
#define MADC_HI(r, a, b)     asm volatile("madc.hi.u32 %0, %1, %2, %0;" : "+r"(r) : "r"(a), "r"(b))
#else
#define B381_CC_DECL uint32_t cc_ = 0; (void)cc_
#define ADD_CC(r, a, b)   do { uint64_t t_ = (uint64_t)(uint32_t)(a) + (uint32_t)(b); (r) = (uint32_t)t_; cc_ = (uint32_t)(t_ >> 32); } while (0)
#define ADDC_CC(r, a, b)  do { uint64_t t_ = (uint64_t)(uint32_t)(a) + (uint32_t)(b) + cc_; (r) = (uint32_t)t_; cc_ = (uint32_t)(t_ >> 32); } while (0)
#define ADDC(r, a, b)     do { (r) = (uint32_t)(a) + (uint32_t)(b) + cc_; } while (0)
#define SUB_CC(r, a, b)   do { uint64_t t_ = (uint64_t)(uint32_t)(a) - (uint32_t)(b); (r) = (uint32_t)t_; cc_ = (uint32_t)(t_ >> 63); } while (0)
#define SUBC_CC(r, a, b)  do { uint64_t t_ = (uint64_t)(uint32_t)(a) - (uint32_t)(b) - cc_; (r) = (uint32_t)t_; cc_ = (uint32_t)(t_ >> 63); } while (0)
#define SUBC(r, a, b)     do { (r) = (uint32_t)(a) - (uint32_t)(b) - cc_; } while (0)
#define MUL_WIDE(lo, hi, a, b) do { uint64_t t_ = (uint64_t)(uint32_t)(a) * (uint32_t)(b); (lo) = (uint32_t)t_; (hi) = (uint32_t)(t_ >> 32); } while (0)
#define MAD_LO_CC(r, a, b)   do { uint64_t t_ = (uint64_t)(uint32_t)((uint64_t)(uint32_t)(a) * (uint32_t)(b)) + (r); (r) = (uint32_t)t_; cc_ = (uint32_t)(t_ >> 32); } while (0)
#define MADC_LO_CC(r, a, b)  do { uint64_t t_ = (uint64_t)(uint32_t)((uint64_t)(uint32_t)(a) * (uint32_t)(b)) + (r) + cc_; (r) = (uint32_t)t_; cc_ = (uint32_t)(t_ >> 32); } while (0)
#define MADC_HI_CC(r, a, b)  do { uint64_t t_ = (((uint64_t)(uint32_t)(a) * (uint32_t)(b)) >> 32) + (r) + cc_; (r) = (uint32_t)t_; cc_ = (uint32_t)(t_ >> 32); } while (0)
#define MADC_HI(r, a, b)     do { (r) = (uint32_t)((((uint64_t)(uint32_t)(a) * (uint32_t)(b)) >> 32) + (r) + cc_); } while (0)
#endif

namespace b381 {

typedef uint32_t limb_t;
constexpr int NL = 13;               // 32-bit words per field element (two's complement)
constexpr int NW = 2 * NL;           // words of a double-width accumulator
constexpr int NPW = 12;              // words of p

struct Fp {
  uint32_t l[NL];
#ifdef B381_TRACK_BOUNDS
  double mag;   // bound on |value| / p
  double lb;    // LOWER bound on value / p (>= 0: provably non-negative)
  double ub;    // UPPER bound on value / p (mag = max(|lb|, |ub|))
  bool nonneg;  // unused in this format
#endif
};

struct Acc {
  uint32_t c[NW];
#ifdef B381_TRACK_BOUNDS
  double cb;    // LOWER bound on value / p^2
  double mag;   // bound on |value| / p^2
#endif
};

#ifdef B381_TRACK_BOUNDS
inline void tb_range(Fp& r, double lo, double hi) { r.lb = lo; r.ub = hi; r.mag = (lo < 0 ? -lo : lo) > (hi < 0 ? -hi : hi) ? (lo < 0 ? -lo : lo) : (hi < 0 ? -hi : hi); r.nonneg = true; }
#define B381_SETRANGE(r, lo, hi) tb_range(r, lo, hi)
#else
#define B381_SETRANGE(r, lo, hi)
#endif
constexpr double FP_MAG_MAX = 1048576.0;        // 2^20 p < 2^401: far inside the 2^415 word range
constexpr double MUL_MAG_MAX = 131072.0;        // operands of a multiplication: |v| < 2^17 p
constexpr double ACC_MAG_MAX = 1.0e9;           // reduction input: result < (1 + 1e9 / 4.2e10) p

B381_HD B381_INL constexpr uint32_t pword(int j) {
  switch (j) {
    case 0: return B381_Q0;   case 1: return B381_Q1;   case 2: return B381_Q2;   case 3: return B381_Q3;
    case 4: return B381_Q4;   case 5: return B381_Q5;   case 6: return B381_Q6;   case 7: return B381_Q7;
    case 8: return B381_Q8;   case 9: return B381_Q9;   case 10: return B381_Q10; case 11: return B381_Q11;
    default: return 0u;
  }
}

// ---------------------------------------------------------------------------------------------
// linear operations (lazy: no modular correction; carry chains are exact)
// ---------------------------------------------------------------------------------------------
B381_HD B381_INL void fp_zero(Fp& r) {
#pragma unroll
  for (int k = 0; k < NL; k++) r.l[k] = 0;
  B381_SETRANGE(r, 0, 0);
}

// r = a + 128 p: makes a difference of two stored values (each below 128 p) non-negative
B381_HD B381_INL void fp_add_p128(Fp& r, const Fp& a) {
  const uint32_t k[NL] = B381_P128;
  B381_CC_DECL;
  ADD_CC(r.l[0], a.l[0], k[0]);
#pragma unroll
  for (int i = 1; i < NL - 1; i++) ADDC_CC(r.l[i], a.l[i], k[i]);
  ADDC(r.l[NL - 1], a.l[NL - 1], k[NL - 1]);
  B381_SETRANGE(r, a.lb + 128, a.ub + 128);
}

// r = a + 5 p: a difference a0 - a1 of stored values (each at most 5 p) made non-negative while staying
// below 2^384 = 9.84 p, so that it still fits the 12-word multiplication
B381_HD B381_INL void fp_add_p5(Fp& r, const Fp& a) {
  const uint32_t k[NL] = B381_P5;
  B381_CC_DECL;
  ADD_CC(r.l[0], a.l[0], k[0]);
#pragma unroll
  for (int i = 1; i < NL - 1; i++) ADDC_CC(r.l[i], a.l[i], k[i]);
  ADDC(r.l[NL - 1], a.l[NL - 1], k[NL - 1]);
  B381_SETRANGE(r, a.lb + 5, a.ub + 5);
}

// r = a + 3 p: offset of the in-register differences b0 - b1 (xi b) and b - f of the fused primitives, whose
// subtrahend is at most 3 p
B381_HD B381_INL void fp_add_p3(Fp& r, const Fp& a) {
  const uint32_t k[NL] = B381_P3;
  B381_CC_DECL;
  ADD_CC(r.l[0], a.l[0], k[0]);
#pragma unroll
  for (int i = 1; i < NL - 1; i++) ADDC_CC(r.l[i], a.l[i], k[i]);
  ADDC(r.l[NL - 1], a.l[NL - 1], k[NL - 1]);
  B381_SETRANGE(r, a.lb + 3, a.ub + 3);
}

// r = a + p
B381_HD B381_INL void fp_add_p(Fp& r, const Fp& a) {
  B381_CC_DECL;
  ADD_CC(r.l[0], a.l[0], pword(0));
#pragma unroll
  for (int i = 1; i < NL - 1; i++) ADDC_CC(r.l[i], a.l[i], pword(i));
  ADDC(r.l[NL - 1], a.l[NL - 1], 0u);
  B381_SETRANGE(r, a.lb + 1, a.ub + 1);
}

B381_HD B381_INL void fp_add(Fp& r, const Fp& a, const Fp& b) {
  B381_CC_DECL;
  ADD_CC(r.l[0], a.l[0], b.l[0]);
#pragma unroll
  for (int k = 1; k < NL - 1; k++) ADDC_CC(r.l[k], a.l[k], b.l[k]);
  ADDC(r.l[NL - 1], a.l[NL - 1], b.l[NL - 1]);
  B381_SETRANGE(r, a.lb + b.lb, a.ub + b.ub);
  B381_CHECK(r.mag < FP_MAG_MAX, "fp_add: magnitude");
}

B381_HD B381_INL void fp_sub(Fp& r, const Fp& a, const Fp& b) {
  B381_CC_DECL;
  SUB_CC(r.l[0], a.l[0], b.l[0]);
#pragma unroll
  for (int k = 1; k < NL - 1; k++) SUBC_CC(r.l[k], a.l[k], b.l[k]);
  SUBC(r.l[NL - 1], a.l[NL - 1], b.l[NL - 1]);
  B381_SETRANGE(r, a.lb - b.ub, a.ub - b.lb);
  B381_CHECK(r.mag < FP_MAG_MAX, "fp_sub: magnitude");
}

B381_HD B381_INL void fp_neg(Fp& r, const Fp& a) {
  B381_CC_DECL;
  SUB_CC(r.l[0], 0u, a.l[0]);
#pragma unroll
  for (int k = 1; k < NL - 1; k++) SUBC_CC(r.l[k], 0u, a.l[k]);
  SUBC(r.l[NL - 1], 0u, a.l[NL - 1]);
  B381_SETRANGE(r, -a.ub, -a.lb);
}

B381_HD B381_INL void fp_dbl(Fp& r, const Fp& a) {
  // shift left by one (funnel shifts; r may alias a: go from the top down)
#pragma unroll
  for (int k = NL - 1; k > 0; k--) r.l[k] = (a.l[k] << 1) | (a.l[k - 1] >> 31);
  r.l[0] = a.l[0] << 1;
  B381_SETRANGE(r, 2 * a.lb, 2 * a.ub);
  B381_CHECK(r.mag < FP_MAG_MAX, "fp_dbl: magnitude");
}

// words are always exact in this format
B381_HD B381_INL void fp_norm(Fp&) {}
B381_HD B381_INL void fp_carry_exact(Fp&) {}

B381_HD B381_INL void fp_set(Fp& r, const uint32_t (&v)[NL]) {
#pragma unroll
  for (int k = 0; k < NL; k++) r.l[k] = v[k];
  B381_SETRANGE(r, 0.0, 1.0);
}

// exact halving mod p: (v + (v odd ? p : 0)) >> 1 (arithmetic).  Equals multiplication by 2^-1
// (ark-ec g2.rs double_in_place's mul_assign_by_fp(two_inv)).
B381_HD B381_INL void fp_half(Fp& r, const Fp& a) {
  const uint32_t odd = 0u - (a.l[0] & 1u);
  Fp t;
  B381_CC_DECL;
  ADD_CC(t.l[0], a.l[0], pword(0) & odd);
#pragma unroll
  for (int k = 1; k < NL - 1; k++) ADDC_CC(t.l[k], a.l[k], pword(k) & odd);
  ADDC(t.l[NL - 1], a.l[NL - 1], 0u);
#pragma unroll
  for (int k = 0; k < NL - 1; k++) r.l[k] = (t.l[k] >> 1) | (t.l[k + 1] << 31);
  r.l[NL - 1] = (uint32_t)((int32_t)t.l[NL - 1] >> 1);
  B381_SETRANGE(r, a.lb >= 0 ? 0 : a.lb, (a.ub + 1) / 2 > 0 ? (a.ub + 1) / 2 : 0);
}

// ---------------------------------------------------------------------------------------------
// double-width accumulators
// ---------------------------------------------------------------------------------------------
B381_HD B381_INL void acc_zero(Acc& t) {
#pragma unroll
  for (int k = 0; k < NW; k++) t.c[k] = 0;
  B381_TB(t.cb = 0; t.mag = 0;)
}

B381_HD B381_INL void acc_add(Acc& r, const Acc& a, const Acc& b) {
  B381_CC_DECL;
  ADD_CC(r.c[0], a.c[0], b.c[0]);
#pragma unroll
  for (int k = 1; k < NW - 1; k++) ADDC_CC(r.c[k], a.c[k], b.c[k]);
  ADDC(r.c[NW - 1], a.c[NW - 1], b.c[NW - 1]);
  B381_TB(r.cb = a.cb + b.cb; r.mag = a.mag + b.mag;)
}

B381_HD B381_INL void acc_sub(Acc& r, const Acc& a, const Acc& b) {
  B381_CC_DECL;
  SUB_CC(r.c[0], a.c[0], b.c[0]);
#pragma unroll
  for (int k = 1; k < NW - 1; k++) SUBC_CC(r.c[k], a.c[k], b.c[k]);
  SUBC(r.c[NW - 1], a.c[NW - 1], b.c[NW - 1]);
  B381_TB(r.cb = a.cb - b.mag; r.mag = a.mag + b.mag;)
}

// t = 2 t (two's complement: the bit pattern shifts the same way for either sign)
B381_HD B381_INL void acc_dbl(Acc& t) {
#pragma unroll
  for (int k = NW - 1; k > 0; k--) t.c[k] = (t.c[k] << 1) | (t.c[k - 1] >> 31);
  t.c[0] <<= 1;
  B381_TB(t.cb *= 2; t.mag *= 2;)
}

// t = a * b as a 26-word two's-complement integer: 169 IMAD.WIDE.  Row i multiplies a by word i of
// b; the even words of a form one carry chain of 64-bit products, the odd words a second one, one
// word higher.  IMAD.WIDE reads and writes ALIGNED register pairs, so the chains that start on an
// even word offset accumulate into E (E[k] = word k) and those on an odd offset into O (O[k] = word
// k + 1): every product then lands on an aligned pair of its array (a single array costs two MOVs
// per misaligned product).  The partial sums after row i fit words 0..i+13, so each chain ends in
// the two words above its last product and never ripples further.  E and O are merged by one
// 26-word addition; the signed interpretation (v = u - 2^416 s) costs two masked 13-word
// subtractions from the high half.
#define B381_MUL_CHAIN(X, s, a, first, cnt, bi, linked, tail)                                          \
  {                                                                                                    \
    if (linked) { MADC_LO_CC(X[(s)], a.l[(first)], bi); } else { MAD_LO_CC(X[(s)], a.l[(first)], bi); }  \
    MADC_HI_CC(X[(s) + 1], a.l[(first)], bi);                                                          \
    _Pragma("unroll") for (int k_ = 1; k_ < (cnt); k_++) {                                              \
      MADC_LO_CC(X[(s) + 2 * k_], a.l[(first) + 2 * k_], bi); MADC_HI_CC(X[(s) + 2 * k_ + 1], a.l[(first) + 2 * k_], bi); \
    }                                                                                                  \
    if (tail) { ADDC_CC(X[(s) + 2 * (cnt)], X[(s) + 2 * (cnt)], 0u); }                                  \
  }

// N = 13: any operands;  N = 12: both operands below 2^384 (top word zero, asserted by the tracker):
// 144 multiplies instead of 169.
template <bool SIGNED, int N>
B381_HD B381_INL void acc_mul_t(Acc& t, const Fp& a, const Fp& b) {
  B381_CC_DECL;
  constexpr int NE = (N + 1) / 2, NO = N / 2;       // even / odd words of a
  uint32_t E[30], O[30];
#pragma unroll
  for (int k = 0; k < 30; k++) { E[k] = 0; O[k] = 0; }
  // Chain ends.  A chain whose last word nothing has written yet cannot carry out; otherwise it
  // pushes its carry into the next word, which is fresh (hi = highest word written so far; the
  // loops are fully unrolled, so all of this folds at compile time).  All chains of one array are
  // LINKED through the carry flag (the flag handed on is always zero): the link is semantically void
  // but makes each array one serial dependency chain, so ptxas interleaves exactly two chains (E
  // and O) instead of a wavefront of thirteen, whose live carries exceed the seven predicate
  // registers and get spilled through LOP3 bit twiddling.
  int hi = -1;
#pragma unroll
  for (int i = 0; i < N; i++) {
    const uint32_t bi = b.l[i];
    if ((i & 1) == 0) {                              // even words of a at words (i + 2k, i + 2k + 1)
      const int last = i + 2 * NE - 1; const bool tail = last <= hi;
      B381_MUL_CHAIN(E, i, a, 0, NE, bi, i != 0, tail);
      hi = tail ? last + 1 : last;
    } else {                                         // odd words of a at words (i + 2k + 1, i + 2k + 2)
      const int last = i + 1 + 2 * NO - 1; const bool tail = last <= hi;
      B381_MUL_CHAIN(E, i + 1, a, 1, NO, bi, 1, tail);
      hi = tail ? last + 1 : (last > hi ? last : hi);
    }
  }
  hi = -1;
#pragma unroll
  for (int i = 0; i < N; i++) {
    const uint32_t bi = b.l[i];
    if ((i & 1) == 0) {
      const int last = i + 2 * NO - 1; const bool tail = last <= hi;
      B381_MUL_CHAIN(O, i, a, 1, NO, bi, i != 0, tail);
      hi = tail ? last + 1 : (last > hi ? last : hi);
    } else {
      const int last = i - 1 + 2 * NE - 1; const bool tail = last <= hi;
      B381_MUL_CHAIN(O, i - 1, a, 0, NE, bi, 1, tail);
      hi = tail ? last + 1 : (last > hi ? last : hi);
    }
  }
  B381_TB(if (N != NL && !(a.ub < 9.8 && b.ub < 9.8 && a.lb >= 0 && b.lb >= 0)) fprintf(stderr, "acc_mul<12>: a in [%g, %g] b in [%g, %g]\n", a.lb, a.ub, b.lb, b.ub);)
  B381_CHECK(N == NL || (a.ub < 9.8 && b.ub < 9.8 && a.lb >= 0 && b.lb >= 0), "acc_mul<12>: operand does not fit 12 words");
  uint32_t (&T)[NW] = t.c;
  T[0] = E[0];
  ADD_CC(T[1], E[1], O[0]);
#pragma unroll
  for (int k = 2; k < NW - 1; k++) ADDC_CC(T[k], E[k], O[k - 1]);
  ADDC(T[NW - 1], E[NW - 1], O[NW - 2]);
  if (SIGNED) {
    // sign corrections: (ua - 2^416 sa)(ub - 2^416 sb) = ua ub - 2^416 (sa ub + sb ua)   (mod 2^832)
    const uint32_t ma = (uint32_t)((int32_t)a.l[NL - 1] >> 31), mb = (uint32_t)((int32_t)b.l[NL - 1] >> 31);
    SUB_CC(T[NL], T[NL], b.l[0] & ma);
#pragma unroll
    for (int k = 1; k < NL - 1; k++) SUBC_CC(T[NL + k], T[NL + k], b.l[k] & ma);
    SUBC(T[NW - 1], T[NW - 1], b.l[NL - 1] & ma);
    SUB_CC(T[NL], T[NL], a.l[0] & mb);
#pragma unroll
    for (int k = 1; k < NL - 1; k++) SUBC_CC(T[NL + k], T[NL + k], a.l[k] & mb);
    SUBC(T[NW - 1], T[NW - 1], a.l[NL - 1] & mb);
  } else {
    B381_CHECK(a.lb >= 0 && b.lb >= 0, "acc_mul: operands of the unsigned multiplication must be non-negative");
  }
  B381_TB(t.cb = SIGNED ? -(a.mag * b.mag) : 0; t.mag = a.mag * b.mag;)
  B381_CHECK(a.mag < MUL_MAG_MAX && b.mag < MUL_MAG_MAX, "acc_mul: operand magnitude");
}

// Stored values are kept NON-NEGATIVE (products are made so by adding p where the double-width value
// can be negative, differences by the weak reduction or a 128 p offset), so the hot path multiplies
// unsigned words and needs no sign correction; acc_mul_signed is for the rare canonicalisation paths.
B381_HD B381_INL void acc_mul(Acc& t, const Fp& a, const Fp& b) { acc_mul_t<false, NL>(t, a, b); }
B381_HD B381_INL void acc_mul12(Acc& t, const Fp& a, const Fp& b) { acc_mul_t<false, 12>(t, a, b); }
B381_HD B381_INL void acc_mul_signed(Acc& t, const Fp& a, const Fp& b) { acc_mul_t<true, NL>(t, a, b); }

// t = a0 b0 + a1 b1 + a2 b2 (non-negative operands) accumulated in ONE pair of even / odd arrays:
// all three products of a row are chained before the next row starts, so the words above the row
// only ever hold small carry counts and each chain still ends with a single carry add.  Compared
// with three separate products this saves two 26-word merges and two 26-word additions.
template <int N>
B381_HD B381_INL void acc_mul3_t(Acc& t, const Fp& a0, const Fp& b0, const Fp& a1, const Fp& b1, const Fp& a2, const Fp& b2) {
  B381_CC_DECL;
  constexpr int NE = (N + 1) / 2, NO = N / 2;
  uint32_t E[30], O[30];
#pragma unroll
  for (int k = 0; k < 30; k++) { E[k] = 0; O[k] = 0; }
#pragma unroll
  for (int i = 0; i < N; i++) {
    const uint32_t x0 = b0.l[i], x1 = b1.l[i], x2 = b2.l[i];
    if ((i & 1) == 0) {
      B381_MUL_CHAIN(E, i, a0, 0, NE, x0, i != 0, 1); B381_MUL_CHAIN(E, i, a1, 0, NE, x1, 1, 1); B381_MUL_CHAIN(E, i, a2, 0, NE, x2, 1, 1);
    } else {
      B381_MUL_CHAIN(E, i + 1, a0, 1, NO, x0, 1, 1); B381_MUL_CHAIN(E, i + 1, a1, 1, NO, x1, 1, 1); B381_MUL_CHAIN(E, i + 1, a2, 1, NO, x2, 1, 1);
    }
  }
#pragma unroll
  for (int i = 0; i < N; i++) {
    const uint32_t x0 = b0.l[i], x1 = b1.l[i], x2 = b2.l[i];
    if ((i & 1) == 0) {
      B381_MUL_CHAIN(O, i, a0, 1, NO, x0, i != 0, 1); B381_MUL_CHAIN(O, i, a1, 1, NO, x1, 1, 1); B381_MUL_CHAIN(O, i, a2, 1, NO, x2, 1, 1);
    } else {
      B381_MUL_CHAIN(O, i - 1, a0, 0, NE, x0, 1, 1); B381_MUL_CHAIN(O, i - 1, a1, 0, NE, x1, 1, 1); B381_MUL_CHAIN(O, i - 1, a2, 0, NE, x2, 1, 1);
    }
  }
  B381_CHECK(N == NL || (a0.ub < 9.8 && b0.ub < 9.8 && a1.ub < 9.8 && b1.ub < 9.8 && a2.ub < 9.8 && b2.ub < 9.8), "acc_mul3<12>: operand does not fit 12 words");
  uint32_t (&T)[NW] = t.c;
  T[0] = E[0];
  ADD_CC(T[1], E[1], O[0]);
#pragma unroll
  for (int k = 2; k < NW - 1; k++) ADDC_CC(T[k], E[k], O[k - 1]);
  ADDC(T[NW - 1], E[NW - 1], O[NW - 2]);
  B381_CHECK(a0.lb >= 0 && b0.lb >= 0 && a1.lb >= 0 && b1.lb >= 0 && a2.lb >= 0 && b2.lb >= 0, "acc_mul3: operands must be non-negative");
  B381_CHECK(a0.mag < MUL_MAG_MAX && b0.mag < MUL_MAG_MAX && a1.mag < MUL_MAG_MAX && b1.mag < MUL_MAG_MAX && a2.mag < MUL_MAG_MAX && b2.mag < MUL_MAG_MAX, "acc_mul3: operand magnitude");
  B381_TB(t.cb = 0; t.mag = a0.mag * b0.mag + a1.mag * b1.mag + a2.mag * b2.mag;)
}
B381_HD B381_INL void acc_mul3(Acc& t, const Fp& a0, const Fp& b0, const Fp& a1, const Fp& b1, const Fp& a2, const Fp& b2) { acc_mul3_t<NL>(t, a0, b0, a1, b1, a2, b2); }
B381_HD B381_INL void acc_mul3_12(Acc& t, const Fp& a0, const Fp& b0, const Fp& a1, const Fp& b1, const Fp& a2, const Fp& b2) { acc_mul3_t<12>(t, a0, b0, a1, b1, a2, b2); }

// t += a * b
B381_HD B381_INL void acc_mac(Acc& t, const Fp& a, const Fp& b) {
  Acc u;
  B381_TB(u.mag = 0; u.cb = 0;)
  acc_mul(u, a, b);
  acc_add(t, t, u);
}
B381_HD B381_INL void acc_mac_cross(Acc& t, const Fp& a, const Fp& b) { acc_mac(t, a, b); }
B381_HD B381_INL void acc_mac12(Acc& t, const Fp& a, const Fp& b) {
  Acc u;
  B381_TB(u.mag = 0; u.cb = 0;)
  acc_mul12(u, a, b);
  acc_add(t, t, u);
}

B381_HD B381_INL void acc_neg(Acc& r, const Acc& a) {
  B381_CC_DECL;
  SUB_CC(r.c[0], 0u, a.c[0]);
#pragma unroll
  for (int k = 1; k < NW - 1; k++) SUBC_CC(r.c[k], 0u, a.c[k]);
  SUBC(r.c[NW - 1], 0u, a.c[NW - 1]);
  B381_TB(r.cb = -a.mag; r.mag = a.mag;)
}

// Montgomery reduction with a sliding two-array window (the even / odd scheme of the public sppark
// field code, rearranged for a separate reduction of a double-width value).  At row i, A[k] is word
// i + k and B[k] word i + k + 1 of U = t_low + M p; A[0] is the exact low word.  m = A[0] n0';
// the odd words of p accumulate into B, the even words into A (both on aligned pairs); then A[0] = 0,
// A[1] is folded into B[0] (its carry enters the next row's first chain), B becomes the new A and
// A shifted by one PAIR the new B.  ROWS * 12 IMAD.WIDE + ROWS IMAD; nothing ever ripples.
#define B381_REDC_ROW(A, B, cin)                                                                       \
  {                                                                                                    \
    const uint32_t m_ = A[0] * (uint32_t)B381_N0Q;                                                     \
    if (cin) { MADC_LO_CC(B[0], m_, (uint32_t)B381_Q1); } else { MAD_LO_CC(B[0], m_, (uint32_t)B381_Q1); }  \
    MADC_HI_CC(B[1], m_, (uint32_t)B381_Q1);                                                           \
    MADC_LO_CC(B[2], m_, (uint32_t)B381_Q3);  MADC_HI_CC(B[3], m_, (uint32_t)B381_Q3);                  \
    MADC_LO_CC(B[4], m_, (uint32_t)B381_Q5);  MADC_HI_CC(B[5], m_, (uint32_t)B381_Q5);                  \
    MADC_LO_CC(B[6], m_, (uint32_t)B381_Q7);  MADC_HI_CC(B[7], m_, (uint32_t)B381_Q7);                  \
    MADC_LO_CC(B[8], m_, (uint32_t)B381_Q9);  MADC_HI_CC(B[9], m_, (uint32_t)B381_Q9);                  \
    MADC_LO_CC(B[10], m_, (uint32_t)B381_Q11); MADC_HI_CC(B[11], m_, (uint32_t)B381_Q11);               \
    ADDC_CC(B[12], B[12], 0u);           /* B[12] was zero: no carry out, the flag links the chains */  \
    MADC_LO_CC(A[0], m_, (uint32_t)B381_Q0);  MADC_HI_CC(A[1], m_, (uint32_t)B381_Q0);                  \
    MADC_LO_CC(A[2], m_, (uint32_t)B381_Q2);  MADC_HI_CC(A[3], m_, (uint32_t)B381_Q2);                  \
    MADC_LO_CC(A[4], m_, (uint32_t)B381_Q4);  MADC_HI_CC(A[5], m_, (uint32_t)B381_Q4);                  \
    MADC_LO_CC(A[6], m_, (uint32_t)B381_Q6);  MADC_HI_CC(A[7], m_, (uint32_t)B381_Q6);                  \
    MADC_LO_CC(A[8], m_, (uint32_t)B381_Q8);  MADC_HI_CC(A[9], m_, (uint32_t)B381_Q8);                  \
    MADC_LO_CC(A[10], m_, (uint32_t)B381_Q10); MADC_HI_CC(A[11], m_, (uint32_t)B381_Q10);               \
    ADDC_CC(A[12], A[12], 0u);           /* later rows: A[12] held at most one earlier carry, no carry out */ \
    if (!(cin)) { ADDC_CC(A[13], A[13], 0u); }  /* row 0: A[12] is word 12 of the INPUT, any value -- 0xffffffff + carry ripples into word 13 */ \
  }

// r = t / 2^(32 ROWS) mod p, result in (t / 2^(32 ROWS), t / 2^(32 ROWS) + p].  Only the low ROWS
// words of t enter the rows; the (signed) high half of t is added at the end.
template <int ROWS>
B381_HD B381_INL void acc_redc_rows(Fp& r, Acc& t) {
  B381_CC_DECL;
  uint32_t A[14], B[14];
#pragma unroll
  for (int k = 0; k < 14; k++) { A[k] = k < ROWS ? t.c[k] : 0u; B[k] = 0; }
#pragma unroll
  for (int i = 0; i < ROWS; i++) {
    B381_REDC_ROW(A, B, i != 0);
    if (i + 1 < ROWS) {
      ADDC_CC(B[0], B[0], A[1]);                      // fold (flag in is zero); its carry enters the next row's first chain
      uint32_t N[14];
#pragma unroll
      for (int k = 0; k < 12; k++) N[k] = A[k + 2];
      N[12] = 0; N[13] = 0;
#pragma unroll
      for (int k = 0; k < 14; k++) { A[k] = B[k]; B[k] = N[k]; }
    }
  }
  // words ROWS + k of U are B[k] + A[k + 1]
  uint32_t h[NL];
  ADD_CC(h[0], B[0], A[1]);
#pragma unroll
  for (int k = 1; k < NL - 1; k++) ADDC_CC(h[k], B[k], A[k + 1]);
  ADDC(h[NL - 1], B[NL - 1], A[NL]);
  ADD_CC(r.l[0], h[0], t.c[ROWS]);
#pragma unroll
  for (int k = 1; k < NL - 1; k++) ADDC_CC(r.l[k], h[k], t.c[ROWS + k]);
  ADDC(r.l[NL - 1], h[NL - 1], t.c[ROWS + NL - 1]);
}

B381_HD B381_INL void acc_redc(Fp& r, Acc& t) {
  B381_CHECK(t.mag < ACC_MAG_MAX, "acc_redc: input too large");
  acc_redc_rows<NL>(r, t);
  B381_SETRANGE(r, t.cb < 0 ? t.cb / 4.2e10 : 0.0, 1.0 + t.mag / 4.2e10 + 1e-9);
}

// two independent reductions with their rows interleaved in source order (the latency of one row's
// m = A[0] n0' chain hides behind the other reduction's twelve multiplies).  The fold carry of each
// reduction is parked in a register across the other reduction's row and put back into the carry
// flag (c + 0xffffffff carries iff c = 1) in front of its next row.
#define B381_REDC_STEP(A, B, c, i)                                                                     \
  {                                                                                                    \
    if ((i) != 0) { uint32_t d_; ADD_CC(d_, c, 0xffffffffu); (void)d_; }                                \
    B381_REDC_ROW(A, B, (i) != 0);                                                                     \
    if ((i) + 1 < ROWS) {                                                                              \
      ADDC_CC(B[0], B[0], A[1]);                                                                       \
      ADDC(c, 0u, 0u);                                                                                 \
      uint32_t N_[14];                                                                                 \
      _Pragma("unroll") for (int k_ = 0; k_ < 12; k_++) N_[k_] = A[k_ + 2];                             \
      N_[12] = 0; N_[13] = 0;                                                                          \
      _Pragma("unroll") for (int k_ = 0; k_ < 14; k_++) { A[k_] = B[k_]; B[k_] = N_[k_]; }              \
    }                                                                                                  \
  }

B381_HD B381_INL void acc_redc2(Fp& r0, Acc& t0, Fp& r1, Acc& t1) {
  B381_CHECK(t0.mag < ACC_MAG_MAX && t1.mag < ACC_MAG_MAX, "acc_redc2: input too large");
  B381_CC_DECL;
  constexpr int ROWS = NL;
  uint32_t A0[14], B0[14], A1[14], B1[14], c0 = 0, c1 = 0;
#pragma unroll
  for (int k = 0; k < 14; k++) { A0[k] = k < ROWS ? t0.c[k] : 0u; B0[k] = 0; A1[k] = k < ROWS ? t1.c[k] : 0u; B1[k] = 0; }
#pragma unroll
  for (int i = 0; i < ROWS; i++) {
    B381_REDC_STEP(A0, B0, c0, i);
    B381_REDC_STEP(A1, B1, c1, i);
  }
  uint32_t h[NL];
  ADD_CC(h[0], B0[0], A0[1]);
#pragma unroll
  for (int k = 1; k < NL - 1; k++) ADDC_CC(h[k], B0[k], A0[k + 1]);
  ADDC(h[NL - 1], B0[NL - 1], A0[NL]);
  ADD_CC(r0.l[0], h[0], t0.c[ROWS]);
#pragma unroll
  for (int k = 1; k < NL - 1; k++) ADDC_CC(r0.l[k], h[k], t0.c[ROWS + k]);
  ADDC(r0.l[NL - 1], h[NL - 1], t0.c[NW - 1]);
  ADD_CC(h[0], B1[0], A1[1]);
#pragma unroll
  for (int k = 1; k < NL - 1; k++) ADDC_CC(h[k], B1[k], A1[k + 1]);
  ADDC(h[NL - 1], B1[NL - 1], A1[NL]);
  ADD_CC(r1.l[0], h[0], t1.c[ROWS]);
#pragma unroll
  for (int k = 1; k < NL - 1; k++) ADDC_CC(r1.l[k], h[k], t1.c[ROWS + k]);
  ADDC(r1.l[NL - 1], h[NL - 1], t1.c[NW - 1]);
  B381_SETRANGE(r0, t0.cb < 0 ? t0.cb / 4.2e10 : 0.0, 1.0 + t0.mag / 4.2e10 + 1e-9);
  B381_SETRANGE(r1, t1.cb < 0 ? t1.cb / 4.2e10 : 0.0, 1.0 + t1.mag / 4.2e10 + 1e-9);
}

// r = t / 2^384 mod p: reduction in the EXTERNAL domain (12 rows), for the element-wise Fp / Fp2
// multiply entry points, which then need no domain conversion.  Canonical operands give (−p, 2p).
B381_HD B381_INL void acc_redc384(Fp& r, Acc& t) {
  B381_CHECK(t.mag < 4.0, "acc_redc384: operands must be canonical");
  acc_redc_rows<12>(r, t);
  B381_SETRANGE(r, t.cb < 0 ? t.cb / 9.8 : 0.0, 1.0 + t.mag / 9.8);
}

B381_HD B381_INL void fp_mul(Fp& r, const Fp& a, const Fp& b) {
  Acc t;
  B381_TB(t.mag = 0; t.cb = 0;)
  acc_mul(t, a, b);
  acc_redc(r, t);
}
// both operands below 2^384 (tracker-asserted): 144 + 156 IMAD.WIDE
B381_HD B381_INL void fp_mul12(Fp& r, const Fp& a, const Fp& b) {
  Acc t;
  B381_TB(t.mag = 0; t.cb = 0;)
  acc_mul12(t, a, b);
  acc_redc(r, t);
}

// weak reduction: any |v| < 2^20 p comes out in [0, 1.02 p).  v' = v + 2^21 p >= 0; q = floor(h / D) with
// h = v' >> 352 (top two words, below 2^51) and D = (p >> 352) + 1, by a 64-bit reciprocal and one
// correction step; q p <= v', and v' - q p < 2^352 (1 + h / D) + p < 1.012 p.  12 IMAD.WIDE + a handful
// of 64-bit operations.  Needed wherever a value feeds back LINEARLY into itself (the -2z term of the
// cyclotomic squaring) and to bring differences back to small non-negative values.
B381_HD B381_INL uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

B381_HD B381_INL void fp_wreduce(Fp& a) {
  B381_CHECK(a.mag < FP_MAG_MAX, "fp_wreduce: value out of range");
  const uint32_t off[NL] = B381_WRED_OFF;       // 2^21 p
  uint32_t v[NL];
  B381_CC_DECL;
  ADD_CC(v[0], a.l[0], off[0]);
#pragma unroll
  for (int k = 1; k < NL - 1; k++) ADDC_CC(v[k], a.l[k], off[k]);
  ADDC(v[NL - 1], a.l[NL - 1], off[NL - 1]);
  const uint64_t h = ((uint64_t)v[NL - 1] << 32) | v[NL - 2];
  uint64_t q64 = mulhi64(h, (uint64_t)B381_WRED_M) >> 28;
  if (h - q64 * (uint64_t)B381_WRED_D >= (uint64_t)B381_WRED_D) q64 += 1;
  const uint32_t q = (uint32_t)q64;
  uint32_t Q[NL];
#pragma unroll
  for (int k = 0; k < 6; k++) MUL_WIDE(Q[2 * k], Q[2 * k + 1], q, pword(2 * k));
  Q[12] = 0;
  MAD_LO_CC(Q[1], q, pword(1)); MADC_HI_CC(Q[2], q, pword(1));
#pragma unroll
  for (int k = 1; k < 6; k++) { MADC_LO_CC(Q[2 * k + 1], q, pword(2 * k + 1)); MADC_HI_CC(Q[2 * k + 2], q, pword(2 * k + 1)); }
  SUB_CC(a.l[0], v[0], Q[0]);
#pragma unroll
  for (int k = 1; k < NL - 1; k++) SUBC_CC(a.l[k], v[k], Q[k]);
  SUBC(a.l[NL - 1], v[NL - 1], Q[NL - 1]);
  B381_SETRANGE(a, 0.0, 1.02);
}

// ---------------------------------------------------------------------------------------------
// canonical form, comparisons, external format (12 x u32, Montgomery R = 2^384)
// ---------------------------------------------------------------------------------------------
// bring a value in (-p, 2p) to [0, p)
B381_HD B381_INL void fp_canon_small(Fp& a) {
  B381_CHECK(a.mag < 2.0, "fp_canon_small: input range");
  B381_CC_DECL;
  const uint32_t neg = (uint32_t)((int32_t)a.l[NL - 1] >> 31);
  ADD_CC(a.l[0], a.l[0], pword(0) & neg);
#pragma unroll
  for (int k = 1; k < NL - 1; k++) ADDC_CC(a.l[k], a.l[k], pword(k) & neg);
  ADDC(a.l[NL - 1], a.l[NL - 1], 0u);
  Fp t;
  SUB_CC(t.l[0], a.l[0], pword(0));
#pragma unroll
  for (int k = 1; k < NL - 1; k++) SUBC_CC(t.l[k], a.l[k], pword(k));
  SUBC(t.l[NL - 1], a.l[NL - 1], 0u);
  const uint32_t ge = ~(uint32_t)((int32_t)t.l[NL - 1] >> 31);       // all ones if a >= p
#pragma unroll
  for (int k = 0; k < NL; k++) a.l[k] = (t.l[k] & ge) | (a.l[k] & ~ge);
  B381_SETRANGE(a, 0.0, 1.0);
}

// full reduction of any stored value to canonical [0,p): one Montgomery multiplication by R' mod p
B381_HD B381_INL void fp_canon(Fp& a) {
  const uint32_t one[NL] = B381_ONE;
  Fp o, n = a;
  fp_set(o, one);
  Acc t;
  B381_TB(t.mag = 0; t.cb = 0;)
  acc_mul_signed(t, n, o);
  acc_redc(a, t);
  fp_canon_small(a);
}

B381_HD B381_INL bool fp_is_zero_canon(const Fp& a) {
  uint32_t o = 0;
#pragma unroll
  for (int k = 0; k < NL; k++) o |= a.l[k];
  return o == 0;
}

B381_HD B381_INL bool fp_eq_canon(const Fp& a, const Fp& b) {
  uint32_t o = 0;
#pragma unroll
  for (int k = 0; k < NL; k++) o |= a.l[k] ^ b.l[k];
  return o == 0;
}

// 12 x u32 (plain integer X < 2^384) -> words (top word zero)
B381_HD B381_INL void fp_unpack32(Fp& r, const uint32_t (&w)[12]) {
#pragma unroll
  for (int k = 0; k < 12; k++) r.l[k] = w[k];
  r.l[12] = 0;
  B381_SETRANGE(r, 0.0, 9.9);     // any 384-bit integer
}

// canonical words -> 12 x u32
B381_HD B381_INL void fp_pack32(uint32_t (&w)[12], const Fp& a) {
#pragma unroll
  for (int j = 0; j < 12; j++) w[j] = a.l[j];
}

// range check of an unpacked plain integer (top word zero): X < p
B381_HD B381_INL bool fp_below_p(const Fp& x) {
  uint32_t t;
  B381_CC_DECL;
  SUB_CC(t, x.l[0], pword(0));
#pragma unroll
  for (int k = 1; k < NL - 1; k++) SUBC_CC(t, x.l[k], pword(k));
  SUBC(t, 0u, 0u);                                   // 0 - borrow: all ones if X < p
  return t != 0;
}

// X (12 x u32, Montgomery R = 2^384, canonical) -> internal.  Returns false if X >= p.
B381_HD B381_INL bool fp_from_ext(Fp& r, const uint32_t (&w)[12]) {
  Fp x;
  fp_unpack32(x, w);
  const bool ok = fp_below_p(x);
  const uint32_t cin[NL] = B381_CIN;
  Fp c;
  fp_set(c, cin);
  fp_mul(r, x, c);
  return ok;
}

// internal -> 12 x u32, Montgomery R = 2^384, canonical
B381_HD B381_INL void fp_to_ext(uint32_t (&w)[12], const Fp& a) {
  const uint32_t cout[NL] = B381_COUT;
  Fp c, o;
  fp_set(c, cout);
  B381_CHECK(a.mag < MUL_MAG_MAX, "fp_to_ext: magnitude");
  fp_mul(o, a, c);
  fp_canon_small(o);
  fp_pack32(w, o);
}

// The external words taken AS the stored value (no multiplication): the element is then off by the fixed factor
// 2^-32 (R = 2^384 against R' = 2^416).  For a bilinear map of two such operands the result is off by 2^-64 and
// fp_to_ext_unscale puts it right with ONE multiplication: stored x~ -> x~ * 2^448 / R' = x~ * 2^32 ... applied to the
// stored product (a 2^-32)(b 2^-32) R' = a b 2^352 it gives a b 2^384 = the external form of a b.
B381_HD B381_INL bool fp_from_ext_asis(Fp& r, const uint32_t (&w)[12]) {
  fp_unpack32(r, w);
  const bool ok = fp_below_p(r);
  B381_SETRANGE(r, 0.0, 1.0);
  return ok;
}
B381_HD B381_INL void fp_to_ext_unscale(uint32_t (&w)[12], const Fp& a) {
  const uint32_t cin[NL] = B381_CIN;              // 2^448 mod p
  Fp c, o;
  fp_set(c, cin);
  B381_CHECK(a.mag < MUL_MAG_MAX, "fp_to_ext_unscale: magnitude");
  fp_mul(o, a, c);
  fp_canon_small(o);
  fp_pack32(w, o);
}

// ---------------------------------------------------------------------------------------------
// modular inversion: Bernstein-Yang "safegcd" division steps (constant control flow, no multiplications
// in the inner loop).  The stored value v = A R' is made canonical, then (f, g) = (p, v) runs through
// batches of 30 division steps on the low words; each batch's 2 x 2 transition matrix is applied to the
// full-length (f, g) and, modulo p, to (d, e) = (0, R'^2 mod p), so that d ends at +-v^-1 R'^2 = +-A^-1 R'.
// Values are 13 signed limbs of 30 bits with 64-bit signed accumulators (the layout of the public
// libsecp256k1 modinv32 code, written out for a 381-bit modulus).  Theorem 11.2 of Bernstein-Yang 2019
// (delta starts at 1; f^2 + 4 g^2 <= 5 * 2^(2 * 381)) bounds the number of division steps by
// floor((49 * 381 + 57) / 17) = 1101 <= 37 * 30.  On the device a warp leaves the loop as soon as every
// active lane has g = 0 (random inputs: 26 - 28 batches), which is the fixed point of the iteration.
// About 20 k ALU instructions + 4 k IMAD.WIDE against 183 k IMAD.WIDE for the Fermat power a^(p-2):
// the same residue, a tenth of the time.  inverse(0) = 0 (as a^(p-2)).
// ---------------------------------------------------------------------------------------------
constexpr int SG_N = 13;
constexpr int SG_BATCHES = 37;
constexpr uint32_t SG_M30 = 0x3fffffffu;

B381_HD B381_INL void sg_divsteps30(uint32_t& eta, uint32_t f, uint32_t g, int32_t& tu, int32_t& tv, int32_t& tq, int32_t& tr) {
  uint32_t u = 1, v = 0, q = 0, r = 1;
#pragma unroll 6
  for (int i = 0; i < 30; i++) {
    uint32_t c1 = (uint32_t)((int32_t)eta >> 31);       // delta > 0
    const uint32_t c2 = 0u - (g & 1u);                  // g odd
    const uint32_t x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;
    g += x & c2; q += y & c2; r += z & c2;
    c1 &= c2;                                           // swap
    eta = (eta ^ c1) - (c1 + 1u);
    f += g & c1; u += q & c1; v += r & c1;
    g >>= 1; u <<= 1; v <<= 1;
  }
  tu = (int32_t)u; tv = (int32_t)v; tq = (int32_t)q; tr = (int32_t)r;
}

// (f, g) <- t (f, g) / 2^30 (exact)
B381_HD B381_INL void sg_update_fg(int32_t (&f)[SG_N], int32_t (&g)[SG_N], int32_t u, int32_t v, int32_t q, int32_t r) {
  int64_t cf = (int64_t)u * f[0] + (int64_t)v * g[0];
  int64_t cg = (int64_t)q * f[0] + (int64_t)r * g[0];
  cf >>= 30; cg >>= 30;
#pragma unroll
  for (int i = 1; i < SG_N; i++) {
    cf += (int64_t)u * f[i] + (int64_t)v * g[i];
    cg += (int64_t)q * f[i] + (int64_t)r * g[i];
    f[i - 1] = (int32_t)((uint32_t)cf & SG_M30); cf >>= 30;
    g[i - 1] = (int32_t)((uint32_t)cg & SG_M30); cg >>= 30;
  }
  f[SG_N - 1] = (int32_t)cf;
  g[SG_N - 1] = (int32_t)cg;
}

// (d, e) <- t (d, e) / 2^30 mod p, both kept in (-2p, p)
B381_HD B381_INL void sg_update_de(int32_t (&d)[SG_N], int32_t (&e)[SG_N], int32_t u, int32_t v, int32_t q, int32_t r) {
  const int32_t pl[SG_N] = B381_SG_P30;
  const int32_t sd = d[SG_N - 1] >> 31, se = e[SG_N - 1] >> 31;
  int32_t md = (u & sd) + (v & se), me = (q & sd) + (r & se);
  int64_t cd = (int64_t)u * d[0] + (int64_t)v * e[0];
  int64_t ce = (int64_t)q * d[0] + (int64_t)r * e[0];
  md -= (int32_t)((B381_SG_PINV30 * (uint32_t)cd + (uint32_t)md) & SG_M30);
  me -= (int32_t)((B381_SG_PINV30 * (uint32_t)ce + (uint32_t)me) & SG_M30);
  cd += (int64_t)pl[0] * md;
  ce += (int64_t)pl[0] * me;
  cd >>= 30; ce >>= 30;
#pragma unroll
  for (int i = 1; i < SG_N; i++) {
    cd += (int64_t)u * d[i] + (int64_t)v * e[i] + (int64_t)pl[i] * md;
    ce += (int64_t)q * d[i] + (int64_t)r * e[i] + (int64_t)pl[i] * me;
    d[i - 1] = (int32_t)((uint32_t)cd & SG_M30); cd >>= 30;
    e[i - 1] = (int32_t)((uint32_t)ce & SG_M30); ce >>= 30;
  }
  d[SG_N - 1] = (int32_t)cd;
  e[SG_N - 1] = (int32_t)ce;
}

// r = a^-1 (Montgomery domain R'), r in [0, p); a any stored value with |a| < 2^17 p
B381_HD B381_INL void fp_inv_safegcd(Fp& r, const Fp& a) {
  Fp c = a;
  fp_canon(c);
  int32_t f[SG_N] = B381_SG_P30, g[SG_N], d[SG_N], e[SG_N] = B381_SG_R2_30;
#pragma unroll
  for (int i = 0; i < SG_N; i++) {                      // limb i = bits 30 i .. 30 i + 29
    const int k = (30 * i) >> 5, s = (30 * i) & 31;
    uint32_t w = c.l[k] >> s;
    if (s > 2 && k + 1 < NL) w |= c.l[k + 1] << (32 - s);
    g[i] = (int32_t)(w & SG_M30);
    d[i] = 0;
  }
  uint32_t eta = 0xffffffffu;                           // -delta, delta = 1
  for (int it = 0; it < SG_BATCHES; it++) {
    int32_t tu, tv, tq, tr;
    sg_divsteps30(eta, (uint32_t)f[0] | ((uint32_t)f[1] << 30), (uint32_t)g[0] | ((uint32_t)g[1] << 30), tu, tv, tq, tr);
    sg_update_de(d, e, tu, tv, tq, tr);
    sg_update_fg(f, g, tu, tv, tq, tr);
#if defined(__CUDA_ARCH__)
    uint32_t nz = 0;
#pragma unroll
    for (int i = 0; i < SG_N; i++) nz |= (uint32_t)g[i];
    if (__all_sync(__activemask(), nz == 0)) break;
#endif
  }
  // f = +-1 (or p when a = 0, with d = 0): d <- sign(f) d, as two's-complement words
  const uint32_t fneg = (uint32_t)(f[SG_N - 1] >> 31);
  Fp t;
  {
    // 13 limbs of 30 bits (top one signed) -> 13 words of 32 bits: word k = bits 32 k .. 32 k + 31
    uint32_t dl[SG_N + 2];
#pragma unroll
    for (int i = 0; i < SG_N; i++) dl[i] = (uint32_t)d[i];
    dl[SG_N] = dl[SG_N + 1] = (uint32_t)(d[SG_N - 1] >> 31);          // sign extension
#pragma unroll
    for (int k = 0; k < NL; k++) {
      const int j = (32 * k) / 30, off = 32 * k - 30 * j;
      // limbs below the top one are non-negative 30-bit values; the top limb (and the extension) is sign-extended
      uint32_t w = dl[j] >> off;
      if (j + 1 < SG_N + 2) w |= dl[j + 1] << (30 - off);
      if (60 - off < 32 && j + 2 < SG_N + 2) w |= dl[j + 2] << (60 - off);
      t.l[k] = w;
    }
    B381_SETRANGE(t, -2.0, 1.0);
  }
  Fp n;
  fp_neg(n, t);
#pragma unroll
  for (int k = 0; k < NL; k++) t.l[k] = (n.l[k] & fneg) | (t.l[k] & ~fneg);
  B381_SETRANGE(t, -2.0, 2.0);
  {                                                     // (-2p, 2p) -> [0, 2p) -> [0, p)
    const uint32_t p2[NL] = B381_P2;
    const uint32_t neg = (uint32_t)((int32_t)t.l[NL - 1] >> 31);
    B381_CC_DECL;
    ADD_CC(t.l[0], t.l[0], p2[0] & neg);
#pragma unroll
    for (int k = 1; k < NL - 1; k++) ADDC_CC(t.l[k], t.l[k], p2[k] & neg);
    ADDC(t.l[NL - 1], t.l[NL - 1], p2[NL - 1] & neg);
    B381_SETRANGE(t, 0.0, 2.0 - 1e-9);
  }
  fp_canon_small(t);
  r = t;
}

}  // namespace b381
