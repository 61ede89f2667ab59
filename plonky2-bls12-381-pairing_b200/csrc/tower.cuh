// tower.cuh -- Fp2/Fp6/Fp12 tower, G2 line functions, Miller loops and final exponentiation
// over memory-resident operands ("slots"), one pairing per thread.
//
// Storage model.  Every per-thread value is an Fp2 "slot" of 24 words (2 x 12: every STORED value is
// below 2^384, asserted by the bound tracker at each store) = 6 uint4 groups.  Group g of a slot
// lives at ptr[g * B381_GS]: on the device B381_GS = block size, so a warp touching the same group
// of its 32 pairings reads 512 contiguous bytes (LDS.128/LDG.128, conflict-free); on the host
// simulation B381_GS = 1.  Slots [0, NS) are in shared memory, slots [NS, NS + nt) in tensor memory,
// the rest in a per-CTA scratch region of global memory (L2 resident: one CTA per SM).
// Fp6 = 3 consecutive slots (c0,c1,c2), Fp12 = 6 consecutive slots
// (c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2) -- the reference's tower order
// (/root/reference/src/fields/helpers.rs:16-37).
//
// Heavy Fp2-level primitives (f2_mul, f2_sqr, ...) are __noinline__: they load their operands
// into registers, do all arithmetic there (fp32.cuh) and store one result.  Everything above is
// orchestration.  Formulas follow the reference's tower twins (cited per function) and, for the
// ARK mode, arkworks 0.4; values are bit-identical to the oracle.
#pragma once
#include "fp32.cuh"

#if defined(__CUDACC__)
typedef uint4 u4;
#define B381_DEV __device__
#define B381_NOINL __device__ __noinline__
#else
struct alignas(16) u4 { uint32_t x, y, z, w; };
#define B381_DEV
#define B381_NOINL __attribute__((noinline))
#endif

#ifndef B381_BLOCK
#define B381_BLOCK 256
#endif
#ifndef B381_SYNC_GROUPS
#define B381_SYNC_GROUPS 1
#endif
#if defined(__CUDA_ARCH__)
#define B381_GS B381_BLOCK
#else
#define B381_GS 1
#endif

namespace b381 {

constexpr int GPS = 6;                 // uint4 groups per slot (2 x 12 words)
constexpr int SW = 12;                 // stored words per Fp
constexpr int SLOT = GPS * B381_GS;    // uint4 stride between slots
#ifndef B381_NS
#define B381_NS 9
#endif
constexpr int NS = B381_NS;            // slots held in shared memory

struct Ctx {
  u4* sm;   // this thread's shared-memory slots
  u4* gm;   // this thread's global-memory slots
  int sync; // 1: the whole CTA runs the same program in lock step -> barrier before every primitive
  uint32_t tm;  // tensor-memory address (lane base << 16 | first column) of this warp's scratch columns
  int nt;       // number of slots held in tensor memory (0 = none); slots [NS, NS + nt)
};

// Tensor memory (TMEM, 256 KB per SM) as a per-thread scratchpad.  The pairing kernels issue no MMA,
// so the whole TMEM of the SM is free: a CTA allocates all 512 columns; warp w owns TMEM lanes
// 32 (w % 4) .. +31 (the only lanes it may address) and columns 256 (w / 4) .. +255, i.e. every
// thread owns 256 words = 10 Fp2 slots of 24 columns, moved with tcgen05.ld/st.32x32b.x16 + .x8 (one thread per lane).
// Pointers into TMEM are tagged (bit 62) so the slot primitives can take either kind.
constexpr int NT_MAX = 10;
constexpr unsigned long long TMEM_TAG = 1ull << 62;

// Warps of a CTA execute the identical instruction stream (control flow does not depend on the
// data).  Left alone they drift apart and each SM sub-partition streams the ~0.5 MB kernel through
// the instruction caches on its own; a barrier before every primitive keeps the four warps within
// one primitive of each other, so one fetch serves all of them (measured: see DESIGN.md).
B381_DEV B381_INL void sync_point(const Ctx& c) {
#if defined(__CUDA_ARCH__)
#if B381_SYNC_GROUPS
  // one named barrier per group of 4 warps (one warp per SM sub-partition): the groups share the
  // code stream loosely but are free to drift against each other, so while one group is in a
  // multiply-heavy stretch the other can be in its carry / load-store stretch.
  if (c.sync) asm volatile("bar.sync %0, 128;" ::"r"(1 + (int)(threadIdx.x >> 7)) : "memory");
#else
  if (c.sync) __syncthreads();
#endif
#else
  (void)c;
#endif
}

B381_DEV B381_INL u4* slot(const Ctx& c, int s) {
  if (s < NS) return c.sm + s * SLOT;
#if defined(__CUDA_ARCH__)
  if (s < NS + c.nt) return reinterpret_cast<u4*>(TMEM_TAG | (unsigned long long)(c.tm + (uint32_t)(s - NS) * 24u));
#endif
  return c.gm + (s - NS) * SLOT;
}

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ bool is_tmem(const u4* p) { return (reinterpret_cast<unsigned long long>(p) & TMEM_TAG) != 0; }
__device__ __forceinline__ void tmem_ld24(uint32_t (&w)[24], uint32_t ta) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
                 "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
               : "r"(ta));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[16]), "=r"(w[17]), "=r"(w[18]), "=r"(w[19]), "=r"(w[20]), "=r"(w[21]), "=r"(w[22]), "=r"(w[23])
               : "r"(ta + 16));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st24(uint32_t ta, const uint32_t (&w)[24]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               :: "r"(ta), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]),
                  "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]), "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15]) : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(ta + 16), "r"(w[16]), "r"(w[17]), "r"(w[18]), "r"(w[19]), "r"(w[20]), "r"(w[21]), "r"(w[22]), "r"(w[23]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
#endif

#ifndef B381_SYNC_LIN
#define B381_SYNC_LIN 1
#endif
B381_DEV B381_INL void sync_point_lin(const Ctx& c) {
#if B381_SYNC_LIN
  sync_point(c);
#else
  (void)c;
#endif
}

// ---------------------------------------------------------------------------------------------
// bound-tracking side table (host simulation only)
// ---------------------------------------------------------------------------------------------
#ifdef B381_TRACK_BOUNDS
}  // namespace b381
#include <unordered_map>
namespace b381 {
struct TrackEnt { double mag[2], lb[2]; };
inline std::unordered_map<const void*, TrackEnt>& track_tab() { static std::unordered_map<const void*, TrackEnt> t; return t; }
inline void track_ld(const u4* p, int h, Fp& a) {
  auto it = track_tab().find(p);
  if (it == track_tab().end()) tb_range(a, 0.0, 1.0); else tb_range(a, it->second.lb[h], it->second.mag[h]);
  a.nonneg = true;
}
inline void track_st(const u4* p, int h, const Fp& a) {
  B381_CHECK(a.lb >= 0, "store of a possibly negative value");
  B381_CHECK(a.mag < 1000.0, "store of oversized value");
  B381_CHECK(a.ub < 9.8, "stored value does not fit 12 words");
  auto& e = track_tab()[p];
  e.mag[h] = a.ub; e.lb[h] = a.lb;
}
#endif

// ---------------------------------------------------------------------------------------------
// slot loads / stores
// ---------------------------------------------------------------------------------------------
B381_DEV B381_INL void ld_f2(Fp& c0, Fp& c1, const u4* p) {
  uint32_t w[2 * SW];
#if defined(__CUDA_ARCH__)
  if (is_tmem(p)) {
    tmem_ld24(w, (uint32_t)reinterpret_cast<unsigned long long>(p));
  } else
#endif
  {
    u4 g[GPS];
#pragma unroll
    for (int i = 0; i < GPS; i++) g[i] = p[i * B381_GS];
#pragma unroll
    for (int i = 0; i < GPS; i++) { w[4 * i] = g[i].x; w[4 * i + 1] = g[i].y; w[4 * i + 2] = g[i].z; w[4 * i + 3] = g[i].w; }
  }
#pragma unroll
  for (int k = 0; k < SW; k++) { c0.l[k] = w[k]; c1.l[k] = w[SW + k]; }
  c0.l[NL - 1] = 0; c1.l[NL - 1] = 0;               // stored values are below 2^384
  B381_TB(track_ld(p, 0, c0); track_ld(p, 1, c1);)
}

B381_DEV B381_INL void st_f2(u4* p, const Fp& c0, const Fp& c1) {
  B381_TB(track_st(p, 0, c0); track_st(p, 1, c1);)
  B381_CHECK(c0.l[NL - 1] == 0 && c1.l[NL - 1] == 0, "stored value has a non-zero thirteenth word");
  uint32_t w[2 * SW];
#pragma unroll
  for (int k = 0; k < SW; k++) { w[k] = c0.l[k]; w[SW + k] = c1.l[k]; }
#if defined(__CUDA_ARCH__)
  if (is_tmem(p)) {
    tmem_st24((uint32_t)reinterpret_cast<unsigned long long>(p), w);
    return;
  }
#endif
#pragma unroll
  for (int i = 0; i < GPS; i++) {
    u4 g;
    g.x = w[4 * i]; g.y = w[4 * i + 1]; g.z = w[4 * i + 2]; g.w = w[4 * i + 3];
    p[i * B381_GS] = g;
  }
}

// one half (h = 0: c0, h = 1: c1) of a slot
B381_DEV B381_INL void ld_fp(Fp& a, const u4* p, int h) {
  Fp c0, c1;
  ld_f2(c0, c1, p);
  if (h) a = c1; else a = c0;
}

// ---------------------------------------------------------------------------------------------
// constants in device constant memory / host statics
// ---------------------------------------------------------------------------------------------
struct ConstTab {
  limb_t frob[3][5][2][NL];   // gamma_k[j] = xi^(j (p^k-1)/6), k = 1..3, j = 1..5, (c0, c1)
  limb_t one[NL];
  uint8_t pm2_nib[96];         // p - 2 as 4-bit windows, MSB first
  limb_t psi[3][2][NL];        // psi: 1 / xi^((p-1)/3), 1 / xi^((p-1)/2); psi^2: 1 / xi^((p^2-1)/3) (in Fp; c1 = 0)
};
#define B381_CONST_INIT { \
  { { {B381_FROB1_1_C0, B381_FROB1_1_C1}, {B381_FROB1_2_C0, B381_FROB1_2_C1}, {B381_FROB1_3_C0, B381_FROB1_3_C1}, \
      {B381_FROB1_4_C0, B381_FROB1_4_C1}, {B381_FROB1_5_C0, B381_FROB1_5_C1} }, \
    { {B381_FROB2_1_C0, B381_FROB2_1_C1}, {B381_FROB2_2_C0, B381_FROB2_2_C1}, {B381_FROB2_3_C0, B381_FROB2_3_C1}, \
      {B381_FROB2_4_C0, B381_FROB2_4_C1}, {B381_FROB2_5_C0, B381_FROB2_5_C1} }, \
    { {B381_FROB3_1_C0, B381_FROB3_1_C1}, {B381_FROB3_2_C0, B381_FROB3_2_C1}, {B381_FROB3_3_C0, B381_FROB3_3_C1}, \
      {B381_FROB3_4_C0, B381_FROB3_4_C1}, {B381_FROB3_5_C0, B381_FROB3_5_C1} } }, \
  B381_ONE, B381_PM2_NIBBLES, \
  { {B381_PSI_CX_C0, B381_PSI_CX_C1}, {B381_PSI_CY_C0, B381_PSI_CY_C1}, {B381_PSI2_CX, {0}} } }

#if defined(__CUDACC__)
static __constant__ ConstTab g_ct = B381_CONST_INIT;
#else
static const ConstTab g_ct = B381_CONST_INIT;
#endif

B381_DEV B381_INL void fp_const(Fp& r, const limb_t* v) {
#pragma unroll
  for (int k = 0; k < NL; k++) r.l[k] = v[k];
  B381_SETRANGE(r, 0.0, 1.0);
}

// ---------------------------------------------------------------------------------------------
// register-level Fp2 helpers
// ---------------------------------------------------------------------------------------------
// B381_W12 (default): the hot primitives multiply 12-word operands (144 instead of 169 IMAD.WIDE per
// product).  Every value they load is a stored value below 5 p and every in-register sum / offset
// difference stays below 2^384 = 9.84 p; the bound tracker asserts it at each multiplication.
#ifndef B381_W12
#define B381_W12 1
#endif
#define HOT_MUL acc_mul12
#define HOT_MAC acc_mac12
#define HOT_MUL3 acc_mul3_12
#define HOT_OFF fp_add_p5
// (r0, r1) = (a0 + a1 u)(b0 + b1 u), Karatsuba in the double-width domain, one reduction per
// coefficient: 3 x 169 + 2 x 156 IMAD.WIDE.  /root/reference/src/fields_as_trees/fq2_target_tree.rs:97-115
B381_DEV B381_INL void f2_mul_reg(Fp& r0, Fp& r1, const Fp& a0, const Fp& a1, const Fp& b0, const Fp& b1) {
  Acc A, B, X;
  Fp sa, sb;
  HOT_MUL(A, a0, b0);
  HOT_MUL(B, a1, b1);
  fp_add(sa, a0, a1);
  fp_add(sb, b0, b1);
  HOT_MUL(X, sa, sb);
  acc_sub(X, X, A);
  acc_sub(X, X, B);                                 // im = (a0 + a1)(b0 + b1) - a0 b0 - a1 b1
  B381_TB(X.cb = 0;)                                // = a0 b1 + a1 b0 >= 0 for non-negative operands
  acc_sub(A, A, B);                                 // re = a0 b0 - a1 b1
  acc_redc2(r0, A, r1, X);
  fp_add_p(r0, r0);                                 // re > -p/128 -> stored values stay non-negative
}

// ((a0+a1)(a0-a1), 2 a0 a1); fq2_target_tree.rs:80-91.  2 x 169 + 2 x 156 IMAD.WIDE.
B381_DEV B381_INL void f2_sqr_reg(Fp& r0, Fp& r1, const Fp& a0, const Fp& a1) {
  Fp s, d, t;
  fp_add(s, a0, a1);
  fp_sub(d, a0, a1);
  HOT_OFF(d, d);                                // a0 - a1 + 128 p >= 0 (same residue)
  fp_dbl(t, a0);
  Acc T, U;
  HOT_MUL(T, s, d);
  HOT_MUL(U, t, a1);
  acc_redc2(r0, T, r1, U);
}

// (a0 s, a1 s) for an Fp scalar s
B381_DEV B381_INL void f2_mulfp_reg(Fp& r0, Fp& r1, const Fp& a0, const Fp& a1, const Fp& s) {
  Acc T, U;
  HOT_MUL(T, a0, s);
  HOT_MUL(U, a1, s);
  acc_redc2(r0, T, r1, U);
}

B381_DEV B381_INL void f2_norm(Fp& a0, Fp& a1) { fp_norm(a0); fp_norm(a1); }

// xi * (a0 + a1 u) = (a0 - a1) + (a0 + a1) u ; fq2_target_tree.rs:137-142
B381_DEV B381_INL void f2_mulxi_reg(Fp& r0, Fp& r1, const Fp& a0, const Fp& a1) {
  Fp t0, t1;
  fp_sub(t0, a0, a1);
  fp_add(t1, a0, a1);
  r0 = t0; r1 = t1;
}

// Fp inversion: Bernstein-Yang division steps (fp32.cuh fp_inv_safegcd); the Fermat power a^(p-2) it replaces
// cost 4 % of a pairing.  B381_FERMAT_INV selects the old square-and-multiply ladder (kept for measurements).
#ifdef B381_FERMAT_INV
#define FP_MUL_SMALL fp_mul12      // x and a are reduction outputs (at most 1.03 p): 12-word products
B381_DEV B381_INL void fp_inv_reg(Fp& r, const Fp& a) {
  Fp x;
  fp_const(x, g_ct.one);
  for (int i = 0; i < 96; i++) {
    int nib = g_ct.pm2_nib[i];
    for (int b = 3; b >= 0; b--) {
      Fp t;
      FP_MUL_SMALL(t, x, x);
      x = t;
      if ((nib >> b) & 1) {
        FP_MUL_SMALL(t, x, a);
        x = t;
      }
    }
  }
  r = x;
}
#else
B381_DEV B381_INL void fp_inv_reg(Fp& r, const Fp& a) { fp_inv_safegcd(r, a); }
#endif

// ---------------------------------------------------------------------------------------------
// noinline slot primitives
// ---------------------------------------------------------------------------------------------
// Pre- / post-operations of the multiplication primitives.  A linear operation on a slot (load, a few carry
// chains, weak reduction, store, plus a call and a barrier) costs about a twelfth of a three-product sum while
// using the multiplier pipe hardly at all, so the hot paths (ARK Miller loop, Fp12 products) fold their
// additions, differences, halvings, triplings and xi-multiplications into the operand loads and result
// stores of the multiplications they feed.  All bounds (operands below 2^384 and non-negative, results
// non-negative) are asserted by the bound tracker on the host simulation.
enum PreOp {
  PRE_NONE = 0,
  PRE_ADD,        // a + a2, weak-reduced (any two stored values)
  PRE_ADD_RAW,    // a + a2 as is (caller guarantees the sums stay below what the product takes)
  PRE_SUB3P,      // a - a2 + 3 p          (a2 at most 3 p)
  PRE_HALFSUM,    // (a + a2) / 2
  PRE_NEG3P,      // 3 p - a               (a at most 3 p)
};
enum PostOp {
  POST_NONE = 0,
  POST_HALF,      // r / 2
  POST_TRIPLE,    // 3 r   (below 6.1 p: feeds an Fp-scalar product or a POST_SUB1)
  POST_SUB1,      // r - p1, weak-reduced
  POST_SUB2,      // r - p1 - p2, weak-reduced
};

B381_DEV B381_INL void f2_pre(Fp& a0, Fp& a1, const u4* a2, int pre) {
  if (pre == PRE_NONE) return;
  if (pre == PRE_NEG3P) {
    fp_neg(a0, a0); fp_neg(a1, a1);
    fp_add_p3(a0, a0); fp_add_p3(a1, a1);
    return;
  }
  Fp t0, t1;
  ld_f2(t0, t1, a2);
  if (pre == PRE_SUB3P) {
    fp_sub(a0, a0, t0); fp_sub(a1, a1, t1);
    fp_add_p3(a0, a0); fp_add_p3(a1, a1);
    return;
  }
  fp_add(a0, a0, t0); fp_add(a1, a1, t1);
  if (pre == PRE_ADD) { fp_wreduce(a0); fp_wreduce(a1); }
  if (pre == PRE_HALFSUM) { fp_half(a0, a0); fp_half(a1, a1); }
}

B381_DEV B381_INL void f2_post(Fp& r0, Fp& r1, int post, const u4* p1, const u4* p2) {
  if (post == POST_NONE) return;
  if (post == POST_HALF) { fp_half(r0, r0); fp_half(r1, r1); return; }
  if (post == POST_TRIPLE) {
    Fp t;
    fp_dbl(t, r0); fp_add(r0, r0, t);
    fp_dbl(t, r1); fp_add(r1, r1, t);
    return;
  }
  Fp x0, x1;
  ld_f2(x0, x1, p1);
  fp_sub(r0, r0, x0); fp_sub(r1, r1, x1);
  if (post == POST_SUB2) {
    ld_f2(x0, x1, p2);
    fp_sub(r0, r0, x0); fp_sub(r1, r1, x1);
  }
  fp_wreduce(r0); fp_wreduce(r1);
}

// r = post( pre_a(a, a2) * pre_b(b, b2) )
B381_NOINL void f2_mul_ex(u4* r, const u4* a, const u4* a2, int pre_a, const u4* b, const u4* b2, int pre_b, int post, const u4* p1, const u4* p2) {
  Fp a0, a1, b0, b1, r0, r1;
  ld_f2(a0, a1, a);
  f2_pre(a0, a1, a2, pre_a);
  ld_f2(b0, b1, b);
  f2_pre(b0, b1, b2, pre_b);
  f2_mul_reg(r0, r1, a0, a1, b0, b1);
  f2_post(r0, r1, post, p1, p2);
  st_f2(r, r0, r1);
}
B381_DEV B381_INL void f2_mul(u4* r, const u4* a, const u4* b) { f2_mul_ex(r, a, nullptr, PRE_NONE, b, nullptr, PRE_NONE, POST_NONE, nullptr, nullptr); }
// r = (a + a2) * (b + b2); a2 / b2 may be null
B381_DEV B381_INL void f2_mul_ss(u4* r, const u4* a, const u4* a2, const u4* b, const u4* b2) {
  f2_mul_ex(r, a, a2, a2 ? PRE_ADD : PRE_NONE, b, b2, b2 ? PRE_ADD : PRE_NONE, POST_NONE, nullptr, nullptr);
}

// Sum of three products with ONE reduction per coefficient:  r = a0 b0 + a1 b1 + a2 b2  in Fp2.
// Karatsuba over the SUMS: P = sum a_i0 b_i0, Q = sum a_i1 b_i1, X = sum (a_i0 + a_i1)(b_i0 + b_i1), each one
// fused three-product accumulation (acc_mul3); re = P - Q (+ p), im = X - P - Q: 9 MAC blocks + 2 reductions
// instead of 3 Fp2 multiplications (9 blocks + 6 reductions) followed by memory-to-memory additions -- the
// row-serial Montgomery reduction is the expensive half of a multiplication, so the
// tower formulas are arranged as sums of products.
// flags: SOP_XIk multiplies b_k by xi = 1 + u while it sits in registers, (b0 - b1 + 3 p, b0 + b1): b_k must be a
// stored value with b0 < 3.4 p and b1 <= 3 p (tracker-asserted through the 12-word products); SOP_DBL doubles
// the result in the double-width domain (2 a b of the Fp12 squaring).
// Result post-operations (the Karatsuba recombinations of the Fp12 product / squaring, folded into the sum that
// produces their first operand; p1 / p2 / r2 are extra slots):
//   SOP_SUB2        r = S - p1 - p2                    weak-reduced
//   SOP_HALFSUB     r = S - (p1 + p2) / 2              weak-reduced      (SOP_HALFSUB_XI: xi p2)
//   SOP_ALSO_ADD    r = S  and  r2 = S + p1            r2 weak-reduced   (SOP_ALSO_XIADD: r2 = xi S + p1)
enum SopFlags { SOP_XI0 = 1, SOP_XI1 = 2, SOP_XI2 = 4, SOP_DBL = 8, SOP_SUB2 = 16, SOP_HALFSUB = 32, SOP_HALFSUB_XI = 64, SOP_ALSO_ADD = 128, SOP_ALSO_XIADD = 256 };
B381_DEV B381_INL void f2_xi_pos(Fp& b0, Fp& b1) {
  Fp d;
  fp_sub(d, b0, b1);
  fp_add(b1, b0, b1);
  fp_add_p3(b0, d);
}
// Two instances: the lean one (operand flags only) is the Miller loop's; the one with result post-operations is
// used by the Fp12 product.  One function with everything cost the Miller loop 2.6 % (measured).
template <bool POST>
B381_DEV B381_INL void f2_sop_t(u4* r, int flags, const u4* a0p, const u4* b0p, const u4* a1p, const u4* b1p, const u4* a2p, const u4* b2p,
                                const u4* p1, const u4* p2, u4* r2) {
  Fp x0, x1, y0, y1, z0, z1, u0, u1, v0, v1, w0, w1;
  ld_f2(x0, x1, a0p); ld_f2(u0, u1, b0p);
  ld_f2(y0, y1, a1p); ld_f2(v0, v1, b1p);
  ld_f2(z0, z1, a2p); ld_f2(w0, w1, b2p);
  if (flags & SOP_XI0) f2_xi_pos(u0, u1);
  if (flags & SOP_XI1) f2_xi_pos(v0, v1);
  if (flags & SOP_XI2) f2_xi_pos(w0, w1);
  Acc P, Q, X;
  HOT_MUL3(P, x0, u0, y0, v0, z0, w0);
  HOT_MUL3(Q, x1, u1, y1, v1, z1, w1);
  fp_add(x0, x0, x1); fp_add(u0, u0, u1);
  fp_add(y0, y0, y1); fp_add(v0, v0, v1);
  fp_add(z0, z0, z1); fp_add(w0, w0, w1);
  HOT_MUL3(X, x0, u0, y0, v0, z0, w0);
  acc_sub(X, X, P);
  acc_sub(X, X, Q);                                 // im = sum (a_i0 b_i1 + a_i1 b_i0) >= 0
  B381_TB(X.cb = 0;)
  acc_sub(P, P, Q);                                 // re
  if (flags & SOP_DBL) { acc_dbl(P); acc_dbl(X); }
  Fp r0, r1;
  acc_redc2(r0, P, r1, X);
  fp_add_p(r0, r0);
  if (POST && (flags & (SOP_SUB2 | SOP_HALFSUB | SOP_HALFSUB_XI))) {
    Fp x0, x1, y0, y1;
    ld_f2(x0, x1, p1);
    ld_f2(y0, y1, p2);
    if (flags & SOP_SUB2) {
      fp_sub(r0, r0, x0); fp_sub(r1, r1, x1);
      fp_sub(r0, r0, y0); fp_sub(r1, r1, y1);
    } else {
      if (flags & SOP_HALFSUB_XI) f2_mulxi_reg(y0, y1, y0, y1);
      fp_add(x0, x0, y0); fp_add(x1, x1, y1);
      fp_half(x0, x0); fp_half(x1, x1);
      fp_sub(r0, r0, x0); fp_sub(r1, r1, x1);
    }
    fp_wreduce(r0); fp_wreduce(r1);
  }
  st_f2(r, r0, r1);
  if (POST && (flags & (SOP_ALSO_ADD | SOP_ALSO_XIADD))) {
    Fp x0, x1;
    ld_f2(x0, x1, p1);
    if (flags & SOP_ALSO_XIADD) f2_mulxi_reg(r0, r1, r0, r1);
    fp_add(r0, r0, x0); fp_add(r1, r1, x1);
    fp_wreduce(r0); fp_wreduce(r1);
    st_f2(r2, r0, r1);
  }
}
B381_NOINL void f2_sop(u4* r, int flags, const u4* a0p, const u4* b0p, const u4* a1p, const u4* b1p, const u4* a2p, const u4* b2p) {
  f2_sop_t<false>(r, flags, a0p, b0p, a1p, b1p, a2p, b2p, nullptr, nullptr, nullptr);
}
B381_NOINL void f2_sop_post(u4* r, int flags, const u4* a0p, const u4* b0p, const u4* a1p, const u4* b1p, const u4* a2p, const u4* b2p,
                            const u4* p1, const u4* p2, u4* r2) {
  f2_sop_t<true>(r, flags, a0p, b0p, a1p, b1p, a2p, b2p, p1, p2, r2);
}

// r = post( pre(a, a2)^2 )
B381_NOINL void f2_sqr_ex(u4* r, const u4* a, const u4* a2, int pre, int post, const u4* p1, const u4* p2) {
  Fp a0, a1, r0, r1;
  ld_f2(a0, a1, a);
  f2_pre(a0, a1, a2, pre);
  f2_sqr_reg(r0, r1, a0, a1);
  f2_post(r0, r1, post, p1, p2);
  st_f2(r, r0, r1);
}
// r = (a + a2)^2 ; a2 may be null
B381_DEV B381_INL void f2_sqr(u4* r, const u4* a, const u4* a2) { f2_sqr_ex(r, a, a2, a2 ? PRE_ADD : PRE_NONE, POST_NONE, nullptr, nullptr); }

// r = (neg ? 3 p - a : a) * s where s is half h of slot sp (an Fp scalar)
B381_NOINL void f2_mulfp(u4* r, const u4* a, const u4* sp, int h, int neg = 0) {
  Fp a0, a1, s, r0, r1;
  ld_f2(a0, a1, a);
  if (neg) f2_pre(a0, a1, nullptr, PRE_NEG3P);
  ld_fp(s, sp, h);
  f2_mulfp_reg(r0, r1, a0, a1, s);
  st_f2(r, r0, r1);
}

// r = frob-coefficient multiply: (conj? conj(a) : a) * gamma_k[j]   (k in 1..3, j in 1..5)
B381_NOINL void f2_mul_gamma(u4* r, const u4* a, int k, int j, int conj) {
  Fp a0, a1, g0, g1, r0, r1;
  ld_f2(a0, a1, a);
  if (conj) {
    fp_neg(a1, a1); fp_norm(a1);
    HOT_OFF(a1, a1);                                // -a1 + 5 p (or 128 p): non-negative, same residue
  }
  fp_const(g0, g_ct.frob[k - 1][j - 1][0]);
  fp_const(g1, g_ct.frob[k - 1][j - 1][1]);
  if (k == 2) {                     // gamma_2[j] lies in Fp
    f2_mulfp_reg(r0, r1, a0, a1, g0);
  } else {
    f2_mul_reg(r0, r1, a0, a1, g0, g1);
  }
  st_f2(r, r0, r1);
}

// r = (conj ? conj(a) : a) * psi-constant idx (0: psi x, 1: psi y, 2: psi^2 x, a real constant)
// -- the G2 endomorphisms of ark-bls12-381 g2.rs p_power_endomorphism / double_p_power_endomorphism
B381_NOINL void f2_mul_psi(u4* r, const u4* a, int idx, int conj) {
  Fp a0, a1, g0, g1, r0, r1;
  ld_f2(a0, a1, a);
  if (conj) { fp_neg(a1, a1); HOT_OFF(a1, a1); }
  fp_const(g0, g_ct.psi[idx][0]);
  fp_const(g1, g_ct.psi[idx][1]);
  if (idx == 2) f2_mulfp_reg(r0, r1, a0, a1, g0);
  else f2_mul_reg(r0, r1, a0, a1, g0, g1);
  st_f2(r, r0, r1);
}

// r = 1 / a ; (a0, -a1)/(a0^2 + a1^2) ; fq2_target_tree.rs:66-78.  a = 0 gives 0.
B381_NOINL void f2_inv(u4* r, const u4* a) {
  Fp a0, a1, n, ni, r0, r1;
  ld_f2(a0, a1, a);
  Acc T;
  acc_zero(T);
  acc_mac(T, a0, a0);
  acc_mac(T, a1, a1);
  acc_redc(n, T);
  fp_inv_reg(ni, n);
  fp_mul(r0, a0, ni);
  fp_neg(a1, a1);
  fp_norm(a1);
  fp_add_p128(a1, a1);
  fp_mul(r1, a1, ni);
  st_f2(r, r0, r1);
}

// two-operand linear ops
enum LinOp {
  L_ADD = 0,      // a + b
  L_SUB,          // a - b
  L_NEG,          // -a
  L_DBL,          // 2a
  L_TRIPLE,       // 3a
  L_MULXI,        // xi a
  L_CONJ,         // conj(a)
  L_COPY,         // a
  L_HALF,         // a / 2
  L_HALFSUM,      // (a + b) / 2
  L_XIADD,        // a + xi b
  L_3A_M2B,       // 3a - 2b
  L_3A_P2B,       // 3a + 2b
  L_MUL12XI,      // 12 xi a          (ark-ec g2.rs: COEFF_B * 3c with COEFF_B = 4 xi)
  L_MUL8,         // 8a
  L_2A_MB,        // 2a - b
  L_MUL4,         // 4a
  L_ADD_R,        // a + b, weak-reduced (results that feed the cyclotomic squarings must be below 1.1 p)
};

B381_NOINL void f2_lin(u4* r, const u4* a, const u4* b, int op) {
  Fp a0, a1, b0, b1, r0, r1;
  ld_f2(a0, a1, a);
  if (b) ld_f2(b0, b1, b); else { fp_zero(b0); fp_zero(b1); }
  switch (op) {
    case L_ADD: case L_ADD_R: fp_add(r0, a0, b0); fp_add(r1, a1, b1); break;
    case L_SUB: fp_sub(r0, a0, b0); fp_sub(r1, a1, b1); break;
    case L_NEG: fp_neg(r0, a0); fp_neg(r1, a1); break;
    case L_DBL: fp_dbl(r0, a0); fp_dbl(r1, a1); break;
    case L_TRIPLE: fp_dbl(r0, a0); fp_add(r0, r0, a0); fp_dbl(r1, a1); fp_add(r1, r1, a1); break;
    case L_MULXI: f2_mulxi_reg(r0, r1, a0, a1); break;
    case L_CONJ: r0 = a0; fp_neg(r1, a1); break;
    case L_COPY: r0 = a0; r1 = a1; break;
    case L_HALF: fp_half(r0, a0); fp_half(r1, a1); break;
    case L_HALFSUM:
      fp_add(r0, a0, b0); fp_add(r1, a1, b1);
      f2_norm(r0, r1);
      fp_half(r0, r0); fp_half(r1, r1);
      break;
    case L_XIADD: {
      Fp t0, t1;
      f2_mulxi_reg(t0, t1, b0, b1);
      fp_add(r0, a0, t0); fp_add(r1, a1, t1);
    } break;
    case L_3A_M2B:
      fp_sub(r0, a0, b0); fp_dbl(r0, r0); fp_add(r0, r0, a0);
      fp_sub(r1, a1, b1); fp_dbl(r1, r1); fp_add(r1, r1, a1);
      break;
    case L_3A_P2B:
      fp_add(r0, a0, b0); fp_dbl(r0, r0); fp_add(r0, r0, a0);
      fp_add(r1, a1, b1); fp_dbl(r1, r1); fp_add(r1, r1, a1);
      break;
    case L_MUL12XI: {
      Fp t0, t1;
      fp_dbl(t0, a0); fp_add(t0, t0, a0); fp_dbl(t1, a1); fp_add(t1, t1, a1);   // 3a
      f2_norm(t0, t1);
      fp_dbl(t0, t0); fp_dbl(t0, t0); fp_dbl(t1, t1); fp_dbl(t1, t1);           // 12a
      f2_norm(t0, t1);
      f2_mulxi_reg(r0, r1, t0, t1);
    } break;
    case L_MUL8:
      fp_dbl(r0, a0); fp_dbl(r0, r0); fp_dbl(r1, a1); fp_dbl(r1, r1);
      f2_norm(r0, r1);
      fp_dbl(r0, r0); fp_dbl(r1, r1);
      break;
    case L_2A_MB:
      fp_dbl(r0, a0); fp_sub(r0, r0, b0); fp_dbl(r1, a1); fp_sub(r1, r1, b1);
      break;
    default:  // L_MUL4
      fp_dbl(r0, a0); fp_dbl(r0, r0); fp_dbl(r1, a1); fp_dbl(r1, r1);
      break;
  }
  // stored values must be non-negative: the weak reduction (output in [0, 11 p)) where a difference occurs
  // ... and (B381_W12) wherever the result could exceed 5 p, so that it fits the 12-word multiplications
  const bool grow = (B381_W12 && (op == L_TRIPLE || op == L_MUL4 || op == L_MUL8 || op == L_MUL12XI || op == L_XIADD)) || op == L_ADD_R;
  const bool n0 = grow || op == L_SUB || op == L_NEG || op == L_MULXI || op == L_XIADD || op == L_3A_M2B || op == L_3A_P2B || op == L_MUL12XI || op == L_2A_MB;
  const bool n1 = grow || op == L_SUB || op == L_NEG || op == L_CONJ || op == L_3A_M2B || op == L_3A_P2B || op == L_2A_MB;
  if (n0) fp_wreduce(r0);
  if (n1) fp_wreduce(r1);
  st_f2(r, r0, r1);
}

// Karatsuba recombination: t = a - b - c (b, c optional); then
//   K_PLAIN     r = t + d
//   K_XI_INNER  r = xi t + d
//   K_XI_D      r = t + xi d
//   K_XI_C      r = a - b - xi c  (+ d)
//   K_HALFSUB    r = a - (b + c) / 2        (Fp12 squaring with 2ab in hand: c0 = s u - (2ab + v 2ab) / 2)
//   K_HALFSUB_XI r = a - (b + xi c) / 2
enum KOp { K_PLAIN = 0, K_XI_INNER, K_XI_D, K_XI_C, K_HALFSUB, K_HALFSUB_XI };

B381_NOINL void f2_kcomb(u4* r, const u4* a, const u4* b, const u4* c, const u4* d, int mode) {
  Fp t0, t1, x0, x1;
  ld_f2(t0, t1, a);
  if (mode == K_HALFSUB || mode == K_HALFSUB_XI) {
    Fp y0, y1;
    ld_f2(x0, x1, b);
    ld_f2(y0, y1, c);
    if (mode == K_HALFSUB_XI) f2_mulxi_reg(y0, y1, y0, y1);
    fp_add(x0, x0, y0); fp_add(x1, x1, y1);
    fp_half(x0, x0); fp_half(x1, x1);
    fp_sub(t0, t0, x0); fp_sub(t1, t1, x1);
    fp_wreduce(t0); fp_wreduce(t1);
    st_f2(r, t0, t1);
    return;
  }
  if (b) { ld_f2(x0, x1, b); fp_sub(t0, t0, x0); fp_sub(t1, t1, x1); }
  if (c) {
    ld_f2(x0, x1, c);
    if (mode == K_XI_C) f2_mulxi_reg(x0, x1, x0, x1);
    fp_sub(t0, t0, x0); fp_sub(t1, t1, x1);
  }
  if (mode == K_XI_INNER) { f2_norm(t0, t1); f2_mulxi_reg(t0, t1, t0, t1); }
  if (d) {
    ld_f2(x0, x1, d);
    if (mode == K_XI_D) f2_mulxi_reg(x0, x1, x0, x1);
    fp_add(t0, t0, x0); fp_add(t1, t1, x1);
  }
  f2_norm(t0, t1);
  fp_wreduce(t0); fp_wreduce(t1);
  st_f2(r, t0, t1);
}

// The three linear values of the ARK doubling step (ark-ec g2.rs double_in_place) in one pass:
//   e = 12 xi c  (= COEFF_B * 3c),   i = e - b,   f = 3 e      all weak-reduced
B381_NOINL void f2_dbl_lin3(u4* re, u4* ri, u4* rf, const u4* c, const u4* b) {
  Fp c0, c1, b0, b1, t0, t1, e0, e1, x0, x1;
  ld_f2(c0, c1, c);
  ld_f2(b0, b1, b);
  fp_dbl(t0, c0); fp_add(t0, t0, c0); fp_dbl(t1, c1); fp_add(t1, t1, c1);   // 3c
  fp_dbl(t0, t0); fp_dbl(t0, t0); fp_dbl(t1, t1); fp_dbl(t1, t1);           // 12c
  f2_mulxi_reg(e0, e1, t0, t1);
  fp_wreduce(e0); fp_wreduce(e1);
  st_f2(re, e0, e1);
  fp_sub(x0, e0, b0); fp_sub(x1, e1, b1);
  fp_wreduce(x0); fp_wreduce(x1);
  st_f2(ri, x0, x1);
  fp_dbl(x0, e0); fp_add(x0, x0, e0); fp_dbl(x1, e1); fp_add(x1, x1, e1);
  fp_wreduce(x0); fp_wreduce(x1);
  st_f2(rf, x0, x1);
}

// Fused step of the Granger-Scott cyclotomic squaring (/root/reference/src/fields_as_trees/miller_loop.rs:29-104):
//   (t0, t1) = fp4_square(a, b) = (a^2 + xi b^2, 2 a b)
//   mode 0:  ra = 3 t0 - 2 za ;      rb = 3 t1 + 2 zb
//   mode 1:  ra = 3 xi t1 + 2 za ;   rb = 3 t0 - 2 zb
// a^2 + xi b^2 is accumulated in the column domain (4 MAC blocks, one reduction per coefficient)
// and 2ab is one Karatsuba product: 7 x 196 + 4 x 225 + 4 x 15 IMAD instead of three separate
// squarings (6 x 196 + 6 x 225) plus seven memory-to-memory linear operations.  The linear feedback
// of z (magnitude M -> 9 + 2M) is absorbed by the weak reduction when `reduce` is set; callers set it
// at least every sixth squaring (2 -> 13 -> 35 -> 79 -> 167 -> 343 stays far below the 14-limb range).
// Six-product form: with A = a^2 = (PA, QA), B = b^2 = (PB, QB), C = (a + b)^2 = (PC, QC), where
// P = (x0 + x1)(x0 - x1 + 5p) and Q = 2 x0 x1 for x = (x0, x1):
//   t0 = a^2 + xi b^2 = (PA + PB - QB, QA + PB + QB);   t1 = 2ab = (PC - PA - PB, QC - QA - QB)
// 6 x 144 + 4 x 156 IMAD.WIDE instead of 7 x 144 + 4 x 156 (the a b product of the seven-product form is
// replaced by one squaring), and only four double-width values are ever live.  Operands must be
// below 1.1 p each (weak-reduced), which every caller guarantees (tracker-asserted through the
// 12-word multiplications).
B381_NOINL void f2_cyc_fp4(u4* ra, u4* rb, const u4* a, const u4* b, const u4* za, const u4* zb, int mode, int reduce) {
  (void)reduce;
  Fp a0, a1, b0, b1, t00, t01, t10, t11;
  ld_f2(a0, a1, a);
  ld_f2(b0, b1, b);
  Acc SP, SQ;
  {
    Fp s, d, e;
    Acc PB, QB;
    fp_add(s, a0, a1); fp_sub(d, a0, a1); HOT_OFF(d, d);
    HOT_MUL(SP, s, d);                              // PA
    fp_dbl(e, a0);
    HOT_MUL(SQ, e, a1);                             // QA
    fp_add(s, b0, b1); fp_sub(d, b0, b1); HOT_OFF(d, d);
    HOT_MUL(PB, s, d);                              // PB
    fp_dbl(e, b0);
    HOT_MUL(QB, e, b1);                             // QB
    acc_add(SP, SP, PB);                            // PA + PB
    acc_add(SQ, SQ, QB);                            // QA + QB
    acc_sub(QB, SP, QB);                            // X = PA + PB - QB
    acc_add(PB, SQ, PB);                            // Y = QA + QB + PB
    acc_redc2(t00, QB, t01, PB);
    fp_add_p(t00, t00);                             // X may be negative
  }
  Fp z0, z1;
  {
    const u4* zt0 = mode == 0 ? za : zb;
    if (zt0 == a) { z0 = a0; z1 = a1; } else ld_f2(z0, z1, zt0);
    Fp w0, w1;
    fp_sub(w0, t00, z0); fp_dbl(w0, w0); fp_add(w0, w0, t00);
    fp_sub(w1, t01, z1); fp_dbl(w1, w1); fp_add(w1, w1, t01);
    fp_wreduce(w0); fp_wreduce(w1);
    st_f2(mode == 0 ? ra : rb, w0, w1);
  }
  const u4* zt1 = mode == 0 ? zb : za;
  if (zt1 == b) { z0 = b0; z1 = b1; } else ld_f2(z0, z1, zt1);
  {
    Fp c0, c1, s, d, e;
    Acc PC, QC;
    fp_add(c0, a0, b0); fp_add(c1, a1, b1);
    fp_add(s, c0, c1); fp_sub(d, c0, c1); HOT_OFF(d, d);
    HOT_MUL(PC, s, d);
    fp_dbl(e, c0);
    HOT_MUL(QC, e, c1);
    acc_sub(PC, PC, SP);                            // re(2ab) = PC - PA - PB
    acc_sub(QC, QC, SQ);                            // im(2ab) = QC - QA - QB = 2 (a0 b1 + a1 b0) >= 0
    B381_TB(QC.cb = 0;)
    acc_redc2(t10, PC, t11, QC);
    fp_add_p(t10, t10);
  }
  if (mode == 1) f2_mulxi_reg(t10, t11, t10, t11);
  // 3 t1 + 2 z
  Fp x0, x1;
  fp_dbl(x0, t10); fp_add(x0, x0, t10); fp_add(x0, x0, z0); fp_add(x0, x0, z0);
  fp_dbl(x1, t11); fp_add(x1, x1, t11); fp_add(x1, x1, z1); fp_add(x1, x1, z1);
  fp_wreduce(x0); fp_wreduce(x1);
  st_f2(mode == 0 ? rb : ra, x0, x1);
}

// set slot to the Fp2 constant (one, 0) or (0, 0)
B381_NOINL void f2_set_small(u4* r, int one) {
  Fp c0, c1;
  fp_zero(c0); fp_zero(c1);
  if (one) fp_const(c0, g_ct.one);
  B381_SETRANGE(c0, 0.0, 1.0); B381_SETRANGE(c1, 0.0, 1.0);
  st_f2(r, c0, c1);
}

// r = cond ? (one ? 1 : 0) : r, executed by EVERY thread (the load and the store are warp-uniform,
// only the selected value is per-thread): slots may live in tensor memory, whose tcgen05.ld/st are
// .sync.aligned and must never sit under a divergent branch.
B381_NOINL void f2_override_if(u4* r, int cond, int one) {
  Fp c0, c1, k0, k1;
  ld_f2(c0, c1, r);
  fp_zero(k0); fp_zero(k1);
  if (one) fp_const(k0, g_ct.one);
#pragma unroll
  for (int i = 0; i < NL; i++) {
    c0.l[i] = cond ? k0.l[i] : c0.l[i];
    c1.l[i] = cond ? k1.l[i] : c1.l[i];
  }
  if (cond) { B381_SETRANGE(c0, 0.0, 1.0); B381_SETRANGE(c1, 0.0, 1.0); }
  st_f2(r, c0, c1);
}

// r = cond ? a : r, executed by every thread of the warp (uniform loads / stores, per-thread selection): see above
B381_NOINL void f2_select_if(u4* r, const u4* a, int cond) {
  Fp c0, c1, k0, k1;
  ld_f2(c0, c1, r);
  ld_f2(k0, k1, a);
#pragma unroll
  for (int i = 0; i < NL; i++) {
    c0.l[i] = cond ? k0.l[i] : c0.l[i];
    c1.l[i] = cond ? k1.l[i] : c1.l[i];
  }
#ifdef B381_TRACK_BOUNDS
  if (cond) { c0.lb = k0.lb; c0.ub = k0.ub; c0.mag = k0.mag; c1.lb = k1.lb; c1.ub = k1.ub; c1.mag = k1.mag; }
#endif
  st_f2(r, c0, c1);
}

// canonical test a == 0 (full reduction; rare path)
B381_NOINL bool f2_is_zero(const u4* a) {
  Fp a0, a1;
  ld_f2(a0, a1, a);
  fp_canon(a0); fp_canon(a1);
  return fp_is_zero_canon(a0) && fp_is_zero_canon(a1);
}

B381_NOINL bool f2_equal(const u4* a, const u4* b) {
  Fp a0, a1, b0, b1;
  ld_f2(a0, a1, a);
  ld_f2(b0, b1, b);
  fp_sub(a0, a0, b0); fp_sub(a1, a1, b1);
  f2_norm(a0, a1);
  fp_canon(a0); fp_canon(a1);
  return fp_is_zero_canon(a0) && fp_is_zero_canon(a1);
}

// external (12 x u32 Montgomery R=2^384) <-> slot.  src = 24 words (c0, c1).  Returns validity.
// External buffers are read / written with 128-bit accesses on the device (every Fq2 of the C-ABI layouts starts
// on a 16-byte boundary when the buffer does: 24-word elements).
B381_DEV B381_INL void ext_ld24(uint32_t (&w0)[12], uint32_t (&w1)[12], const uint32_t* src) {
#if defined(__CUDA_ARCH__)
  if ((reinterpret_cast<unsigned long long>(src) & 15ull) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    const uint4 v0 = s4[0], v1 = s4[1], v2 = s4[2], v3 = s4[3], v4 = s4[4], v5 = s4[5];
    w0[0] = v0.x; w0[1] = v0.y; w0[2] = v0.z; w0[3] = v0.w; w0[4] = v1.x; w0[5] = v1.y; w0[6] = v1.z; w0[7] = v1.w;
    w0[8] = v2.x; w0[9] = v2.y; w0[10] = v2.z; w0[11] = v2.w;
    w1[0] = v3.x; w1[1] = v3.y; w1[2] = v3.z; w1[3] = v3.w; w1[4] = v4.x; w1[5] = v4.y; w1[6] = v4.z; w1[7] = v4.w;
    w1[8] = v5.x; w1[9] = v5.y; w1[10] = v5.z; w1[11] = v5.w;
    return;
  }
#endif
#pragma unroll
  for (int i = 0; i < 12; i++) { w0[i] = src[i]; w1[i] = src[12 + i]; }
}
B381_DEV B381_INL void ext_st24(uint32_t* dst, const uint32_t (&w0)[12], const uint32_t (&w1)[12]) {
#if defined(__CUDA_ARCH__)
  if ((reinterpret_cast<unsigned long long>(dst) & 15ull) == 0) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    d4[0] = make_uint4(w0[0], w0[1], w0[2], w0[3]); d4[1] = make_uint4(w0[4], w0[5], w0[6], w0[7]); d4[2] = make_uint4(w0[8], w0[9], w0[10], w0[11]);
    d4[3] = make_uint4(w1[0], w1[1], w1[2], w1[3]); d4[4] = make_uint4(w1[4], w1[5], w1[6], w1[7]); d4[5] = make_uint4(w1[8], w1[9], w1[10], w1[11]);
    return;
  }
#endif
#pragma unroll
  for (int i = 0; i < 12; i++) { dst[i] = w0[i]; dst[12 + i] = w1[i]; }
}

B381_NOINL bool f2_load_ext(u4* r, const uint32_t* src) {
  uint32_t w0[12], w1[12];
  ext_ld24(w0, w1, src);
  Fp c0, c1;
  bool ok0 = fp_from_ext(c0, w0);
  bool ok1 = fp_from_ext(c1, w1);
  st_f2(r, c0, c1);
  return ok0 && ok1;
}

B381_NOINL void f2_store_ext(uint32_t* dst, const u4* a) {
  Fp c0, c1;
  ld_f2(c0, c1, a);
  uint32_t w0[12], w1[12];
  fp_to_ext(w0, c0);
  fp_to_ext(w1, c1);
  ext_st24(dst, w0, w1);
}

// Streaming products (config #2): operands taken as they are (see fp_from_ext_asis), the two halves from separate
// places (tower order: s1 = s0 + 12; w-basis order: wherever the coefficient map puts them), and the matching store
// that removes the 2^-64 of a product of two such operands.
B381_DEV B381_INL void ext_ld12(uint32_t (&w)[12], const uint32_t* src) {
#if defined(__CUDA_ARCH__)
  if ((reinterpret_cast<unsigned long long>(src) & 15ull) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    const uint4 v0 = s4[0], v1 = s4[1], v2 = s4[2];
    w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
    w[8] = v2.x; w[9] = v2.y; w[10] = v2.z; w[11] = v2.w;
    return;
  }
#endif
#pragma unroll
  for (int i = 0; i < 12; i++) w[i] = src[i];
}
B381_DEV B381_INL void ext_st12(uint32_t* dst, const uint32_t (&w)[12]) {
#if defined(__CUDA_ARCH__)
  if ((reinterpret_cast<unsigned long long>(dst) & 15ull) == 0) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    d4[0] = make_uint4(w[0], w[1], w[2], w[3]); d4[1] = make_uint4(w[4], w[5], w[6], w[7]); d4[2] = make_uint4(w[8], w[9], w[10], w[11]);
    return;
  }
#endif
#pragma unroll
  for (int i = 0; i < 12; i++) dst[i] = w[i];
}
B381_NOINL bool f2_load_asis2(u4* r, const uint32_t* s0, const uint32_t* s1) {
  uint32_t w0[12], w1[12];
  ext_ld12(w0, s0); ext_ld12(w1, s1);
  Fp c0, c1;
  const bool ok0 = fp_from_ext_asis(c0, w0);
  const bool ok1 = fp_from_ext_asis(c1, w1);
  st_f2(r, c0, c1);
  return ok0 && ok1;
}
B381_NOINL void f2_store_unscale2(uint32_t* d0, uint32_t* d1, const u4* a) {
  Fp c0, c1;
  ld_f2(c0, c1, a);
  uint32_t w0[12], w1[12];
  fp_to_ext_unscale(w0, c0);
  fp_to_ext_unscale(w1, c1);
  ext_st12(d0, w0); ext_st12(d1, w1);
}

// ---------------------------------------------------------------------------------------------
// orchestration helpers (slot indices)
// ---------------------------------------------------------------------------------------------
#define S_(i) slot(cx, (i))

B381_DEV B381_INL void lin(const Ctx& cx, int r, int a, int b, int op) { sync_point_lin(cx); f2_lin(S_(r), S_(a), b >= 0 ? S_(b) : nullptr, op); }
B381_DEV B381_INL void mul(const Ctx& cx, int r, int a, int b) { sync_point(cx); f2_mul(S_(r), S_(a), S_(b)); }
B381_DEV B381_INL void mul_ss(const Ctx& cx, int r, int a, int a2, int b, int b2) {
  sync_point(cx);
  f2_mul_ss(S_(r), S_(a), a2 >= 0 ? S_(a2) : nullptr, S_(b), b2 >= 0 ? S_(b2) : nullptr);
}
B381_DEV B381_INL void sop3(const Ctx& cx, int r, int a0, int b0, int a1, int b1, int a2, int b2, int flags = 0, int p1 = -1, int p2 = -1, int r2 = -1) {
  sync_point(cx);
  if (flags & (SOP_SUB2 | SOP_HALFSUB | SOP_HALFSUB_XI | SOP_ALSO_ADD | SOP_ALSO_XIADD))
    f2_sop_post(S_(r), flags, S_(a0), S_(b0), S_(a1), S_(b1), S_(a2), S_(b2), p1 >= 0 ? S_(p1) : nullptr, p2 >= 0 ? S_(p2) : nullptr, r2 >= 0 ? S_(r2) : nullptr);
  else
    f2_sop(S_(r), flags, S_(a0), S_(b0), S_(a1), S_(b1), S_(a2), S_(b2));
}
B381_DEV B381_INL void mul_ex(const Ctx& cx, int r, int a, int b, int b2, int pre_b, int post, int p1 = -1, int p2 = -1) {
  sync_point(cx);
  f2_mul_ex(S_(r), S_(a), nullptr, PRE_NONE, S_(b), b2 >= 0 ? S_(b2) : nullptr, pre_b, post, p1 >= 0 ? S_(p1) : nullptr, p2 >= 0 ? S_(p2) : nullptr);
}
B381_DEV B381_INL void sqr_ex(const Ctx& cx, int r, int a, int a2, int pre, int post, int p1 = -1, int p2 = -1) {
  sync_point(cx);
  f2_sqr_ex(S_(r), S_(a), a2 >= 0 ? S_(a2) : nullptr, pre, post, p1 >= 0 ? S_(p1) : nullptr, p2 >= 0 ? S_(p2) : nullptr);
}
B381_DEV B381_INL void sqr(const Ctx& cx, int r, int a) { sync_point(cx); f2_sqr(S_(r), S_(a), nullptr); }
B381_DEV B381_INL void sqr_s(const Ctx& cx, int r, int a, int a2) { sync_point(cx); f2_sqr(S_(r), S_(a), S_(a2)); }
B381_DEV B381_INL void kcomb(const Ctx& cx, int r, int a, int b, int c, int d, int mode) {
  sync_point_lin(cx);
  f2_kcomb(S_(r), S_(a), b >= 0 ? S_(b) : nullptr, c >= 0 ? S_(c) : nullptr, d >= 0 ? S_(d) : nullptr, mode);
}

// Fp6 multiplication r = a * b; same value as the Karatsuba of fq6_target_tree.rs:172-214, arranged
// as three sums of three products (schoolbook over v with v^3 = xi, one reduction per coefficient):
//   c0 = a0 b0 + a1 (xi b2) + a2 (xi b1);  c1 = a0 b1 + a1 b0 + a2 (xi b2);  c2 = a0 b2 + a1 b1 + a2 b0
// 27 MAC blocks + 6 reductions instead of 18 + 12.  r must not alias a or b; t = 2 scratch slots.
B381_DEV B381_INL void f6_mul(const Ctx& cx, int r, int a, int b, int t) {
  const int xb1 = t, xb2 = t + 1;
  lin(cx, xb1, b + 1, -1, L_MULXI);
  lin(cx, xb2, b + 2, -1, L_MULXI);
  sop3(cx, r, a, b, a + 1, xb2, a + 2, xb1);
  sop3(cx, r + 1, a, b + 1, a + 1, b, a + 2, xb2);
  sop3(cx, r + 2, a, b + 2, a + 1, b + 1, a + 2, b);
}

// The same product with the two xi-multiplications folded into the operand loads of the sums (no scratch, no
// linear passes); b1, b2 must satisfy the SOP_XI bounds (single stored values, or weak-reduced sums).  dbl: 2 a b.
struct SopPost { int flags, p1, p2, r2; };
B381_DEV B381_INL void f6_mul_x(const Ctx& cx, int r, int a, int b, int dbl = 0, const SopPost* post = nullptr) {
  const int d = dbl ? SOP_DBL : 0;
  const SopPost none = {0, -1, -1, -1};
  const SopPost& q0 = post ? post[0] : none; const SopPost& q1 = post ? post[1] : none; const SopPost& q2 = post ? post[2] : none;
  sop3(cx, r, a, b, a + 1, b + 2, a + 2, b + 1, SOP_XI1 | SOP_XI2 | d | q0.flags, q0.p1, q0.p2, q0.r2);
  sop3(cx, r + 1, a, b + 1, a + 1, b, a + 2, b + 2, SOP_XI2 | d | q1.flags, q1.p1, q1.p2, q1.r2);
  sop3(cx, r + 2, a, b + 2, a + 1, b + 1, a + 2, b, d | q2.flags, q2.p1, q2.p2, q2.r2);
}

// Fp6 squaring via f6_mul-style Karatsuba with squarings (3 sqr + 3 mul)
B381_DEV B381_INL void f6_sqr(const Ctx& cx, int r, int a, int t) {
  const int v0 = t, v1 = t + 1, v2 = t + 2, m = t + 3;
  sqr(cx, v0, a);
  sqr(cx, v1, a + 1);
  sqr(cx, v2, a + 2);
  sqr_s(cx, m, a + 1, a + 2);
  kcomb(cx, r, m, v1, v2, v0, K_XI_INNER);
  sqr_s(cx, m, a, a + 1);
  kcomb(cx, r + 1, m, v0, v1, v2, K_XI_D);
  sqr_s(cx, m, a, a + 2);
  kcomb(cx, r + 2, m, v0, v2, v1, K_PLAIN);
}

// Fp12 multiplication r = a * b (3 Fp6 muls); fq12_target_tree.rs:130-141.
// r may alias a or b.  Scratch: t1 = 6 slots (aa, bb), t2 = 6 slots (sa, sb).  The xi-multiplications of the
// Fp6 products are folded into the sums of products (f6_mul_x): b and a + 3.. are stored values, the two
// coefficients of sb that get multiplied by xi are weak-reduced sums.
B381_DEV B381_INL void f12_mul(const Ctx& cx, int r, int a, int b, int t1, int t2) {
  const int aa = t1, bb = t1 + 3, sa = t2, sb = t2 + 3;
  for (int i = 0; i < 3; i++) {
    lin(cx, sa + i, a + i, a + 3 + i, L_ADD);
    lin(cx, sb + i, b + i, b + 3 + i, i == 0 ? L_ADD : L_ADD_R);
  }
  f6_mul_x(cx, aa, a, b);
  // bb = a1 b1, and with it c0 = aa + v bb = (aa0 + xi bb2, aa1 + bb0, aa2 + bb1) straight into r (a0, b0 are dead)
  const SopPost pb[3] = {{SOP_ALSO_ADD, aa + 1, -1, r + 1}, {SOP_ALSO_ADD, aa + 2, -1, r + 2}, {SOP_ALSO_XIADD, aa, -1, r}};
  f6_mul_x(cx, bb, a + 3, b + 3, 0, pb);
  // c1 = (a0 + a1)(b0 + b1) - aa - bb
  const SopPost pc[3] = {{SOP_SUB2, aa, bb, -1}, {SOP_SUB2, aa + 1, bb + 1, -1}, {SOP_SUB2, aa + 2, bb + 2, -1}};
  f6_mul_x(cx, r + 3, sa, sb, 0, pc);
}

// Fp12 complex squaring in place; fq12_target_tree.rs:143-155.  Scratch: t = 5 slots, s3 = 3 more
// slots (the Miller loops pass the line-coefficient slots, which are dead while f is squared).
B381_DEV void f12_sqr(const Ctx& cx, int f, int t, int s3) {
  const int ab = t, s = s3, w = t + 3;            // w: 2 slots
  f6_mul(cx, ab, f, f + 3, w);                    // ab = a0 a1
  for (int i = 0; i < 3; i++) lin(cx, s + i, f + i, f + 3 + i, L_ADD);      // s = a0 + a1
  // u = a0 + v a1 = (a00 + xi a12, a01 + a10, a02 + a11), built in place over a1
  lin(cx, w, f + 5, -1, L_COPY);                  // keep a12
  lin(cx, f + 5, f + 2, f + 4, L_ADD);            // u2
  lin(cx, f + 4, f + 1, f + 3, L_ADD);            // u1
  lin(cx, f + 3, f, w, L_XIADD);                  // u0
  f6_mul(cx, f, s, f + 3, w);                     // c0 = s * u  (a0 dead)
  kcomb(cx, f, f, ab, ab + 2, -1, K_XI_C);        // c0 -= ab + v ab ; v ab = (xi ab2, ab0, ab1)
  kcomb(cx, f + 1, f + 1, ab + 1, ab, -1, K_PLAIN);
  kcomb(cx, f + 2, f + 2, ab + 2, ab + 1, -1, K_PLAIN);
  for (int i = 0; i < 3; i++) lin(cx, f + 3 + i, ab + i, -1, L_DBL);        // c1 = 2 ab
}

// sparse multiplication f *= (c0 + c1 v + c4 v w), in place; same value as fq12_target_tree.rs:157-176
// and the native twin /root/reference/src/miller_loop_native.rs:118-137 (mul_by_01 / mul_by_1
// Karatsuba, 13 Fp2 muls), arranged as six sums of three products (18 products, 12 reductions
// instead of 26).  With f = (a0,a1,a2 | b0,b1,b2):
//   f0' = (a0 c0 + a2 xi c1 + b1 xi c4,  a0 c1 + a1 c0 + b2 xi c4,  a1 c1 + a2 c0 + b0 c4)
//   f1' = (b0 c0 + b2 xi c1 + a2 xi c4,  b0 c1 + b1 c0 + a0 c4,     b1 c1 + b2 c0 + a1 c4)
// t = 8 scratch slots (six outputs, xi c1, xi c4); the result is copied back over f.
B381_DEV void f12_mul_by_014(const Ctx& cx, int f, int c0, int c1, int c4, int t) {
  const int a0 = f, a1 = f + 1, a2 = f + 2, b0 = f + 3, b1 = f + 4, b2 = f + 5;
  const int xc1 = t + 6, xc4 = t + 7;
  lin(cx, xc1, c1, -1, L_MULXI);
  lin(cx, xc4, c4, -1, L_MULXI);
  sop3(cx, t + 0, a0, c0, a2, xc1, b1, xc4);
  sop3(cx, t + 1, a0, c1, a1, c0, b2, xc4);
  sop3(cx, t + 2, a1, c1, a2, c0, b0, c4);
  sop3(cx, t + 3, b0, c0, b2, xc1, a2, xc4);
  sop3(cx, t + 4, b0, c1, b1, c0, a0, c4);
  sop3(cx, t + 5, b1, c1, b2, c0, a1, c4);
  for (int i = 0; i < 6; i++) lin(cx, f + i, t + i, -1, L_COPY);
}

// Fp12 complex squaring OUT OF PLACE, d = f^2 (d, f distinct slot ranges), for the ARK Miller loops:
//   c1 = 2 a0 a1 straight from the sums (SOP_DBL) into d + 3..5;  c0 = (a0 + a1)(a0 + v a1) - (c1 + v c1) / 2.
// Six linear passes (s, u) and three recombinations instead of fourteen and three.  s3, t: 3 scratch slots each.
B381_DEV void f12_sqr_oop(const Ctx& cx, int d, int f, int t, int s3) {
  const int s = s3, u = t;
  for (int i = 0; i < 3; i++) lin(cx, s + i, f + i, f + 3 + i, L_ADD);      // s = a0 + a1
  lin(cx, u, f, f + 5, L_XIADD);                  // u = a0 + v a1 = (a00 + xi a12, a01 + a10, a02 + a11), weak-reduced
  lin(cx, u + 1, f + 1, f + 3, L_ADD_R);
  lin(cx, u + 2, f + 2, f + 4, L_ADD_R);
  f6_mul_x(cx, d + 3, f + 3, f, 1);               // c1 = 2 a1 a0
#ifdef B381_SQR_FOLD
  // c0 = s u - (c1 + v c1) / 2 ; v c1 = (xi c12, c10, c11): the recombination rides on the sums that produce s u
  const SopPost pc[3] = {{SOP_HALFSUB_XI, d + 3, d + 5, -1}, {SOP_HALFSUB, d + 4, d + 3, -1}, {SOP_HALFSUB, d + 5, d + 4, -1}};
  f6_mul_x(cx, d, s, u, 0, pc);
#else
  // (folding these three recombinations into the sums, as f12_mul does, was measured: -2 % on k_miller)
  f6_mul_x(cx, d, s, u);                          // s u
  kcomb(cx, d, d, d + 3, d + 5, -1, K_HALFSUB_XI);     // c0 = s u - (c1 + v c1) / 2 ; v c1 = (xi c12, c10, c11)
  kcomb(cx, d + 1, d + 1, d + 4, d + 3, -1, K_HALFSUB);
  kcomb(cx, d + 2, d + 2, d + 5, d + 4, -1, K_HALFSUB);
#endif
}

// sparse multiplication d = f * (c0 + c1 v + c4 v w) OUT OF PLACE, xi c1 and xi c4 formed in registers:
// six sums of three products and nothing else.  c1, c4 are Fp-scalar products (below 1.03 p).
B381_DEV void f12_mul_by_014_oop(const Ctx& cx, int d, int f, int c0, int c1, int c4) {
  const int a0 = f, a1 = f + 1, a2 = f + 2, b0 = f + 3, b1 = f + 4, b2 = f + 5;
  sop3(cx, d + 0, a0, c0, a2, c1, b1, c4, SOP_XI1 | SOP_XI2);
  sop3(cx, d + 1, a0, c1, a1, c0, b2, c4, SOP_XI2);
  sop3(cx, d + 2, a1, c1, a2, c0, b0, c4);
  sop3(cx, d + 3, b0, c0, b2, c1, a2, c4, SOP_XI1 | SOP_XI2);
  sop3(cx, d + 4, b0, c1, b1, c0, a0, c4);
  sop3(cx, d + 5, b1, c1, b2, c0, a1, c4);
}

B381_DEV B381_INL void f12_set_one(const Ctx& cx, int f) {
  f2_set_small(S_(f), 1);
  for (int i = 1; i < 6; i++) f2_set_small(S_(f + i), 0);
}

B381_DEV B381_INL void f12_copy(const Ctx& cx, int r, int a) {
  for (int i = 0; i < 6; i++) lin(cx, r + i, a + i, -1, L_COPY);
}

// conjugation (c0, -c1) in place; fq12_target_tree.rs:53-58; /root/reference/src/miller_loop_native.rs:194-200
B381_DEV B381_INL void f12_conj(const Ctx& cx, int f) {
  for (int i = 3; i < 6; i++) lin(cx, f + i, f + i, -1, L_NEG);
}

// Frobenius map f -> f^(p^k) in place, k in 1..3; fq12_target_tree.rs:92-128, fq6_target_tree.rs:129-169.
// Slot f + 3 i + j holds the coefficient of w^(2j+i).
B381_DEV B381_INL void f12_frobenius(const Ctx& cx, int f, int k) {
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 3; j++) {
      const int tw = 2 * j + i;
      u4* p = S_(f + 3 * i + j);
      sync_point(cx);
      if (tw == 0) { if (k & 1) f2_lin(p, p, nullptr, L_CONJ); }
      else f2_mul_gamma(p, p, k, tw, k & 1);
    }
}

// Fp6 inverse; fq6_target_tree.rs:59-89.  r may alias a.  t = 5 scratch slots.
B381_DEV B381_INL void f6_inv(const Ctx& cx, int r, int a, int t) {
  const int c0 = t, c1 = t + 1, c2 = t + 2, x = t + 3, y = t + 4;
  sqr(cx, c0, a); mul(cx, x, a + 1, a + 2); kcomb(cx, c0, c0, -1, x, -1, K_XI_C);        // c0 = a0^2 - xi a1 a2
  sqr(cx, c1, a + 2); lin(cx, c1, c1, -1, L_MULXI); mul(cx, x, a, a + 1); lin(cx, c1, c1, x, L_SUB);   // c1 = xi a2^2 - a0 a1
  sqr(cx, c2, a + 1); mul(cx, x, a, a + 2); lin(cx, c2, c2, x, L_SUB);                   // c2 = a1^2 - a0 a2
  mul(cx, x, a + 2, c1); mul(cx, y, a + 1, c2); lin(cx, x, x, y, L_ADD); lin(cx, x, x, -1, L_MULXI);
  mul(cx, y, a, c0); lin(cx, x, x, y, L_ADD);                                            // t = xi(a2 c1 + a1 c2) + a0 c0
  sync_point(cx);
  f2_inv(S_(x), S_(x));
  mul(cx, r, x, c0); mul(cx, r + 1, x, c1); mul(cx, r + 2, x, c2);
}

// Fp12 inverse in place; fq12_target_tree.rs:77-90.  t = 15 scratch slots.
B381_DEV B381_INL void f12_inv(const Ctx& cx, int f, int t) {
  const int u = t, v = t + 3, w = t + 6;          // w: up to 5 + 4
  f6_sqr(cx, u, f, w);
  f6_sqr(cx, v, f + 3, w);
  kcomb(cx, u, u, -1, v + 2, -1, K_XI_C);         // u = c0^2 - v c1^2 ; v x = (xi x2, x0, x1)
  lin(cx, u + 1, u + 1, v, L_SUB);
  lin(cx, u + 2, u + 2, v + 1, L_SUB);
  f6_inv(cx, u, u, w);
  f6_mul(cx, v, f, u, w);
  for (int i = 0; i < 3; i++) lin(cx, f + i, v + i, -1, L_COPY);
  f6_mul(cx, v, f + 3, u, w);
  for (int i = 0; i < 3; i++) lin(cx, f + 3 + i, v + i, -1, L_NEG);
}

// Granger-Scott cyclotomic squaring d = s^2, OUT OF PLACE (d and s distinct Fp12 slot ranges);
// /root/reference/src/fields_as_trees/miller_loop.rs:46-104.  With z0=c0.c0, z4=c0.c1, z3=c0.c2,
// z2=c1.c0, z1=c1.c1, z5=c1.c2:  (z0',z1') from fp4(z0,z1); (z4',z5') from fp4(z2,z3);
// (z2',z3') from fp4(z4,z5) with the xi twist.  Three fused primitive calls, no scratch.
B381_DEV B381_INL void f12_cyclotomic_square(const Ctx& cx, int d, int s, int reduce = 1) {
  const int z0 = 0, z4 = 1, z3 = 2, z2 = 3, z1 = 4, z5 = 5;
  sync_point(cx);
  f2_cyc_fp4(S_(d + z0), S_(d + z1), S_(s + z0), S_(s + z1), S_(s + z0), S_(s + z1), 0, reduce);
  sync_point(cx);
  f2_cyc_fp4(S_(d + z4), S_(d + z5), S_(s + z2), S_(s + z3), S_(s + z4), S_(s + z5), 0, reduce);
  sync_point(cx);
  f2_cyc_fp4(S_(d + z2), S_(d + z3), S_(s + z4), S_(s + z5), S_(s + z2), S_(s + z3), 1, reduce);
}

// r = conj(a^|x|): ark Bls12::exp_by_x (x < 0), plain left-to-right form over Granger-Scott squarings.  The running
// value ping-pongs between the two fixed slot ranges acc / acc2; r may alias a.  t = 12 scratch slots.  Not used by
// the kernels any more: the tests keep it as the cross-check of the compressed form below.
B381_DEV B381_INL void f12_exp_by_x_gs(const Ctx& cx, int r, int a, int acc, int acc2, int t) {
  int cur = acc, nxt = acc2;
  const uint64_t xabs = B381_X_ABS;
  for (int b = 62; b >= 0; b--) {
    f12_cyclotomic_square(cx, nxt, b == 62 ? a : cur);
    const int sw = cur; cur = nxt; nxt = sw;
    if ((xabs >> b) & 1) f12_mul(cx, cur, cur, a, t, t + 6);
  }
  for (int i = 0; i < 3; i++) lin(cx, r + i, cur + i, -1, L_COPY);
  for (int i = 3; i < 6; i++) lin(cx, r + i, cur + i, -1, L_NEG);
}

// Karabina's compressed squaring (Squaring in cyclotomic subgroups, Math. Comp. 2013): the Granger-Scott formulas
// for (z2, z3, z4, z5) do not involve (z0, z1), so a run of squarings carries four coefficients and costs two fused
// Fp4 squarings instead of three.  d, s: Fp12 slot ranges of which only +1, +2, +3, +5 are touched.
B381_DEV B381_INL void f12_cyclotomic_square_c(const Ctx& cx, int d, int s) {
  const int z4 = 1, z3 = 2, z2 = 3, z5 = 5;
  sync_point(cx);
  f2_cyc_fp4(S_(d + z4), S_(d + z5), S_(s + z2), S_(s + z3), S_(s + z4), S_(s + z5), 0, 1);
  sync_point(cx);
  f2_cyc_fp4(S_(d + z2), S_(d + z3), S_(s + z4), S_(s + z5), S_(s + z2), S_(s + z3), 1, 1);
}

// n compressed squarings src -> ... -> last (n >= 1); intermediate values alternate p0, p1, p0, ...; the caller
// picks them so that the one before the last is not `last` itself (the squaring is out of place)
B381_DEV B381_INL void f12_csqr_run(const Ctx& cx, int last, int src, int n, int p0, int p1) {
  int cur = src, nxt = p0, oth = p1;
  for (int i = 1; i <= n; i++) {
    const int dst = i == n ? last : nxt;
    f12_cyclotomic_square_c(cx, dst, cur);
    cur = dst;
    const int sw = nxt; nxt = oth; oth = sw;
  }
}

// r = conj(a^|x|), |x| = 2^63 + 2^62 + 2^60 + 2^57 + 2^48 + 2^16, right to left:  a^|x| is the product of a^(2^i) over
// those six i.  The first 57 squarings run in compressed form; the three values needed in full (i = 16, 48, 57)
//   z1 = (xi z5^2 + 3 z4^2 - 2 z3) / (4 z2),   z0 = (2 z1^2 + z2 z5 - 3 z3 z4) xi + 1
// share ONE Fp2 inversion (Montgomery's trick; the inversion itself is the division-step one, fp32.cuh); the last six
// squarings are ordinary Granger-Scott ones.  132 + 18 fused Fp4 squarings instead of 189 per exponentiation, paid
// with 6 sums of products, 9 Fp2 products and one inversion: the value is the same field element, so the output of
// the final exponentiation is unchanged bit for bit.
// z2 = 0: the relation a b = s c^2 + conj(b) between the Fp4 coefficients of a cyclotomic element (the t-coefficient
// of the Granger-Scott identity) gives z1 = 2 z4 z5 / z3 there, and z2 = z3 = 0 only for the identity (z1 = 0, z0 = 1;
// gcd(p^4 - 1, p^4 - p^2 + 1) = 1).  Those cases are patched by per-thread selections inside a WARP-uniform branch
// that runs without the lock-step barriers, so the barrier sequence of the CTA does not depend on the data.
// r must not alias a; acc, acc2: 6 slots each (hot); t: 12 scratch slots.
B381_DEV B381_INL void f12_exp_by_x(const Ctx& cx, int r, int a, int acc, int acc2, int t) {
  const int z0 = 0, z4 = 1, z3 = 2, z2 = 3, z1 = 4, z5 = 5;
  f12_csqr_run(cx, r, a, 16, acc, acc2);                 // compressed a^(2^16) -> r
  f12_csqr_run(cx, t, r, 32, acc, acc2);                 // compressed a^(2^48) -> t .. t + 5
  f12_csqr_run(cx, acc, t, 9, acc, acc2);                // compressed a^(2^57) -> acc (the 8th lands in acc2)
  const int B[3] = {r, t, acc};
  const int u = t + 6, w = acc2;                         // 6 + 6 free slots
  const int ZERO = u, ONE = u + 1, X = u + 2, Y = u + 3;
  sync_point(cx); f2_set_small(S_(ZERO), 0);
  sync_point(cx); f2_set_small(S_(ONE), 1);
  bool zf[3], any = false;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int b = B[k];
    sync_point(cx);
    zf[k] = f2_is_zero(S_(b + z2));
    any = any || zf[k];
    lin(cx, X, b + z4, -1, L_TRIPLE);
    sop3(cx, b + z1, b + z5, b + z5, b + z4, X, ZERO, ZERO, SOP_XI0 | SOP_SUB2, b + z3, b + z3);   // numerator
    lin(cx, b + z0, b + z2, -1, L_MUL4);                                                            // denominator
  }
#if defined(__CUDA_ARCH__)
  any = __any_sync(__activemask(), any);
#endif
  if (any) {                                             // z2 = 0 somewhere in this warp (the identity, mostly)
    Ctx c2 = cx;
    c2.sync = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int b = B[k];
      mul(c2, X, b + z4, b + z5);
      lin(c2, X, X, X, L_ADD_R);                         // 2 z4 z5
      f2_select_if(slot(c2, b + z1), slot(c2, X), zf[k]);
      f2_select_if(slot(c2, b + z0), slot(c2, b + z3), zf[k]);
      const bool dz = f2_is_zero(slot(c2, b + z0));
      f2_override_if(slot(c2, b + z0), dz, 1);           // the identity: numerator 0, denominator 1
    }
  }
  mul(cx, w, B[0] + z0, B[1] + z0);                      // d0 d1
  mul(cx, w + 1, w, B[2] + z0);                          // d0 d1 d2
  sync_point(cx); f2_inv(S_(w + 1), S_(w + 1));
  mul(cx, w + 2, w + 1, w);                              // 1 / d2
  mul(cx, w + 1, w + 1, B[2] + z0);                      // 1 / (d0 d1)
  mul(cx, w + 3, w + 1, B[0] + z0);                      // 1 / d1
  mul(cx, w, w + 1, B[1] + z0);                          // 1 / d0
  const int inv[3] = {w, w + 3, w + 2};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int b = B[k];
    mul(cx, b + z1, b + z1, inv[k]);
    lin(cx, X, b + z4, -1, L_TRIPLE);
    lin(cx, X, X, -1, L_NEG);                            // -3 z4
    lin(cx, Y, b + z1, -1, L_DBL);                       // 2 z1
    sop3(cx, b + z0, Y, b + z1, b + z2, b + z5, b + z3, X, SOP_XI0 | SOP_XI1 | SOP_XI2);   // xi goes on the undoubled factor (operand bounds)
    lin(cx, b + z0, b + z0, ONE, L_ADD_R);
  }
  f12_mul(cx, r, r, t, w, u);                            // a^(2^16 + 2^48)
  f12_mul(cx, r, r, acc, t, u);                          // ... + 2^57
  f12_cyclotomic_square(cx, w, acc);
  f12_cyclotomic_square(cx, acc, w);
  f12_cyclotomic_square(cx, w, acc);                     // a^(2^60)
  f12_mul(cx, r, r, w, t, u);
  f12_cyclotomic_square(cx, acc, w);
  f12_cyclotomic_square(cx, w, acc);                     // a^(2^62)
  f12_mul(cx, r, r, w, t, u);
  f12_cyclotomic_square(cx, acc, w);                     // a^(2^63)
  f12_mul(cx, r, r, acc, t, u);
  for (int i = 3; i < 6; i++) lin(cx, r + i, r + i, -1, L_NEG);
}

// ---------------------------------------------------------------------------------------------
// ARK mode: G2 homogeneous-projective steps (ark-ec 0.4 models/bls12/g2.rs)
// R = slots (X, Y, Z) = (R, R+1, R+2); line coefficients -> (L, L+1, L+2); t = 5 scratch slots.
// ---------------------------------------------------------------------------------------------
B381_DEV void ark_double_step(const Ctx& cx, int R, int L, int t) {
  const int X = R, Y = R + 1, Z = R + 2, T0 = t, T1 = t + 1, T2 = t + 2, T3 = t + 3, T4 = t + 4;
  mul(cx, T0, X, Y); lin(cx, T0, T0, -1, L_HALF);                  // a = X Y / 2
  sqr(cx, T1, Y);                                                  // b = Y^2
  sqr(cx, T2, Z);                                                  // c = Z^2
  sqr(cx, L + 1, X); lin(cx, L + 1, L + 1, -1, L_TRIPLE);          // 3 j = 3 X^2
  sqr_s(cx, T3, Y, Z); kcomb(cx, T3, T3, T1, T2, -1, K_PLAIN);     // h = (Y+Z)^2 - (b + c)
  lin(cx, T2, T2, -1, L_MUL12XI);                                  // e = B' 3c = 12 xi c
  lin(cx, L, T2, T1, L_SUB);                                       // i = e - b
  lin(cx, T4, T2, -1, L_TRIPLE);                                   // f = 3e
  lin(cx, X, T1, T4, L_SUB); mul(cx, X, T0, X);                    // X' = a (b - f)
  lin(cx, T0, T1, T4, L_HALFSUM);                                  // g = (b + f) / 2
  mul(cx, Z, T1, T3);                                              // Z' = b h
  lin(cx, L + 2, T3, -1, L_NEG);                                   // -h
  sqr(cx, T2, T2); lin(cx, T2, T2, -1, L_TRIPLE);                  // 3 e^2
  sqr(cx, Y, T0); lin(cx, Y, Y, T2, L_SUB);                        // Y' = g^2 - 3 e^2
}

// mixed addition R += Q, Q = slots (Qs, Qs+1) affine; coefficients (j, -theta, lambda)
B381_DEV void ark_add_step(const Ctx& cx, int R, int Qs, int L, int t) {
  const int X = R, Y = R + 1, Z = R + 2, qx = Qs, qy = Qs + 1;
  const int th = t, la = t + 1, c = t + 2, d = t + 3, e = t + 4;
  mul(cx, th, qy, Z); lin(cx, th, Y, th, L_SUB);                   // theta = Y - qy Z
  mul(cx, la, qx, Z); lin(cx, la, X, la, L_SUB);                   // lambda = X - qx Z
  sqr(cx, c, th);                                                  // c = theta^2
  sqr(cx, d, la);                                                  // d = lambda^2
  mul(cx, e, la, d);                                               // e = lambda d
  mul(cx, c, Z, c);                                                // f = Z c
  mul(cx, d, X, d);                                                // g = X d
  lin(cx, c, e, c, L_ADD); lin(cx, X, d, -1, L_DBL); lin(cx, c, c, X, L_SUB);   // h = e + f - 2g
  mul(cx, X, la, c);                                               // X' = lambda h
  lin(cx, d, d, c, L_SUB); mul(cx, d, th, d);                      // theta (g - h)
  mul(cx, Y, e, Y); lin(cx, Y, d, Y, L_SUB);                       // Y' = theta (g-h) - e Y
  mul(cx, Z, Z, e);                                                // Z' = Z e
  mul(cx, c, th, qx); mul(cx, d, la, qy); lin(cx, L, c, d, L_SUB); // j = theta qx - lambda qy
  lin(cx, L + 1, th, -1, L_NEG);
  lin(cx, L + 2, la, -1, L_COPY);
}

// ell for the M-twist: c2 *= py ; c1 *= px ; f.mul_by_014(c0, c1, c2).  Pt = slot (px, py).
B381_DEV B381_INL void ark_ell(const Ctx& cx, int f, int L, int Pt, int t) {
  sync_point(cx);
  f2_mulfp(S_(L + 2), S_(L + 2), S_(Pt), 1);
  f2_mulfp(S_(L + 1), S_(L + 1), S_(Pt), 0);
  f12_mul_by_014(cx, f, L, L + 1, L + 2, t);
}

// The doubling step with its linear operations folded into the multiplications (same values as
// ark_double_step; coefficients (i, 3j, h) -- the sign of h is applied by ark_ell_oop's scalar product).
// 9 multiplication primitives + 1 linear pass instead of 9 + 11.  t = 5 scratch slots (t + 3 unused).
B381_DEV void ark_double_step_fused(const Ctx& cx, int R, int L, int t) {
  const int X = R, Y = R + 1, Z = R + 2, T0 = t, T1 = t + 1, T2 = t + 2, T4 = t + 4;
  mul_ex(cx, T0, X, Y, -1, PRE_NONE, POST_HALF);                   // a = X Y / 2
  sqr(cx, T1, Y);                                                  // b = Y^2
  sqr(cx, T2, Z);                                                  // c = Z^2
  sqr_ex(cx, L + 1, X, -1, PRE_NONE, POST_TRIPLE);                 // 3 j = 3 X^2
  sqr_ex(cx, L + 2, Y, Z, PRE_ADD_RAW, POST_SUB2, T1, T2);         // h = (Y+Z)^2 - (b + c)
  sync_point_lin(cx);
  f2_dbl_lin3(S_(T2), S_(L), S_(T4), S_(T2), S_(T1));              // e = 12 xi c ; i = e - b ; f = 3 e
  mul_ex(cx, X, T0, T1, T4, PRE_SUB3P, POST_NONE);                 // X' = a (b - f)
  mul(cx, Z, T1, L + 2);                                           // Z' = b h
  sqr_ex(cx, T2, T2, -1, PRE_NONE, POST_TRIPLE);                   // 3 e^2
  sqr_ex(cx, Y, T1, T4, PRE_HALFSUM, POST_SUB1, T2);               // Y' = ((b + f) / 2)^2 - 3 e^2
}

// ell, out of place: d = f * line;  c2 *= py (negated first when the step left h instead of -h), c1 *= px
B381_DEV B381_INL void ark_ell_oop(const Ctx& cx, int d, int f, int L, int Pt, int neg2) {
  sync_point(cx);
  f2_mulfp(S_(L + 2), S_(L + 2), S_(Pt), 1, neg2);
  f2_mulfp(S_(L + 1), S_(L + 1), S_(Pt), 0);
  f12_mul_by_014_oop(cx, d, f, L, L + 1, L + 2);
}

// The same against a line held OUTSIDE the arena (packed G2Prepared, programs.cuh): `line` points at three consecutive
// slot-shaped values (i, 3j, h).  c1 / c2 are scaled by px / py on their way into the arena slots L + 1, L + 2; c0 is
// read in place by the six sums of products (one cold read, then L1 / L2 hits).
B381_DEV B381_INL void ark_ell_oop_ext(const Ctx& cx, int d, int f, const u4* line, int L, int Pt, int neg2) {
  sync_point(cx);
  f2_mulfp(S_(L + 2), line + 2 * SLOT, S_(Pt), 1, neg2);
  f2_mulfp(S_(L + 1), line + SLOT, S_(Pt), 0);
  const int a0 = f, a1 = f + 1, a2 = f + 2, b0 = f + 3, b1 = f + 4, b2 = f + 5, c1 = L + 1, c4 = L + 2;
  sync_point(cx); f2_sop(S_(d + 0), SOP_XI1 | SOP_XI2, S_(a0), line, S_(a2), S_(c1), S_(b1), S_(c4));
  sync_point(cx); f2_sop(S_(d + 1), SOP_XI2, S_(a0), S_(c1), S_(a1), line, S_(b2), S_(c4));
  sync_point(cx); f2_sop(S_(d + 2), 0, S_(a1), S_(c1), S_(a2), line, S_(b0), S_(c4));
  sync_point(cx); f2_sop(S_(d + 3), SOP_XI1 | SOP_XI2, S_(b0), line, S_(b2), S_(c1), S_(a2), S_(c4));
  sync_point(cx); f2_sop(S_(d + 4), 0, S_(b0), S_(c1), S_(b1), line, S_(a0), S_(c4));
  sync_point(cx); f2_sop(S_(d + 5), 0, S_(b1), S_(c1), S_(b2), line, S_(a1), S_(c4));
}

// slot plan shared by the Miller-loop kernels
struct MillerSlots { int f, L, T, R, Q, P; };
// Ping-pong plan of the ARK loops: f alternates between the banks A and B (every squaring and every line
// multiplication is out of place); up to KPP pairs per thread share the squarings.  Result in A.
constexpr int KPP = 4;
struct MillerSlotsPP { int A, B, L, T; int R[KPP], Q[KPP], P[KPP]; };

B381_DEV B381_INL void line_to_one_if(const Ctx& cx, int L, bool ident, int one_at) {
  for (int i = 0; i < 3; i++) f2_override_if(S_(L + i), ident, i == one_at);   // uniform: the line may sit in tensor memory
}

// Bls12::multi_miller_loop for k pairs per thread with SHARED squarings (ark squares f once per
// bit for a whole chunk of pairs): f = 1; per bit: f = f^2; ell(double) per pair; if bit:
// ell(add) per pair; finally conjugate (x < 0).  Pair j: Q (affine) at slots (Q[j], Q[j]+1), P at P[j], running
// point at R[j]..+2.  A pair flagged in `ident` (may be null) multiplies by the line (1, 0, 0), i.e. by one, so
// control flow stays uniform.  T = 5 scratch slots.
// One bit of the loop with f in bank `cur` on entry.  All bank swaps are resolved at compile time (the function
// is inlined with constant cur / oth: runtime bank indices cost the single-pair loop 4 %): a bit does 1 + K out-of-
// place operations, K more on an addition bit; for odd K the addition lines are followed by a copy back, so the
// bank after the bit does not depend on the bit's value.  Returns nothing: the caller knows the parity.
template <int K>
B381_DEV B381_INL void ark_bit_pp(const Ctx& cx, const MillerSlotsPP& s, int cur, int oth, bool square, bool add, const bool* ident) {
  if (square) {
    f12_sqr_oop(cx, oth, cur, s.T, s.L);
    const int sw = cur; cur = oth; oth = sw;
  }
#pragma unroll
  for (int j = 0; j < K; j++) {
    ark_double_step_fused(cx, s.R[j], s.L, s.T);
    if (ident) line_to_one_if(cx, s.L, ident[j], 0);
    ark_ell_oop(cx, oth, cur, s.L, s.P[j], 1);
    const int sw = cur; cur = oth; oth = sw;
  }
  if (add) {
#pragma unroll
    for (int j = 0; j < K; j++) {
      ark_add_step(cx, s.R[j], s.Q[j], s.L, s.T);
      if (ident) line_to_one_if(cx, s.L, ident[j], 0);
      ark_ell_oop(cx, oth, cur, s.L, s.P[j], 0);
      const int sw = cur; cur = oth; oth = sw;
    }
    if (K & 1) f12_copy(cx, oth, cur);              // odd K: back to the bank a bit without addition ends in
  }
}

template <int K>
B381_DEV void ark_miller_loop_pp(const Ctx& cx, const MillerSlotsPP& s, const bool* ident) {
  const uint64_t xabs = B381_X_ABS;
  // the leading one of |x| (bit 63) is implicit; bit 62 is handled like any other, only the squaring of f = 1 is skipped
  if (K & 1) {
    // odd K: the first bit (no squaring) flips the bank, every later bit keeps it: start in B, live in A
    f12_set_one(cx, s.B);
    for (int j = 0; j < K; j++) { lin(cx, s.R[j], s.Q[j], -1, L_COPY); lin(cx, s.R[j] + 1, s.Q[j] + 1, -1, L_COPY); f2_set_small(S_(s.R[j] + 2), 1); }
    ark_bit_pp<K>(cx, s, s.B, s.A, false, (xabs >> 62) & 1, ident);
    for (int b = 61; b >= 0; b--) ark_bit_pp<K>(cx, s, s.A, s.B, true, (xabs >> b) & 1, ident);
  } else {
    // even K: the first bit keeps the bank, every later bit flips it: 62 later bits, two per iteration
    f12_set_one(cx, s.A);
    for (int j = 0; j < K; j++) { lin(cx, s.R[j], s.Q[j], -1, L_COPY); lin(cx, s.R[j] + 1, s.Q[j] + 1, -1, L_COPY); f2_set_small(S_(s.R[j] + 2), 1); }
    ark_bit_pp<K>(cx, s, s.A, s.B, false, (xabs >> 62) & 1, ident);
    for (int b = 61; b >= 0; b -= 2) {
      ark_bit_pp<K>(cx, s, s.A, s.B, true, (xabs >> b) & 1, ident);
      ark_bit_pp<K>(cx, s, s.B, s.A, true, (xabs >> (b - 1)) & 1, ident);
    }
  }
  f12_conj(cx, s.A);
}

// slot plan of the ZK-mode two-pair loop (in-place primitives)
struct MultiSlots { int f, L, T; int R[KPP], Q[KPP], P[KPP]; };

// ---------------------------------------------------------------------------------------------
// ZK mode: /root/reference/src/miller_loop_native.rs:27-116 with ell (:139-152) wired in.
// ---------------------------------------------------------------------------------------------
// Alg. 26 doubling, miller_loop_native.rs:27-55.  t = 7 scratch slots.  coeffs -> (L, L+1, L+2)
B381_DEV void zk_double_step(const Ctx& cx, int R, int L, int t) {
  const int x = R, y = R + 1, z = R + 2;
  const int t0 = t, t1 = t + 1, t2 = t + 2, t3 = t + 3, t4 = t + 4, t5 = t + 5, zs = t + 6;
  sqr(cx, t0, x);                                                  // tmp0 = x^2
  sqr(cx, t1, y);                                                  // tmp1 = y^2
  sqr(cx, t2, t1);                                                 // tmp2 = tmp1^2
  sqr_s(cx, t3, t1, x); kcomb(cx, t3, t3, t0, t2, -1, K_PLAIN); lin(cx, t3, t3, -1, L_DBL);   // tmp3
  lin(cx, t4, t0, -1, L_TRIPLE);                                   // tmp4 = 3 tmp0
  lin(cx, L + 2, x, t4, L_ADD);                                    // tmp6 = x + tmp4
  sqr(cx, t5, t4);                                                 // tmp5 = tmp4^2
  sqr(cx, zs, z);                                                  // zsquared
  kcomb(cx, x, t5, t3, t3, -1, K_PLAIN);                           // x' = tmp5 - 2 tmp3
  sqr_s(cx, z, z, y); kcomb(cx, z, z, t1, zs, -1, K_PLAIN);        // z' = (z+y)^2 - tmp1 - zsq
  lin(cx, y, t3, x, L_SUB); mul(cx, y, y, t4);                     // y' = (tmp3 - x') tmp4
  lin(cx, t2, t2, -1, L_MUL8); lin(cx, y, y, t2, L_SUB);           // y' -= 8 tmp2
  mul(cx, L + 1, t4, zs); lin(cx, L + 1, L + 1, -1, L_DBL); lin(cx, L + 1, L + 1, -1, L_NEG);   // -2 tmp4 zsq
  sqr(cx, L + 2, L + 2); kcomb(cx, L + 2, L + 2, t0, t5, -1, K_PLAIN);                          // tmp6^2 - tmp0 - tmp5
  lin(cx, t1, t1, -1, L_MUL4); lin(cx, L + 2, L + 2, t1, L_SUB);   // - 4 tmp1
  mul(cx, L, z, zs); lin(cx, L, L, -1, L_DBL);                     // 2 z' zsq
}

// Alg. 27 mixed addition, miller_loop_native.rs:58-87.  t = 10 scratch slots.
B381_DEV void zk_add_step(const Ctx& cx, int R, int Qs, int L, int t) {
  const int x = R, y = R + 1, z = R + 2, qx = Qs, qy = Qs + 1;
  const int zs = t, ys = t + 1, t0 = t + 2, t1 = t + 3, t2 = t + 4, t3 = t + 5, t4 = t + 6, t5 = t + 7, t6 = t + 8, t7 = t + 9;
  sqr(cx, zs, z);
  sqr(cx, ys, qy);
  mul(cx, t0, zs, qx);
  sqr_s(cx, t1, qy, z); kcomb(cx, t1, t1, ys, zs, -1, K_PLAIN); mul(cx, t1, t1, zs);
  lin(cx, t2, t0, x, L_SUB);
  sqr(cx, t3, t2);
  lin(cx, t4, t3, -1, L_MUL4);
  mul(cx, t5, t4, t2);
  kcomb(cx, t6, t1, y, y, -1, K_PLAIN);                            // t6 = t1 - 2y
  mul(cx, L + 2, t6, qx);                                          // t9
  mul(cx, t7, t4, x);
  sqr(cx, x, t6); kcomb(cx, x, x, t5, t7, -1, K_PLAIN); lin(cx, x, x, t7, L_SUB);   // x' = t6^2 - t5 - 2 t7
  sqr_s(cx, z, z, t2); kcomb(cx, z, z, zs, t3, -1, K_PLAIN);       // z' = (z + t2)^2 - zs - t3
  lin(cx, t7, t7, x, L_SUB); mul(cx, t7, t7, t6);                  // t8 = (t7 - x') t6
  mul(cx, t0, y, t5); lin(cx, t0, t0, -1, L_DBL);                  // t0 = 2 y t5
  lin(cx, y, t7, t0, L_SUB);                                       // y' = t8 - t0
  sqr_s(cx, t0, qy, z); lin(cx, t0, t0, ys, L_SUB);                // t10 = (qy + z')^2 - ysq
  sqr(cx, t1, z); lin(cx, t0, t0, t1, L_SUB);                      // t10 -= z'^2
  lin(cx, L + 2, L + 2, t0, L_2A_MB);                              // t9 = 2 t9 - t10
  lin(cx, L, z, -1, L_DBL);                                        // t10 = 2 z'
  lin(cx, L + 1, t6, -1, L_DBL); lin(cx, L + 1, L + 1, -1, L_NEG); // t1 = -2 t6
}

// miller_loop_native.rs:139-152: c0 *= py; c1 *= px; f.mul_by_014(coeffs.2, c1, c0)
B381_DEV B381_INL void zk_ell(const Ctx& cx, int f, int L, int Pt, int t) {
  sync_point(cx);
  f2_mulfp(S_(L), S_(L), S_(Pt), 1);
  f2_mulfp(S_(L + 1), S_(L + 1), S_(Pt), 0);
  f12_mul_by_014(cx, f, L + 2, L + 1, L, t);
}

// driver miller_loop_native.rs:89-116
B381_DEV void zk_miller_loop(const Ctx& cx, const MillerSlots& s) {
  f12_set_one(cx, s.f);
  lin(cx, s.R, s.Q, -1, L_COPY);
  lin(cx, s.R + 1, s.Q + 1, -1, L_COPY);
  f2_set_small(S_(s.R + 2), 1);
  const uint64_t xh = B381_X_ABS >> 1;
  bool found_one = false;
  for (int b = 63; b >= 0; b--) {
    bool bit = (xh >> b) & 1;
    if (!found_one) { found_one = bit; continue; }
    zk_double_step(cx, s.R, s.L, s.T);
    zk_ell(cx, s.f, s.L, s.P, s.T);
    if (bit) {
      zk_add_step(cx, s.R, s.Q, s.L, s.T);
      zk_ell(cx, s.f, s.L, s.P, s.T);
    }
    f12_sqr(cx, s.f, s.T, s.L);
  }
  zk_double_step(cx, s.R, s.L, s.T);
  zk_ell(cx, s.f, s.L, s.P, s.T);
  f12_conj(cx, s.f);
}

// ZK-mode multi-Miller loop with shared squarings (the Adder driver of
// /root/reference/src/miller_loop_native.rs:154-212 loops over all terms inside each step).
B381_DEV void zk_miller_loop_multi(const Ctx& cx, const MultiSlots& s, int k, const bool* ident) {
  f12_set_one(cx, s.f);
  for (int j = 0; j < k; j++) {
    lin(cx, s.R[j], s.Q[j], -1, L_COPY);
    lin(cx, s.R[j] + 1, s.Q[j] + 1, -1, L_COPY);
    f2_set_small(S_(s.R[j] + 2), 1);
  }
  const uint64_t xh = B381_X_ABS >> 1;
  bool found_one = false;
  for (int b = 63; b >= -1; b--) {
    bool bit = b >= 0 ? ((xh >> b) & 1) : false;
    if (b >= 0 && !found_one) { found_one = bit; continue; }
    for (int j = 0; j < k; j++) {
      zk_double_step(cx, s.R[j], s.L, s.T);
      line_to_one_if(cx, s.L, ident[j], 2);          // zk_ell uses (L+2, L+1, L) as (c0, c1, c4)
      zk_ell(cx, s.f, s.L, s.P[j], s.T);
    }
    if (b < 0) break;                              // the trailing doubling step of the driver (:109)
    if (bit) {
      for (int j = 0; j < k; j++) {
        zk_add_step(cx, s.R[j], s.Q[j], s.L, s.T);
        line_to_one_if(cx, s.L, ident[j], 2);
        zk_ell(cx, s.f, s.L, s.P[j], s.T);
      }
    }
    f12_sqr(cx, s.f, s.T, s.L);
  }
  f12_conj(cx, s.f);
}

// ---------------------------------------------------------------------------------------------
// final exponentiation f -> f^(3 (p^12 - 1)/r): ark Bls12::final_exponentiation chain
// (same exponent as the zkcrypto chain at
// /root/reference/src/fields_as_trees/miller_loop.rs:128-178).  In place on f.
// y0, y1, y2, r2 = four Fp12 scratch values (6 slots each); t = 16 scratch slots.
// ---------------------------------------------------------------------------------------------
struct FexpSlots { int f, y0, y1, y2, r, acc, acc2, T; };

B381_DEV B381_INL void final_exponentiation(const Ctx& cx, const FexpSlots& s) {
  const int f = s.f, y0 = s.y0, y1 = s.y1, y2 = s.y2, r = s.r, acc = s.acc, acc2 = s.acc2, T = s.T;
  // easy part: r = f^((p^6-1)(p^2+1))
  f12_copy(cx, r, f);
  f12_conj(cx, r);                                // f1 = conj(f)
  f12_inv(cx, f, T);                              // f2 = f^-1
  f12_mul(cx, r, r, f, T, T + 6);                        // r = f1 f2
  f12_copy(cx, f, r);                             // f2 = r
  f12_frobenius(cx, r, 2);
  f12_mul(cx, r, r, f, T, T + 6);
  // hard part
  f12_cyclotomic_square(cx, y0, r);                                 // y0 = r^2
  f12_exp_by_x(cx, y1, r, acc, acc2, T);                                       // y1 = r^x
  f12_copy(cx, y2, r); f12_conj(cx, y2);                            // y2 = r^-1
  f12_mul(cx, y1, y1, y2, T, T + 6);
  f12_exp_by_x(cx, y2, y1, acc, acc2, T);
  f12_conj(cx, y1);
  f12_mul(cx, y1, y1, y2, T, T + 6);
  f12_exp_by_x(cx, y2, y1, acc, acc2, T);
  f12_frobenius(cx, y1, 1);
  f12_mul(cx, y1, y1, y2, T, T + 6);
  f12_mul(cx, r, r, y0, T, T + 6);
  f12_exp_by_x(cx, y0, y1, acc, acc2, T);
  f12_exp_by_x(cx, y2, y0, acc, acc2, T);
  f12_copy(cx, y0, y1); f12_frobenius(cx, y0, 2);
  f12_conj(cx, y1);
  f12_mul(cx, y1, y1, y2, T, T + 6);
  f12_mul(cx, y1, y1, y0, T, T + 6);
  f12_mul(cx, f, r, y1, T, T + 6);
}

// ---------------------------------------------------------------------------------------------
// LITERAL mode: /root/reference/src/miller_loop_native_optimized.rs:8-127 exactly as written
// (everything in Fq2; Jacobian coordinates fed to a homogeneous line formula; one squaring as
// "final exponentiation").  Slots: R (3), Q (3), P (x,y | z,-) in 2 slots as Fq2-embedded Fp,
// fn, fd, n, d, and t = 8 scratch slots.  Returns false where the reference panics (f_den = 0).
// ---------------------------------------------------------------------------------------------
struct LiteralSlots { int R, Q, Pp, fn, fd, n, d, T; };   // Pp: 3 slots xp, yp, zp as Fq2 (c1 = 0)

// ark-ec 0.4 Projective::double_in_place (a = 0, dbl-2009-l); t = 5 scratch
B381_DEV void jac_double(const Ctx& cx, int R, int t) {
  const int X = R, Y = R + 1, Z = R + 2, A = t, B = t + 1, C = t + 2, D = t + 3, E = t + 4;
  if (f2_is_zero(S_(Z))) return;
  sqr(cx, A, X);
  sqr(cx, B, Y);
  sqr(cx, C, B);
  sqr_s(cx, D, X, B); kcomb(cx, D, D, A, C, -1, K_PLAIN); lin(cx, D, D, -1, L_DBL);
  lin(cx, E, A, -1, L_TRIPLE);
  mul(cx, Z, Y, Z); lin(cx, Z, Z, -1, L_DBL);                      // Z3 = 2 Y Z
  sqr(cx, X, E); kcomb(cx, X, X, D, D, -1, K_PLAIN);               // X3 = E^2 - 2D
  lin(cx, D, D, X, L_SUB); mul(cx, Y, E, D);                       // E (D - X3)
  lin(cx, C, C, -1, L_MUL8); lin(cx, Y, Y, C, L_SUB);              // - 8C
}

// ark-ec 0.4 Projective += Projective (add-2007-bl); result in R1.  t = 9 scratch
B381_DEV void jac_add(const Ctx& cx, int R1, int R2, int t) {
  const int X1 = R1, Y1 = R1 + 1, Z1 = R1 + 2, X2 = R2, Y2 = R2 + 1, Z2 = R2 + 2;
  const int z1z1 = t, z2z2 = t + 1, u1 = t + 2, u2 = t + 3, s1 = t + 4, s2 = t + 5, h = t + 6, i = t + 7, j = t + 8;
  if (f2_is_zero(S_(Z1))) { for (int k = 0; k < 3; k++) lin(cx, R1 + k, R2 + k, -1, L_COPY); return; }
  if (f2_is_zero(S_(Z2))) return;
  sqr(cx, z1z1, Z1);
  sqr(cx, z2z2, Z2);
  mul(cx, u1, X1, z2z2);
  mul(cx, u2, X2, z1z1);
  mul(cx, s1, Y1, Z2); mul(cx, s1, s1, z2z2);
  mul(cx, s2, Y2, Z1); mul(cx, s2, s2, z1z1);
  if (f2_equal(S_(u1), S_(u2)) && f2_equal(S_(s1), S_(s2))) { jac_double(cx, R1, t); return; }
  lin(cx, h, u2, u1, L_SUB);
  lin(cx, i, h, -1, L_DBL); sqr(cx, i, i);
  mul(cx, j, h, i);
  lin(cx, s2, s2, s1, L_SUB); lin(cx, s2, s2, -1, L_DBL);          // r = 2 (S2 - S1)
  mul(cx, u1, u1, i);                                              // V = U1 I
  sqr_s(cx, Z1, Z1, Z2); kcomb(cx, Z1, Z1, z1z1, z2z2, -1, K_PLAIN); mul(cx, Z1, Z1, h);   // Z3
  sqr(cx, X1, s2); kcomb(cx, X1, X1, j, u1, -1, K_PLAIN); lin(cx, X1, X1, u1, L_SUB);      // X3 = r^2 - J - 2V
  lin(cx, u1, u1, X1, L_SUB); mul(cx, u1, s2, u1);                 // r (V - X3)
  mul(cx, s1, s1, j); lin(cx, s1, s1, -1, L_DBL);                  // 2 S1 J
  lin(cx, Y1, u1, s1, L_SUB);
}

// R1 += Q with Q affine in slots (Q, Q + 1) (madd-2007-bl: 7 products + 4 squarings instead of 11 + 5); the same
// group element as jac_add on (Q, 1), a different Jacobian representative.  t = 9 scratch.  Bucket sums only: results
// leave through a canonical affine conversion.
B381_DEV void jac_add_mixed(const Ctx& cx, int R1, int Q, int t) {
  const int X1 = R1, Y1 = R1 + 1, Z1 = R1 + 2, X2 = Q, Y2 = Q + 1;
  const int z1z1 = t, u2 = t + 1, s2 = t + 2, h = t + 3, hh = t + 4, i = t + 5, j = t + 6, v = t + 7;
  if (f2_is_zero(S_(Z1))) {
    lin(cx, X1, X2, -1, L_COPY); lin(cx, Y1, Y2, -1, L_COPY); f2_set_small(S_(Z1), 1);
    return;
  }
  sqr(cx, z1z1, Z1);
  mul(cx, u2, X2, z1z1);
  mul(cx, s2, Y2, Z1); mul(cx, s2, s2, z1z1);
  if (f2_equal(S_(u2), S_(X1)) && f2_equal(S_(s2), S_(Y1))) { jac_double(cx, R1, t); return; }
  lin(cx, h, u2, X1, L_SUB);
  sqr(cx, hh, h);
  lin(cx, i, hh, -1, L_MUL4);                                      // I = 4 H^2
  mul(cx, j, h, i);
  lin(cx, s2, s2, Y1, L_SUB); lin(cx, s2, s2, -1, L_DBL);          // r = 2 (S2 - Y1)
  mul(cx, v, X1, i);                                               // V = X1 I
  sqr_s(cx, Z1, Z1, h); kcomb(cx, Z1, Z1, z1z1, hh, -1, K_PLAIN);  // Z3 = (Z1 + H)^2 - Z1Z1 - HH
  sqr(cx, X1, s2); kcomb(cx, X1, X1, j, v, -1, K_PLAIN); lin(cx, X1, X1, v, L_SUB);   // X3 = r^2 - J - 2 V
  lin(cx, v, v, X1, L_SUB); mul(cx, v, s2, v);                     // r (V - X3)
  mul(cx, j, Y1, j); lin(cx, j, j, -1, L_DBL);                     // 2 Y1 J
  lin(cx, Y1, v, j, L_SUB);
}

// optimized_line_function (:8-78): (n, d) for the line through Q1, Q2 at P.  t = 6 scratch
B381_DEV void literal_line(const Ctx& cx, int Q1, int Q2, int Pp, int n, int d, int t) {
  const int x1 = Q1, y1 = Q1 + 1, z1 = Q1 + 2, x2 = Q2, y2 = Q2 + 1, z2 = Q2 + 2, xp = Pp, yp = Pp + 1, zp = Pp + 2;
  const int num = t, den = t + 1, A = t + 2, B = t + 3, w = t + 4, w2 = t + 5;
  mul(cx, A, xp, z1); mul(cx, w, x1, zp); lin(cx, A, A, w, L_SUB);
  mul(cx, B, yp, z1); mul(cx, w, y1, zp); lin(cx, B, B, w, L_SUB);
  bool den_zero = true, num_zero = true;           // Q1 == Q2 (same object): both differences are 0
  if (Q1 != Q2) {
    mul(cx, num, y2, z1); mul(cx, w, y1, z2); lin(cx, num, num, w, L_SUB);
    mul(cx, den, x2, z1); mul(cx, w, x1, z2); lin(cx, den, den, w, L_SUB);
    den_zero = f2_is_zero(S_(den));
    num_zero = den_zero && f2_is_zero(S_(num));
  }
  if (den_zero && !num_zero) {                     // vertical line
    lin(cx, n, A, -1, L_COPY);
    mul(cx, d, z1, zp);
    return;
  }
  if (den_zero) {                                  // tangent
    sqr(cx, num, x1); lin(cx, num, num, -1, L_TRIPLE);             // 3 x1^2  (= Fq2::from(3) * x1 * x1)
    mul(cx, den, y1, z1); lin(cx, den, den, -1, L_DBL);            // 2 y1 z1
  }
  mul(cx, w, num, A); mul(cx, w2, den, B); lin(cx, n, w, w2, L_SUB);
  mul(cx, d, den, zp); mul(cx, d, d, z1);
}

B381_DEV bool literal_optimized_miller_loop(const Ctx& cx, const LiteralSlots& s, int out /* slot for c0.c0 */) {
  for (int k = 0; k < 3; k++) lin(cx, s.R + k, s.Q + k, -1, L_COPY);
  f2_set_small(S_(s.fn), 1);
  f2_set_small(S_(s.fd), 1);
  const uint64_t xabs = B381_X_ABS;
  for (int v = 0; v < 64; v++) {                   // PSEUDO_BINARY_ENCODING, forward (LSB first)
    literal_line(cx, s.R, s.R, s.Pp, s.n, s.d, s.T);
    sqr(cx, s.fn, s.fn); mul(cx, s.fn, s.fn, s.n);
    sqr(cx, s.fd, s.fd); mul(cx, s.fd, s.fd, s.d);
    jac_double(cx, s.R, s.T);                      // R + R: ark's add detects U1==U2 && S1==S2 -> double
    if ((xabs >> v) & 1) {
      literal_line(cx, s.R, s.Q, s.Pp, s.n, s.d, s.T);
      mul(cx, s.fn, s.fn, s.n);
      mul(cx, s.fd, s.fd, s.d);
      jac_add(cx, s.R, s.Q, s.T);
    }
  }
  if (f2_is_zero(S_(s.fd))) return false;
  f2_inv(S_(s.fd), S_(s.fd));
  mul(cx, out, s.fn, s.fd);
  sqr(cx, out, out);
  return true;
}

#undef S_

}  // namespace b381
