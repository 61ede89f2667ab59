// programs.cuh -- per-thread programs (one unit of work per thread) shared by the CUDA kernels
// (kernels.cu) and the host simulation used by the CPU-side tests (tests/hostsim/hostsim.cpp).
// External buffers use the C-ABI layout of include/b381.h.
#pragma once
#include "tower.cuh"

namespace b381 {

enum Mode { MODE_ARK = 0, MODE_ZK = 1, MODE_LITERAL = 2 };

// error bits accumulated per call (include/b381.h)
enum ErrBits { ERR_NOT_CANONICAL = 1, ERR_ZERO_DIVISION = 2 };

// ---- slot plans -------------------------------------------------------------------------------
// Miller loop: f and the line coefficients and the first scratch slots are the hot set (shared
// memory when NS = 16); R, Q, P are touched only by the curve steps.
constexpr int ML_F = 0, ML_L = 6, ML_T = 9, ML_R = 19, ML_Q = 22, ML_P = 24, ML_ACC = 25, ML_NSLOTS = 31;
// ARK Miller loop, ping-pong plan (tower.cuh ark_miller_loop_pp): banks A (= ML_F) and B, line, 5 scratch slots
constexpr int PP_A = 0, PP_B = 6, PP_L = 12, PP_T = 15, PP_R = 20, PP_Q = 23, PP_P = 25;
// ... second pair of the shared-squaring multi-pair loop and the running product of k_multi_miller (pairs 3, 4: M4_*)
constexpr int PP_R2 = 26, PP_Q2 = 29, PP_P2 = 31, PP_ACC = 32, PP_NSLOTS = 38;
// final exponentiation: ACC (the value being squared) and the first scratch slots are hot.
// T = 12 scratch slots for the Fp12 products; the inversion of the easy part needs 15 and runs over into Y1, which is
// not live yet.  48 slots: 29 of them in global memory = 105 MB of scratch for 148 CTAs, inside the 126 MB L2.
constexpr int FE_ACC = 0, FE_ACC2 = 6, FE_T = 12, FE_Y1 = 24, FE_Y2 = 30, FE_F = 36, FE_Y0 = 36 /* f is dead once y0 is first written */, FE_R = 42, FE_NSLOTS = 48;
// literal loop
constexpr int LT_R = 0, LT_Q = 3, LT_P = 6, LT_FN = 9, LT_FD = 10, LT_N = 11, LT_D = 12, LT_T = 13, LT_OUT = 22, LT_NSLOTS = 23;
// shared-squaring multi-Miller, ZK plan: second pair's R, Q, P and the running product
constexpr int M2_R2 = 25, M2_Q2 = 28, M2_P2 = 30, M2_ACC = 32 /* = PP_ACC: one accumulator slot range for both modes */, M2_NSLOTS = 38;
constexpr int M2_SCRATCH = 6;   // 14 scratch slots of the accumulating Fp12 product: dead line / scratch / bank-B slots of either plan
// pairs three and four of the shared-squaring multi-Miller loop (either plan): R (3 slots), Q (2), P (1) each
constexpr int MK = 4;                 // pairs per thread of the on-the-fly multi-Miller loop
constexpr int M4_R3 = 38, M4_Q3 = 41, M4_P3 = 43, M4_R4 = 44, M4_Q4 = 47, M4_P4 = 49, M4_NSLOTS = 50;
constexpr int MAX_NSLOTS = 50;
// Stores under a per-thread `if (ident)` (miller_to_slots & co.) must never hit tensor memory (tcgen05.st is
// .sync.aligned): the P / Q input slots live in the global-memory part of the arena, f in shared memory.
static_assert(ML_Q >= NS + NT_MAX && ML_P >= NS + NT_MAX && M2_Q2 >= NS + NT_MAX && M2_P2 >= NS + NT_MAX, "P / Q slots must be global-memory slots");
static_assert(ML_F + 5 < NS, "f must be a shared-memory slot range");
static_assert(PP_Q >= NS + NT_MAX && PP_P >= NS + NT_MAX && PP_A == ML_F, "ping-pong plan: P / Q in global memory, result where ML_F is");
static_assert(PP_NSLOTS <= MAX_NSLOTS && M2_NSLOTS <= MAX_NSLOTS && M4_NSLOTS <= MAX_NSLOTS && PP_ACC == M2_ACC && M4_R3 >= PP_ACC + 6, "arena size / shared accumulator");

#define S_(i) slot(cx, (i))

B381_DEV B381_INL bool f12_load_ext(const Ctx& cx, int f, const uint32_t* src) {
  bool ok = true;
  for (int i = 0; i < 6; i++) ok &= f2_load_ext(S_(f + i), src + 24 * i);
  return ok;
}

B381_DEV B381_INL void f12_store_ext(const Ctx& cx, uint32_t* dst, int f) {
  for (int i = 0; i < 6; i++) f2_store_ext(dst + 24 * i, S_(f + i));
}

// raw (internal-format) dump / load of an Fp12: 6 slots x 28 words, used for partial products
B381_DEV B381_INL void f12_store_raw(const Ctx& cx, uint32_t* dst, int f) {
  for (int i = 0; i < 6; i++) {
    Fp c0, c1;
    ld_f2(c0, c1, S_(f + i));
    for (int k = 0; k < NL; k++) { dst[28 * i + k] = (uint32_t)c0.l[k]; dst[28 * i + NL + k] = (uint32_t)c1.l[k]; }
  }
}

B381_DEV B381_INL void f12_load_raw(const Ctx& cx, int f, const uint32_t* src) {
  for (int i = 0; i < 6; i++) {
    Fp c0, c1;
    for (int k = 0; k < NL; k++) { c0.l[k] = (limb_t)src[28 * i + k]; c1.l[k] = (limb_t)src[28 * i + NL + k]; }
    B381_SETRANGE(c0, 0.0, 5.0); B381_SETRANGE(c1, 0.0, 5.0);     // raw values are stored values
    st_f2(S_(f + i), c0, c1);
  }
}

// Miller loop of one pair into slots ML_F.. (internal format).  Returns error bits.
// g1: 24 words (x, y), g2: 48 words (x.c0, x.c1, y.c0, y.c1); inf bit0 = P is identity, bit1 = Q.
B381_DEV B381_INL int miller_to_slots(const Ctx& cx, const uint32_t* g1, const uint32_t* g2, int inf, int mode) {
  int err = 0;
  const bool ident = (inf & 3) != 0;               // identity pairs contribute 1 (ark drops them)
  const int sP = mode == MODE_ZK ? ML_P : PP_P, sQ = mode == MODE_ZK ? ML_Q : PP_Q;
  if (ident) {                                     // run the (uniform) loop on zeros, discard the result
    f2_set_small(S_(sP), 0); f2_set_small(S_(sQ), 0); f2_set_small(S_(sQ + 1), 0);
  } else {
    if (!f2_load_ext(S_(sP), g1)) err |= ERR_NOT_CANONICAL;
    if (!f2_load_ext(S_(sQ), g2)) err |= ERR_NOT_CANONICAL;
    if (!f2_load_ext(S_(sQ + 1), g2 + 24)) err |= ERR_NOT_CANONICAL;
  }
  if (mode == MODE_ZK) {
    MillerSlots s;
    s.f = ML_F; s.L = ML_L; s.T = ML_T; s.R = ML_R; s.Q = ML_Q; s.P = ML_P;
    zk_miller_loop(cx, s);
  } else {
    MillerSlotsPP s;
    s.A = PP_A; s.B = PP_B; s.L = PP_L; s.T = PP_T; s.R[0] = PP_R; s.Q[0] = PP_Q; s.P[0] = PP_P;
    ark_miller_loop_pp<1>(cx, s, nullptr);
  }
  if (ident) f12_set_one(cx, ML_F);
  return err;
}

// Miller loops of MK = FOUR pairs with shared squarings into slots ML_F.. ; identity pairs contribute 1.
B381_DEV B381_INL int miller_multi_to_slots(const Ctx& cx, const uint32_t* const* g1, const uint32_t* const* g2, const int* inf, int mode) {
  int err = 0;
  bool ident[MK];
  const bool zk = mode == MODE_ZK;
  const int Rj[MK] = {zk ? ML_R : PP_R, zk ? M2_R2 : PP_R2, M4_R3, M4_R4};
  const int Qj[MK] = {zk ? ML_Q : PP_Q, zk ? M2_Q2 : PP_Q2, M4_Q3, M4_Q4};
  const int Pj[MK] = {zk ? ML_P : PP_P, zk ? M2_P2 : PP_P2, M4_P3, M4_P4};
  for (int j = 0; j < MK; j++) {
    ident[j] = (inf[j] & 3) != 0;
    if (ident[j]) {
      f2_set_small(S_(Pj[j]), 0); f2_set_small(S_(Qj[j]), 0); f2_set_small(S_(Qj[j] + 1), 0);
    } else {
      if (!f2_load_ext(S_(Pj[j]), g1[j])) err |= ERR_NOT_CANONICAL;
      if (!f2_load_ext(S_(Qj[j]), g2[j])) err |= ERR_NOT_CANONICAL;
      if (!f2_load_ext(S_(Qj[j] + 1), g2[j] + 24)) err |= ERR_NOT_CANONICAL;
    }
  }
  if (zk) {
    MultiSlots s;
    s.f = ML_F; s.L = ML_L; s.T = ML_T;
    for (int j = 0; j < MK; j++) { s.R[j] = Rj[j]; s.Q[j] = Qj[j]; s.P[j] = Pj[j]; }
    zk_miller_loop_multi(cx, s, MK, ident);
  } else {
    MillerSlotsPP s;
    s.A = PP_A; s.B = PP_B; s.L = PP_L; s.T = PP_T;
    for (int j = 0; j < MK; j++) { s.R[j] = Rj[j]; s.Q[j] = Qj[j]; s.P[j] = Pj[j]; }
    ark_miller_loop_pp<MK>(cx, s, ident);
  }
  return err;
}

// final exponentiation of the value in slots `src`..src+5, result left in FE_F.
B381_DEV B381_INL int final_exp_slots(const Ctx& cx, int src) {
  if (src != FE_F) f12_copy(cx, FE_F, src);
  bool zero = true;
  for (int i = 0; i < 6; i++) zero = f2_is_zero(S_(FE_F + i)) && zero;
  FexpSlots s;
  s.f = FE_F; s.y0 = FE_Y0; s.y1 = FE_Y1; s.y2 = FE_Y2; s.r = FE_R; s.acc = FE_ACC; s.acc2 = FE_ACC2; s.T = FE_T;
  final_exponentiation(cx, s);                     // uniform control flow: f = 0 just flows through as 0
  return zero ? ERR_ZERO_DIVISION : 0;             // ark final_exponentiation(0) is None
}

// ---- programs -----------------------------------------------------------------------------------
B381_DEV B381_INL int prog_miller(const Ctx& cx, const uint32_t* g1, const uint32_t* g2, int inf, uint32_t* out, int mode) {
  int err = miller_to_slots(cx, g1, g2, inf, mode);
  f12_store_ext(cx, out, ML_F);
  return err;
}

// ---- G2Prepared: the line coefficients of Q as a cached stage --------------------
// ark-ec G2Prepared { ell_coeffs: Vec<(Fp2, Fp2, Fp2)>, infinity } (the shape the reference's circuit
// side mirrors at /root/reference/src/miller_loop_target.rs:23-76): 68 triples per Q in the order the
// Miller loop consumes them (63 doublings, 5 additions), each triple 3 x Fq2 in the external format
// = 72 words; B381_G2PREP_WORDS = 68 * 72 = 4896 words per point.
constexpr int G2PREP_TRIPLES = 68;
constexpr int G2PREP_WORDS = G2PREP_TRIPLES * 72;

B381_DEV B381_INL void store_triple(const Ctx& cx, uint32_t* dst, int L) {
  for (int k = 0; k < 3; k++) f2_store_ext(dst + 24 * k, S_(L + k));
}
B381_DEV B381_INL bool load_triple(const Ctx& cx, int L, const uint32_t* src) {
  bool ok = true;
  for (int k = 0; k < 3; k++) ok &= f2_load_ext(S_(L + k), src + 24 * k);
  return ok;
}

// coefficients of one Q (affine, external format) -> coeffs[0 .. 4896)
B381_DEV B381_INL int prog_g2_prepare(const Ctx& cx, const uint32_t* g2, uint32_t* coeffs, int mode) {
  int err = 0;
  if (!f2_load_ext(S_(ML_Q), g2)) err |= ERR_NOT_CANONICAL;
  if (!f2_load_ext(S_(ML_Q + 1), g2 + 24)) err |= ERR_NOT_CANONICAL;
  lin(cx, ML_R, ML_Q, -1, L_COPY);
  lin(cx, ML_R + 1, ML_Q + 1, -1, L_COPY);
  f2_set_small(S_(ML_R + 2), 1);
  int idx = 0;
  if (mode == MODE_ZK) {                           // schedule of zk_miller_loop
    const uint64_t xh = B381_X_ABS >> 1;
    bool found_one = false;
    for (int b = 63; b >= 0; b--) {
      bool bit = (xh >> b) & 1;
      if (!found_one) { found_one = bit; continue; }
      zk_double_step(cx, ML_R, ML_L, ML_T); store_triple(cx, coeffs + 72 * idx++, ML_L);
      if (bit) { zk_add_step(cx, ML_R, ML_Q, ML_L, ML_T); store_triple(cx, coeffs + 72 * idx++, ML_L); }
    }
    zk_double_step(cx, ML_R, ML_L, ML_T); store_triple(cx, coeffs + 72 * idx++, ML_L);
  } else {                                         // schedule of ark_miller_loop
    const uint64_t xabs = B381_X_ABS;
    for (int b = 62; b >= 0; b--) {
      ark_double_step(cx, ML_R, ML_L, ML_T); store_triple(cx, coeffs + 72 * idx++, ML_L);
      if ((xabs >> b) & 1) { ark_add_step(cx, ML_R, ML_Q, ML_L, ML_T); store_triple(cx, coeffs + 72 * idx++, ML_L); }
    }
  }
  return err;
}

// Miller loop of (P, prepared Q) -> slots ML_F..; same values as miller_to_slots on (P, Q)
B381_DEV B381_INL int miller_prepared_to_slots(const Ctx& cx, const uint32_t* g1, const uint32_t* coeffs, int inf, int mode) {
  int err = 0;
  const bool ident = (inf & 3) != 0;
  if (ident) f2_set_small(S_(ML_P), 0);
  else if (!f2_load_ext(S_(ML_P), g1)) err |= ERR_NOT_CANONICAL;
  f12_set_one(cx, ML_F);
  int idx = 0;
  bool ok = true;
  if (mode == MODE_ZK) {
    const uint64_t xh = B381_X_ABS >> 1;
    bool found_one = false;
    for (int b = 63; b >= 0; b--) {
      bool bit = (xh >> b) & 1;
      if (!found_one) { found_one = bit; continue; }
      ok &= load_triple(cx, ML_L, coeffs + 72 * idx++); zk_ell(cx, ML_F, ML_L, ML_P, ML_T);
      if (bit) { ok &= load_triple(cx, ML_L, coeffs + 72 * idx++); zk_ell(cx, ML_F, ML_L, ML_P, ML_T); }
      f12_sqr(cx, ML_F, ML_T, ML_L);
    }
    ok &= load_triple(cx, ML_L, coeffs + 72 * idx++); zk_ell(cx, ML_F, ML_L, ML_P, ML_T);
  } else {
    const uint64_t xabs = B381_X_ABS;
    // ping-pong banks with static roles: squaring A -> B, line B -> A; an addition line goes A -> B and is copied back
    for (int b = 62; b >= 0; b--) {
      if (b != 62) f12_sqr_oop(cx, PP_B, PP_A, PP_T, PP_L);
      else f12_set_one(cx, PP_B);
      ok &= load_triple(cx, PP_L, coeffs + 72 * idx++);
      ark_ell_oop(cx, PP_A, PP_B, PP_L, ML_P, 0);
      if ((xabs >> b) & 1) {
        ok &= load_triple(cx, PP_L, coeffs + 72 * idx++);
        ark_ell_oop(cx, PP_B, PP_A, PP_L, ML_P, 0);
        f12_copy(cx, PP_A, PP_B);
      }
    }
  }
  f12_conj(cx, ML_F);
  if (ident) f12_set_one(cx, ML_F);
  else if (!ok) err |= ERR_NOT_CANONICAL;
  return err;
}

B381_DEV B381_INL int prog_miller_prepared(const Ctx& cx, const uint32_t* g1, const uint32_t* coeffs, int inf, uint32_t* out, int mode, int do_fe) {
  int err = miller_prepared_to_slots(cx, g1, coeffs, inf, mode);
  if (do_fe) { err |= final_exp_slots(cx, ML_F); f12_store_ext(cx, out, FE_F); }
  else f12_store_ext(cx, out, ML_F);
  return err;
}

// ---- G2Prepared in the library's own layout ("packed": device memory only) ----------------------------
// The same 68 triples per Q, but (1) in the INTERNAL number format -- the values the on-the-fly loop would hand from
// its curve step to its line multiplication, so consuming a coefficient costs no conversion (the external format
// costs six Fp multiplications per line) -- and (2) laid out exactly like a range of arena slots: points are grouped
// in tiles of B381_GS (= the CTA size on the device, 1 on the host simulation); inside a tile, uint4 group g of
// coefficient c of point j sits at (c * GPS + g) * B381_GS + j.  A warp therefore reads 512 contiguous bytes per
// group, and a coefficient is addressed like a slot: the Miller loop multiplies straight out of the buffer.
// ARK mode stores what ark_double_step_fused leaves, (i, 3j, h): the sign of h is applied by the consumer.
constexpr int G2PACK_SLOTS = G2PREP_TRIPLES * 3;
B381_DEV B381_INL size_t g2pack_u4_per_tile() { return (size_t)G2PACK_SLOTS * SLOT; }
B381_DEV B381_INL size_t g2pack_index(size_t i) { return (i / B381_GS) * g2pack_u4_per_tile() + (i % B381_GS); }

B381_DEV B381_INL void pack_triple(const Ctx& cx, u4* dst, int L) {
  for (int k = 0; k < 3; k++) { sync_point_lin(cx); f2_lin(dst + k * SLOT, S_(L + k), nullptr, L_COPY); }
}
B381_DEV B381_INL void prefetch_triple(const u4* src) {
#if defined(__CUDA_ARCH__)
  for (int g = 0; g < 3 * GPS; g++) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + g * B381_GS));
#else
  (void)src;
#endif
}

// pk = this point's first group (buffer + g2pack_index(i))
B381_DEV B381_INL int prog_g2_prepare_packed(const Ctx& cx, const uint32_t* g2, u4* pk, int mode) {
  int err = 0;
  const int sQ = mode == MODE_ZK ? ML_Q : PP_Q, sR = mode == MODE_ZK ? ML_R : PP_R;
  if (!f2_load_ext(S_(sQ), g2)) err |= ERR_NOT_CANONICAL;
  if (!f2_load_ext(S_(sQ + 1), g2 + 24)) err |= ERR_NOT_CANONICAL;
  lin(cx, sR, sQ, -1, L_COPY);
  lin(cx, sR + 1, sQ + 1, -1, L_COPY);
  f2_set_small(S_(sR + 2), 1);
  int idx = 0;
  if (mode == MODE_ZK) {
    const uint64_t xh = B381_X_ABS >> 1;
    bool found_one = false;
    for (int b = 63; b >= 0; b--) {
      bool bit = (xh >> b) & 1;
      if (!found_one) { found_one = bit; continue; }
      zk_double_step(cx, ML_R, ML_L, ML_T); pack_triple(cx, pk + (size_t)3 * SLOT * idx++, ML_L);
      if (bit) { zk_add_step(cx, ML_R, ML_Q, ML_L, ML_T); pack_triple(cx, pk + (size_t)3 * SLOT * idx++, ML_L); }
    }
    zk_double_step(cx, ML_R, ML_L, ML_T); pack_triple(cx, pk + (size_t)3 * SLOT * idx++, ML_L);
  } else {
    const uint64_t xabs = B381_X_ABS;
    for (int b = 62; b >= 0; b--) {
      ark_double_step_fused(cx, PP_R, PP_L, PP_T); pack_triple(cx, pk + (size_t)3 * SLOT * idx++, PP_L);
      if ((xabs >> b) & 1) { ark_add_step(cx, PP_R, PP_Q, PP_L, PP_T); pack_triple(cx, pk + (size_t)3 * SLOT * idx++, PP_L); }
    }
  }
  return err;
}

// Miller loop of (P, packed Q) -> slots ML_F..; same values as miller_to_slots on (P, Q)
B381_DEV B381_INL int miller_packed_to_slots(const Ctx& cx, const uint32_t* g1, const u4* pk, int inf, int mode) {
  int err = 0;
  const bool ident = (inf & 3) != 0;
  const int sP = mode == MODE_ZK ? ML_P : PP_P;
  if (ident) f2_set_small(S_(sP), 0);
  else if (!f2_load_ext(S_(sP), g1)) err |= ERR_NOT_CANONICAL;
  int idx = 0;
  prefetch_triple(pk);
  if (mode == MODE_ZK) {
    f12_set_one(cx, ML_F);
    const uint64_t xh = B381_X_ABS >> 1;
    bool found_one = false;
    auto line = [&]() {
      const u4* src = pk + (size_t)3 * SLOT * idx++;
      if (idx < G2PREP_TRIPLES) prefetch_triple(src + 3 * SLOT);
      for (int k = 0; k < 3; k++) { sync_point_lin(cx); f2_lin(S_(ML_L + k), src + k * SLOT, nullptr, L_COPY); }
      zk_ell(cx, ML_F, ML_L, ML_P, ML_T);
    };
    for (int b = 63; b >= 0; b--) {
      bool bit = (xh >> b) & 1;
      if (!found_one) { found_one = bit; continue; }
      line();
      if (bit) line();
      f12_sqr(cx, ML_F, ML_T, ML_L);
    }
    line();
  } else {
    const uint64_t xabs = B381_X_ABS;
    // static bank roles as in miller_prepared_to_slots: squaring A -> B, doubling line B -> A, addition line A -> B + copy back
    for (int b = 62; b >= 0; b--) {
      if (b != 62) f12_sqr_oop(cx, PP_B, PP_A, PP_T, PP_L);
      else f12_set_one(cx, PP_B);
      const u4* src = pk + (size_t)3 * SLOT * idx++;
      if (idx < G2PREP_TRIPLES) prefetch_triple(src + 3 * SLOT);
      ark_ell_oop_ext(cx, PP_A, PP_B, src, PP_L, PP_P, 1);
      if ((xabs >> b) & 1) {
        src = pk + (size_t)3 * SLOT * idx++;
        if (idx < G2PREP_TRIPLES) prefetch_triple(src + 3 * SLOT);
        ark_ell_oop_ext(cx, PP_B, PP_A, src, PP_L, PP_P, 0);
        f12_copy(cx, PP_A, PP_B);
      }
    }
  }
  f12_conj(cx, ML_F);
  if (ident) f12_set_one(cx, ML_F);
  return err;
}

B381_DEV B381_INL int prog_miller_packed(const Ctx& cx, const uint32_t* g1, const u4* pk, int inf, uint32_t* out, int mode, int do_fe) {
  int err = miller_packed_to_slots(cx, g1, pk, inf, mode);
  if (do_fe) { err |= final_exp_slots(cx, ML_F); f12_store_ext(cx, out, FE_F); }
  else f12_store_ext(cx, out, ML_F);
  return err;
}

// ---- multi-Miller loop over packed prepared Q's: K pairs per thread share every squaring ----------------------
// (ark squares f once per bit for the whole batch; with the curve steps cached the per-thread state of a
// pair is just P, so four pairs fit where the on-the-fly loop holds two).  ARK mode.  Pair j reads its lines at
// line[j] + 3 * SLOT * (triple index); a pair that contributes 1 (identity, or past the end of the batch) points at a
// constant line (1, 0, 0) instead, so control flow stays uniform.  Result in bank A (= ML_F).
constexpr int PK_K = 4;
struct MillerSlotsPK { int A, B, L, T; int P[PK_K]; };
constexpr int PK_P0 = PP_P, PK_P1 = PP_P2, PK_P2 = PP_R2, PK_P3 = PP_Q2;   // global-memory slots (written under `if (ident)`)
static_assert(PK_P2 >= NS + NT_MAX && PK_P3 >= NS + NT_MAX, "P slots must be global-memory slots");

template <int K>
B381_DEV B381_INL void ark_bit_pk(const Ctx& cx, const MillerSlotsPK& s, int cur, int oth, bool square, bool add, const u4* const* line, const size_t* stride, int& idx) {
  static_assert((K & 1) == 0, "even K: the bank after a bit does not depend on the bit");
  if (square) {
    f12_sqr_oop(cx, oth, cur, s.T, s.L);
    const int sw = cur; cur = oth; oth = sw;
  }
  for (int pass = 0; pass < (add ? 2 : 1); pass++) {
#pragma unroll
    for (int j = 0; j < K; j++) {
      const u4* src = line[j] + stride[j] * idx;
      if (idx + 1 < G2PREP_TRIPLES) prefetch_triple(src + stride[j]);
      ark_ell_oop_ext(cx, oth, cur, src, s.L, s.P[j], pass == 0);
      const int sw = cur; cur = oth; oth = sw;
    }
    idx++;
  }
}

template <int K>
B381_DEV void ark_miller_loop_pk(const Ctx& cx, const MillerSlotsPK& s, const u4* const* line, const size_t* stride) {
  const uint64_t xabs = B381_X_ABS;
  int idx = 0;
  f12_set_one(cx, s.A);
  ark_bit_pk<K>(cx, s, s.A, s.B, false, (xabs >> 62) & 1, line, stride, idx);        // no squaring of f = 1: the bank is kept
  for (int b = 61; b >= 0; b -= 2) {
    ark_bit_pk<K>(cx, s, s.A, s.B, true, (xabs >> b) & 1, line, stride, idx);
    ark_bit_pk<K>(cx, s, s.B, s.A, true, (xabs >> (b - 1)) & 1, line, stride, idx);
  }
  f12_conj(cx, s.A);
}

// g1[j]: 24 words; pk[j]: first group of pair j's packed lines; inf[j] != 0: the pair contributes 1; one_line: this
// thread's view of the constant line (1, 0, 0).  Leaves the product of the K Miller values in ML_F.
B381_DEV B381_INL int miller_pk_to_slots(const Ctx& cx, const uint32_t* const* g1, const u4* const* pk, const int* inf, const u4* one_line) {
  int err = 0;
  MillerSlotsPK s;
  s.A = PP_A; s.B = PP_B; s.L = PP_L; s.T = PP_T;
  s.P[0] = PK_P0; s.P[1] = PK_P1; s.P[2] = PK_P2; s.P[3] = PK_P3;
  const u4* line[PK_K];
  size_t stride[PK_K];                             // uint4 distance between consecutive triples: 0 for the constant line
  for (int j = 0; j < PK_K; j++) {
    const bool ident = (inf[j] & 3) != 0;
    if (ident) f2_set_small(S_(s.P[j]), 0);
    else if (!f2_load_ext(S_(s.P[j]), g1[j])) err |= ERR_NOT_CANONICAL;
    line[j] = ident ? one_line : pk[j];
    stride[j] = ident ? 0 : (size_t)3 * SLOT;
  }
  ark_miller_loop_pk<PK_K>(cx, s, line, stride);
  return err;
}

// the constant line (1, 0, 0) for this thread (stride 0: every triple index reads the same three slots)
B381_DEV B381_INL void fill_one_line(u4* tile_lane) {
  f2_set_small(tile_lane, 1);
  f2_set_small(tile_lane + SLOT, 0);
  f2_set_small(tile_lane + 2 * SLOT, 0);
}

// Fq12 / Fq6 inverse (witness helpers): /root/reference/src/fields/fq12_target.rs:340-374,
// fq6_target.rs:384-418 run x.inverse() natively; formulas fq12_target_tree.rs:77-90, fq6_target_tree.rs:59-89
B381_DEV B381_INL int prog_f12_inv(const Ctx& cx, const uint32_t* in, uint32_t* out) {
  int err = 0;
  if (!f12_load_ext(cx, 0, in)) err |= ERR_NOT_CANONICAL;
  bool zero = true;
  for (int i = 0; i < 6; i++) zero = f2_is_zero(S_(i)) && zero;
  if (zero) err |= ERR_ZERO_DIVISION;
  f12_inv(cx, 0, 6);
  f12_store_ext(cx, out, 0);
  return err;
}

B381_DEV B381_INL int prog_f6_inv(const Ctx& cx, const uint32_t* in, uint32_t* out) {
  int err = 0;
  bool zero = true;
  for (int i = 0; i < 3; i++) {
    if (!f2_load_ext(S_(i), in + 24 * i)) err |= ERR_NOT_CANONICAL;
    zero = f2_is_zero(S_(i)) && zero;
  }
  if (zero) err |= ERR_ZERO_DIVISION;
  f6_inv(cx, 3, 0, 6);
  for (int i = 0; i < 3; i++) f2_store_ext(out + 24 * i, S_(3 + i));
  return err;
}

B381_DEV B381_INL int prog_final_exp(const Ctx& cx, const uint32_t* in, uint32_t* out) {
  int err = 0;
  if (!f12_load_ext(cx, FE_F, in)) err |= ERR_NOT_CANONICAL;
  err |= final_exp_slots(cx, FE_F);
  f12_store_ext(cx, out, FE_F);
  return err;
}

B381_DEV B381_INL int prog_pairing(const Ctx& cx, const uint32_t* g1, const uint32_t* g2, int inf, uint32_t* out, int mode) {
  int err = miller_to_slots(cx, g1, g2, inf, mode);
  err |= final_exp_slots(cx, ML_F);
  f12_store_ext(cx, out, FE_F);
  return err;
}

// Fp12 product of `cnt` raw values spaced `stride` words apart, result raw
B381_DEV B381_INL void prog_f12_product_raw(const Ctx& cx, const uint32_t* in, size_t cnt, size_t stride, uint32_t* out) {
  f12_load_raw(cx, 0, in);
  for (size_t i = 1; i < cnt; i++) {
    f12_load_raw(cx, 6, in + i * stride);
    f12_mul(cx, 0, 0, 6, 12, 18);
  }
  f12_store_raw(cx, out, 0);
}

// tower-order Fp12 multiply on external buffers (config #2 microbench, b381_fp12_mul).  The operands are used as
// they are and the product is put right on the way out (fp32.cuh fp_from_ext_asis): 12 conversions instead of 36.
B381_DEV B381_INL int prog_f12_mul(const Ctx& cx, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  int err = 0;
  for (int i = 0; i < 6; i++) {
    if (!f2_load_asis2(S_(i), a + 24 * i, a + 24 * i + 12)) err |= ERR_NOT_CANONICAL;
    if (!f2_load_asis2(S_(6 + i), b + 24 * i, b + 24 * i + 12)) err |= ERR_NOT_CANONICAL;
  }
  f12_mul(cx, 0, 0, 6, 12, 18);
  for (int i = 0; i < 6; i++) f2_store_unscale2(out + 24 * i, out + 24 * i + 12, S_(i));
  return err;
}

// MyFq12 (w-basis) multiply: /root/reference/src/fields/helpers.rs:90-152.  The reference's
// schoolbook formula and the tower product are the same field element (its own test
// helpers.rs:248-267 asserts it), so the product is computed in the tower; the coefficient map of
// helpers.rs:39-41 only decides where each half of a tower slot is read and written:
// coeffs = [c000,c100,c010,c110,c020,c120, c001,c101,c011,c111,c021,c121]; c_ijk: i = Fp6 half, j = Fp2 idx, k = u,
// so tower slot 3 i + j has its two halves at w-basis indices 2 j + i and 6 + 2 j + i.
B381_DEV B381_INL int prog_wbasis_mul(const Ctx& cx, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  int err = 0;
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 3; j++) {
      const int s = 3 * i + j, w = 2 * j + i;
      if (!f2_load_asis2(S_(s), a + 12 * w, a + 12 * (6 + w))) err |= ERR_NOT_CANONICAL;
      if (!f2_load_asis2(S_(6 + s), b + 12 * w, b + 12 * (6 + w))) err |= ERR_NOT_CANONICAL;
    }
  f12_mul(cx, 0, 0, 6, 12, 18);
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 3; j++) {
      const int s = 3 * i + j, w = 2 * j + i;
      f2_store_unscale2(out + 12 * w, out + 12 * (6 + w), S_(s));
    }
  return err;
}

// LITERAL optimized_miller_loop on (G1Projective x,y,z: 36 words; G2Projective x,y,z: 72 words).
// Output Fq12 with only c0.c0 populated (/root/reference/src/miller_loop_native_optimized.rs:18-36).
B381_DEV B381_INL int prog_literal(const Ctx& cx, const uint32_t* g1p, const uint32_t* g2p, uint32_t* out) {
  int err = 0;
  uint32_t w[24];
  for (int c = 0; c < 3; c++) {                    // xp, yp, zp as Fq2::new(v, 0)
    for (int k = 0; k < 12; k++) { w[k] = g1p[12 * c + k]; w[12 + k] = 0; }
    if (!f2_load_ext(S_(LT_P + c), w)) err |= ERR_NOT_CANONICAL;
    if (!f2_load_ext(S_(LT_Q + c), g2p + 24 * c)) err |= ERR_NOT_CANONICAL;
  }
  LiteralSlots s;
  s.R = LT_R; s.Q = LT_Q; s.Pp = LT_P; s.fn = LT_FN; s.fd = LT_FD; s.n = LT_N; s.d = LT_D; s.T = LT_T;
  bool ok = literal_optimized_miller_loop(cx, s, LT_OUT);
  for (int k = 0; k < 144; k++) out[k] = 0;
  if (!ok) return err | ERR_ZERO_DIVISION;
  f2_store_ext(out, S_(LT_OUT));
  return err;
}

// ---- G2 endomorphism programs: subgroup membership and cofactor clearing -----------------
// ark-bls12-381 0.4 curves/g2.rs, restated in the test oracle (g2_psi, g2_in_subgroup_fast,
// g2_clear_cofactor).  The group law is the ark-ec Jacobian add / double of tower.cuh (jac_add / jac_double), the
// scalar |x| is public: uniform ladders apart from the special cases inside the formulas.  (G1: g1.cuh.)

// R <- [|x|] Base, Base Jacobian in slots Base..Base+2 (63 doublings + 5 additions); T = 9 scratch slots
B381_DEV B381_INL void jac_mul_x_abs(const Ctx& cx, int R, int Base, int T) {
  for (int k = 0; k < 3; k++) lin(cx, R + k, Base + k, -1, L_COPY);
  const uint64_t xabs = B381_X_ABS;
  for (int b = 62; b >= 0; b--) {
    jac_double(cx, R, T);
    if ((xabs >> b) & 1) jac_add(cx, R, Base, T);
  }
}
B381_DEV B381_INL void jac_neg(const Ctx& cx, int R) { lin(cx, R + 1, R + 1, -1, L_NEG); }

// psi(Q) for affine Q in slots Q, Q+1 -> slots D, D+1 (Z = 1 in D+2)
B381_DEV B381_INL void g2_psi_affine(const Ctx& cx, int D, int Q) {
  f2_mul_psi(S_(D), S_(Q), 0, 1);
  f2_mul_psi(S_(D + 1), S_(Q + 1), 1, 1);
  f2_set_small(S_(D + 2), 1);
}

// is the Jacobian point in R equal to the AFFINE point in A, A+1?  X == ax Z^2 and Y == ay Z^3; t = 3 scratch slots
B381_DEV B381_INL bool jac_equals_affine(const Ctx& cx, int R, int A, int t) {
  if (f2_is_zero(S_(R + 2))) return false;
  sqr(cx, t, R + 2);
  mul(cx, t + 1, A, t);
  if (!f2_equal(S_(t + 1), S_(R))) return false;
  mul(cx, t, t, R + 2);
  mul(cx, t + 1, A + 1, t);
  return f2_equal(S_(t + 1), S_(R + 1));
}

// ark g2.rs is_in_correct_subgroup_assuming_on_curve (eprint 2021/1130 section 4): psi(P) == [x] P, x = -|x|
B381_DEV B381_INL int prog_g2_in_subgroup(const Ctx& cx, const uint32_t* pt, int inf, uint8_t* out) {
  int err = 0;
  if (inf & 1) { *out = 1; return 0; }              // the identity is in every subgroup
  const int Q = 0, R = 3, PS = 6, T = 9;            // Q (affine, Z = 1), [|x|] Q, psi(Q), 9 scratch slots
  if (!f2_load_ext(S_(Q), pt)) err |= ERR_NOT_CANONICAL;
  if (!f2_load_ext(S_(Q + 1), pt + 24)) err |= ERR_NOT_CANONICAL;
  f2_set_small(S_(Q + 2), 1);
  jac_mul_x_abs(cx, R, Q, T);
  jac_neg(cx, R);                                   // [x] Q
  g2_psi_affine(cx, PS, Q);
  *out = jac_equals_affine(cx, R, PS, T) ? 1 : 0;
  return err;
}

// Jacobian (X, Y, Z) in slots R..R+2 -> affine (X / Z^2, Y / Z^3) in the C-ABI layout + identity flag;
// t = 2 scratch slots
B381_DEV B381_INL void jac_store_affine(const Ctx& cx, int R, int t, int is_g2, uint32_t* out, uint8_t* out_inf) {
  const int w = is_g2 ? 48 : 24, ZI = t, ZI2 = t + 1;
  if (f2_is_zero(S_(R + 2))) {
    for (int i = 0; i < w; i++) out[i] = 0;
    *out_inf = 1;
    return;
  }
  *out_inf = 0;
  f2_inv(S_(ZI), S_(R + 2));
  sqr(cx, ZI2, ZI);
  mul(cx, R, R, ZI2);
  mul(cx, ZI2, ZI2, ZI);
  mul(cx, R + 1, R + 1, ZI2);
  if (is_g2) {
    f2_store_ext(out, S_(R));
    f2_store_ext(out + 24, S_(R + 1));
  } else {
    uint32_t wd[24];
    f2_store_ext(wd, S_(R));
    for (int j = 0; j < 12; j++) out[j] = wd[j];
    f2_store_ext(wd, S_(R + 1));
    for (int j = 0; j < 12; j++) out[12 + j] = wd[j];
  }
}

// affine point (C-ABI layout) -> slots Q, Q+1 with Z = 1 in Q+2; G1 coordinates are embedded in Fq2
B381_DEV B381_INL int load_affine_point(const Ctx& cx, int Q, const uint32_t* pt, int is_g2) {
  int err = 0;
  if (is_g2) {
    if (!f2_load_ext(S_(Q), pt)) err |= ERR_NOT_CANONICAL;
    if (!f2_load_ext(S_(Q + 1), pt + 24)) err |= ERR_NOT_CANONICAL;
  } else {
    uint32_t wd[24];
    for (int c = 0; c < 2; c++) {
      for (int j = 0; j < 12; j++) { wd[j] = pt[12 * c + j]; wd[12 + j] = 0; }
      if (!f2_load_ext(S_(Q + c), wd)) err |= ERR_NOT_CANONICAL;
    }
  }
  f2_set_small(S_(Q + 2), 1);
  return err;
}

// G2: sum of cnt affine points (48 words each, identity flags in a byte array, may be null) -> one affine point +
// flag.  One level of the reduction tree behind b381_g2_sum and the G2 multi-scalar multiplication.
B381_DEV B381_INL int prog_g2_point_sum(const Ctx& cx, const uint32_t* in, const uint8_t* in_inf, size_t cnt, uint32_t* out, uint8_t* out_inf) {
  int err = 0;
  const int Q = 0, R = 3, T = 6, ZI = 15;
  f2_set_small(S_(R), 1); f2_set_small(S_(R + 1), 1); f2_set_small(S_(R + 2), 0);     // identity (1, 1, 0)
  for (size_t i = 0; i < cnt; i++) {
    if (in_inf && (in_inf[i] & 1)) continue;
    err |= load_affine_point(cx, Q, in + 48 * i, 1);
    jac_add(cx, R, Q, T);
  }
  jac_store_affine(cx, R, ZI, 1, out, out_inf);
  return err;
}

// ---- G2 bucket method (Pippenger) over the slot arena -----------------------------------------------------------
// Same stages as the G1 pipeline (g1_kernels.cu; the digit histogram / scan / scatter kernels are shared, they never
// look at a point): one thread per bucket sums its points, one thread per chunk of buckets forms the weighted chunk
// sum by running sums, two levels of plain sums per window, Horner over the windows.  Intermediate points travel as
// raw Jacobian triples: 3 slots x 24 words in the internal format.  Slot plan: Q 0..2, R 3..5, T 6..14 (jac_add's nine
// scratch slots), R2 15..17, R3 18..20.
constexpr int G2_RAW_JAC = 72;
constexpr int GM_Q = 0, GM_R = 3, GM_T = 6, GM_R2 = 15, GM_R3 = 18;

B381_DEV B381_INL void jac_store_raw(const Ctx& cx, uint32_t* dst, int R) {
  for (int i = 0; i < 3; i++) {
    Fp c0, c1;
    ld_f2(c0, c1, S_(R + i));
    B381_CHECK(c0.mag <= 4.2 && c1.mag <= 4.2, "raw Jacobian store: coordinate above the bound jac_load_raw claims");
    for (int k = 0; k < SW; k++) { dst[24 * i + k] = (uint32_t)c0.l[k]; dst[24 * i + SW + k] = (uint32_t)c1.l[k]; }
  }
}
B381_DEV B381_INL void jac_load_raw(const Ctx& cx, int R, const uint32_t* src) {
  for (int i = 0; i < 3; i++) {
    Fp c0, c1;
    fp_zero(c0); fp_zero(c1);
    for (int k = 0; k < SW; k++) { c0.l[k] = (limb_t)src[24 * i + k]; c1.l[k] = (limb_t)src[24 * i + SW + k]; }
    B381_SETRANGE(c0, 0.0, 4.2); B381_SETRANGE(c1, 0.0, 4.2);     // outputs of jac_add / jac_double: Z = 2 Y Z < 4.2 p (checked at the store)
    st_f2(S_(R + i), c0, c1);
  }
}
B381_DEV B381_INL void jac_set_identity(const Ctx& cx, int R) {     // (1, 1, 0), ark-ec 0.4
  f2_set_small(S_(R), 1); f2_set_small(S_(R + 1), 1); f2_set_small(S_(R + 2), 0);
}

// sum of the affine points pts[idx[lo .. hi)] (C-ABI layout, 48 words) -> raw Jacobian
B381_DEV B381_INL int prog_g2_bucket_sum(const Ctx& cx, const uint32_t* pts, const uint32_t* idx, size_t lo, size_t hi, uint32_t* dst) {
  int err = 0;
  jac_set_identity(cx, GM_R);
  for (size_t t = lo; t < hi; t++) {
    const uint32_t* pt = pts + (size_t)48 * idx[t];
    if (!f2_load_ext(S_(GM_Q), pt)) err |= ERR_NOT_CANONICAL;
    if (!f2_load_ext(S_(GM_Q + 1), pt + 24)) err |= ERR_NOT_CANONICAL;
    jac_add_mixed(cx, GM_R, GM_Q, GM_T);
  }
  jac_store_raw(cx, dst, GM_R);
  return err;
}

// chunk [lo, hi) of one window's buckets: sum_{d in chunk} d B_d = running sums + (lo - 1) x (sum of the chunk)
B381_DEV B381_INL void prog_g2_chunk_weighted(const Ctx& cx, const uint32_t* buckets, uint32_t lo, uint32_t hi, uint32_t* dst) {
  jac_set_identity(cx, GM_R);                       // running sum
  jac_set_identity(cx, GM_R2);                      // accumulated
  for (uint32_t d = hi; d-- > lo;) {
    jac_load_raw(cx, GM_Q, buckets + (size_t)G2_RAW_JAC * d);
    jac_add(cx, GM_R, GM_Q, GM_T);
    jac_add(cx, GM_R2, GM_R, GM_T);
  }
  const uint32_t m = lo - 1;
  if (m != 0 && lo < hi) {
    jac_set_identity(cx, GM_R3);
    bool started = false;
    for (int bit = 31; bit >= 0; bit--) {
      if (started) jac_double(cx, GM_R3, GM_T);
      if ((m >> bit) & 1u) { jac_add(cx, GM_R3, GM_R, GM_T); started = true; }
    }
    jac_add(cx, GM_R2, GM_R3, GM_T);
  }
  jac_store_raw(cx, dst, GM_R2);
}

// sum of the raw Jacobian points in[lo .. hi)
B381_DEV B381_INL void prog_g2_jac_sum(const Ctx& cx, const uint32_t* in, size_t lo, size_t hi, uint32_t* dst) {
  jac_set_identity(cx, GM_R);
  for (size_t t = lo; t < hi; t++) {
    jac_load_raw(cx, GM_Q, in + (size_t)G2_RAW_JAC * t);
    jac_add(cx, GM_R, GM_Q, GM_T);
  }
  jac_store_raw(cx, dst, GM_R);
}

// Horner over the W window sums (c doublings per window) and conversion to affine
B381_DEV B381_INL void prog_g2_msm_final(const Ctx& cx, const uint32_t* sums, int W, int c, uint32_t* out48, uint8_t* out_inf) {
  jac_load_raw(cx, GM_R, sums + (size_t)G2_RAW_JAC * (W - 1));
  for (int w = W - 2; w >= 0; w--) {
    for (int i = 0; i < c; i++) jac_double(cx, GM_R, GM_T);
    jac_load_raw(cx, GM_Q, sums + (size_t)G2_RAW_JAC * w);
    jac_add(cx, GM_R, GM_Q, GM_T);
  }
  jac_store_affine(cx, GM_R, GM_R2, 1, out48, out_inf);
}

// ark g2.rs clear_cofactor (Budroni-Pintore, eprint 2017/419 section 4.1):
//   [x^2 - x - 1] P + [x - 1] psi(P) + psi^2(2 P)  =  psi^2(2P) + (-[X](-[X] P + psi P)) - (-[X] P) - psi P - P,  X = |x|
B381_DEV B381_INL int prog_g2_clear_cofactor(const Ctx& cx, const uint32_t* pt, int inf, uint32_t* out, uint8_t* out_inf) {
  int err = 0;
  if (inf & 1) { for (int i = 0; i < 48; i++) out[i] = 0; *out_inf = 1; return 0; }
  const int P0 = 0, XP = 3, PS = 6, P2 = 9, TM = 12, AC = 15, T = 18, ZI = 27;
  if (!f2_load_ext(S_(P0), pt)) err |= ERR_NOT_CANONICAL;
  if (!f2_load_ext(S_(P0 + 1), pt + 24)) err |= ERR_NOT_CANONICAL;
  f2_set_small(S_(P0 + 2), 1);
  jac_mul_x_abs(cx, XP, P0, T);
  jac_neg(cx, XP);                                  // x_p = [x] P
  g2_psi_affine(cx, PS, P0);                        // psi_p
  for (int k = 0; k < 3; k++) lin(cx, P2 + k, P0 + k, -1, L_COPY);
  jac_double(cx, P2, T);                            // 2 P ; psi^2 (X, Y, Z) = (c X, -Y, Z), c in Fq
  f2_mul_psi(S_(P2), S_(P2), 2, 0);
  jac_neg(cx, P2);                                  // psi2_p2
  for (int k = 0; k < 3; k++) lin(cx, TM + k, XP + k, -1, L_COPY);
  jac_add(cx, TM, PS, T);                           // tmp = x_p + psi_p
  jac_mul_x_abs(cx, AC, TM, T);
  jac_neg(cx, AC);                                  // tmp2 = [x] tmp = [x^2] P + [x] psi(P)
  jac_add(cx, AC, P2, T);                           // + psi2_p2
  jac_neg(cx, XP); jac_add(cx, AC, XP, T);          // - x_p
  jac_neg(cx, PS); jac_add(cx, AC, PS, T);          // - psi_p
  jac_neg(cx, P0); jac_add(cx, AC, P0, T);          // - P
  jac_store_affine(cx, AC, ZI, 1, out, out_inf);
  return err;
}

// ---- scalar multiplication (first half): out = [k] P, k = 256-bit scalar (8 LE words) ----
// Left-to-right double-and-add over the Jacobian group law the reference's native loop already uses
// (`R + R`, `R + Q` at /root/reference/src/miller_loop_native_optimized.rs:93,98 = ark-ec Projective add /
// double, tower.cuh jac_add / jac_double), then one inversion back to affine.  Per-thread scalars:
// control flow diverges, so the kernel runs without the lock-step barriers.  Not constant time.
B381_DEV B381_INL int prog_scalar_mul(const Ctx& cx, const uint32_t* pt, int is_g2, int inf, const uint32_t* k, uint32_t* out, uint8_t* out_inf) {
  int err = 0;
  const int Q = 0, R = 3, T = 6, ZI = 15, ZI2 = 16;
  const int w = is_g2 ? 48 : 24;
  uint32_t kk[8], nz = 0;
  for (int i = 0; i < 8; i++) { kk[i] = k[i]; nz |= kk[i]; }
  if ((inf & 1) || nz == 0) {
    for (int i = 0; i < w; i++) out[i] = 0;
    *out_inf = 1;
    return 0;
  }
  if (is_g2) {
    if (!f2_load_ext(S_(Q), pt)) err |= ERR_NOT_CANONICAL;
    if (!f2_load_ext(S_(Q + 1), pt + 24)) err |= ERR_NOT_CANONICAL;
  } else {
    uint32_t wd[24];
    for (int c = 0; c < 2; c++) {
      for (int j = 0; j < 12; j++) { wd[j] = pt[12 * c + j]; wd[12 + j] = 0; }
      if (!f2_load_ext(S_(Q + c), wd)) err |= ERR_NOT_CANONICAL;
    }
  }
  f2_set_small(S_(Q + 2), 1);
  f2_set_small(S_(R), 0); f2_set_small(S_(R + 1), 1); f2_set_small(S_(R + 2), 0);     // identity (Z = 0)
  bool started = false;
  for (int b = 255; b >= 0; b--) {
    const bool bit = (kk[b >> 5] >> (b & 31)) & 1u;
    if (started) jac_double(cx, R, T);
    if (bit) {
      if (started) jac_add(cx, R, Q, T);
      else { for (int c = 0; c < 3; c++) lin(cx, R + c, Q + c, -1, L_COPY); started = true; }
    }
  }
  jac_store_affine(cx, R, ZI, is_g2, out, out_inf);
  return err;
}

#undef S_

}  // namespace b381
