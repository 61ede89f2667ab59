// hostsim.cpp -- TEST-ONLY host build of the device arithmetic (fp32.cuh | fp28.cuh / tower.cuh / programs.cuh).
// The CUDA sources are plain C++ on the host (the PTX carry-chain macros of fp32.cuh expand to uint64
// arithmetic with an explicit carry variable), so compiling them with g++ gives a bit-exact
// simulation of what every GPU thread computes.  tests/ use it on the CPU box to check the device
// algorithms against the oracle without a GPU, and (built with -DB381_TRACK_BOUNDS) to verify the
// worst-case limb / column / magnitude bounds of the lazy arithmetic along the executed path.
// It is NOT part of the product: libb381.so never links or loads it.
#include <cstring>
#include <vector>
#include "../../plonky2-bls12-381-pairing_b200/csrc/programs.cuh"
#include "../../plonky2-bls12-381-pairing_b200/csrc/helpers.cuh"
#include "../../plonky2-bls12-381-pairing_b200/csrc/g1.cuh"

using namespace b381;

static thread_local u4 g_arena[(MAX_NSLOTS + 8) * GPS];

static Ctx make_ctx() {
  Ctx cx;
  cx.sm = g_arena;
  cx.gm = g_arena + NS * SLOT;
  cx.sync = 0;
#ifdef B381_TRACK_BOUNDS
  track_tab().clear();
#endif
  return cx;
}

extern "C" {

int hs_fp_mul(const uint32_t* a, const uint32_t* b, uint32_t* out) {
  // Montgomery product in the external domain: out = a * b / 2^384, via internal conversion
  uint32_t wa[12], wb[12], wo[12];
  memcpy(wa, a, 48); memcpy(wb, b, 48);
  Fp x, y, z;
  bool ok = fp_from_ext(x, wa);
  ok &= fp_from_ext(y, wb);
  fp_mul(z, x, y);
  fp_to_ext(wo, z);
  memcpy(out, wo, 48);
  return ok ? 0 : 1;
}

int hs_fp_roundtrip(const uint32_t* a, uint32_t* out) {
  uint32_t wa[12], wo[12];
  memcpy(wa, a, 48);
  Fp x;
  bool ok = fp_from_ext(x, wa);
  fp_to_ext(wo, x);
  memcpy(out, wo, 48);
  return ok ? 0 : 1;
}

int hs_fp2_mul(const uint32_t* a, const uint32_t* b, uint32_t* out) {
  Ctx cx = make_ctx();
  bool ok = f2_load_ext(slot(cx, 0), a);
  ok &= f2_load_ext(slot(cx, 1), b);
  f2_mul(slot(cx, 2), slot(cx, 0), slot(cx, 1));
  f2_store_ext(out, slot(cx, 2));
  return ok ? 0 : 1;
}

int hs_fp2_inv(const uint32_t* a, uint32_t* out) {
  Ctx cx = make_ctx();
  bool ok = f2_load_ext(slot(cx, 0), a);
  f2_inv(slot(cx, 1), slot(cx, 0));
  f2_store_ext(out, slot(cx, 1));
  return ok ? 0 : 1;
}

int hs_fp12_mul(const uint32_t* a, const uint32_t* b, uint32_t* out) { Ctx cx = make_ctx(); return prog_f12_mul(cx, a, b, out); }
int hs_fp12_mul_wbasis(const uint32_t* a, const uint32_t* b, uint32_t* out) { Ctx cx = make_ctx(); return prog_wbasis_mul(cx, a, b, out); }

int hs_fp12_sqr(const uint32_t* a, uint32_t* out) {
  Ctx cx = make_ctx();
  bool ok = f12_load_ext(cx, 0, a);
  f12_sqr(cx, 0, 6, 13);
  f12_store_ext(cx, out, 0);
  return ok ? 0 : 1;
}

int hs_fp12_inv(const uint32_t* a, uint32_t* out) {
  Ctx cx = make_ctx();
  bool ok = f12_load_ext(cx, 0, a);
  f12_inv(cx, 0, 6);
  f12_store_ext(cx, out, 0);
  return ok ? 0 : 1;
}

int hs_fp12_frobenius(const uint32_t* a, int k, uint32_t* out) {
  Ctx cx = make_ctx();
  bool ok = f12_load_ext(cx, 0, a);
  f12_frobenius(cx, 0, k);
  f12_store_ext(cx, out, 0);
  return ok ? 0 : 1;
}

int hs_fp12_cyclotomic_square(const uint32_t* a, uint32_t* out) {
  Ctx cx = make_ctx();
  bool ok = f12_load_ext(cx, 0, a);
  f12_cyclotomic_square(cx, 6, 0);
  f12_store_ext(cx, out, 6);
  return ok ? 0 : 1;
}

int hs_fp12_mul_by_014(const uint32_t* f, const uint32_t* c0, const uint32_t* c1, const uint32_t* c4, uint32_t* out) {
  Ctx cx = make_ctx();
  bool ok = f12_load_ext(cx, 0, f);
  ok &= f2_load_ext(slot(cx, 6), c0);
  ok &= f2_load_ext(slot(cx, 7), c1);
  ok &= f2_load_ext(slot(cx, 8), c4);
  f12_mul_by_014(cx, 0, 6, 7, 8, 9);
  f12_store_ext(cx, out, 0);
  return ok ? 0 : 1;
}

int hs_miller_loop(const uint32_t* g1, const uint32_t* g2, int inf, uint32_t* out, int mode) { Ctx cx = make_ctx(); return prog_miller(cx, g1, g2, inf, out, mode); }
int hs_final_exp(const uint32_t* in, uint32_t* out) { Ctx cx = make_ctx(); return prog_final_exp(cx, in, out); }
int hs_pairing(const uint32_t* g1, const uint32_t* g2, int inf, uint32_t* out, int mode) { Ctx cx = make_ctx(); return prog_pairing(cx, g1, g2, inf, out, mode); }
int hs_literal(const uint32_t* g1p, const uint32_t* g2p, uint32_t* out) { Ctx cx = make_ctx(); return prog_literal(cx, g1p, g2p, out); }

// product of n Miller values, MK = four pairs at a time with shared squarings (the k_multi_miller path)
int hs_multi_miller(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, size_t n, uint32_t* out, int mode) {
  Ctx cx = make_ctx();
  int err = 0;
  f12_set_one(cx, M2_ACC);
  for (size_t i = 0; i < n; i += MK) {
    const uint32_t* pg1[MK]; const uint32_t* pg2[MK]; int pinf[MK];
    for (int j = 0; j < MK; j++) {
      const bool act = i + j < n;
      const size_t k = act ? i + j : n - 1;
      pg1[j] = g1 + 24 * k; pg2[j] = g2 + 48 * k;
      pinf[j] = act ? (inf ? inf[k] : 0) : 3;
    }
    err |= miller_multi_to_slots(cx, pg1, pg2, pinf, mode);
    f12_mul(cx, M2_ACC, M2_ACC, ML_F, M2_SCRATCH, M2_SCRATCH + 6);   // scratch: slots that are dead once the loop has left f in ML_F
  }
  f12_store_ext(cx, out, M2_ACC);
  return err;
}

int hs_g2_prepare(const uint32_t* g2, uint32_t* coeffs, int mode) { Ctx cx = make_ctx(); return prog_g2_prepare(cx, g2, coeffs, mode); }
int hs_miller_prepared(const uint32_t* g1, const uint32_t* coeffs, int inf, uint32_t* out, int mode, int do_fe) {
  Ctx cx = make_ctx();
  return prog_miller_prepared(cx, g1, coeffs, inf, out, mode, do_fe);
}

// packed G2Prepared (internal format; on the host a tile is one point: 68 x 3 slots of 6 uint4 = 4896 words)
int hs_g2_prepare_packed(const uint32_t* g2, uint32_t* packed, int mode) { Ctx cx = make_ctx(); return prog_g2_prepare_packed(cx, g2, reinterpret_cast<u4*>(packed), mode); }
int hs_miller_packed(const uint32_t* g1, const uint32_t* packed, int inf, uint32_t* out, int mode, int do_fe) {
  Ctx cx = make_ctx();
  return prog_miller_packed(cx, g1, reinterpret_cast<const u4*>(packed), inf, out, mode, do_fe);
}

// product of n Miller values against packed prepared Q's, four pairs at a time (the k_multi_miller_packed path);
// packed: n host tiles of 4896 words
int hs_multi_miller_packed(const uint32_t* g1, const uint32_t* packed, const uint8_t* inf, size_t n, uint32_t* out) {
  Ctx cx = make_ctx();
  alignas(16) static u4 one_line[3 * GPS];
  fill_one_line(one_line);
  int err = 0;
  f12_set_one(cx, M2_ACC);
  for (size_t i = 0; i < n; i += PK_K) {
    const uint32_t* pg1[PK_K]; const u4* ppk[PK_K]; int pinf[PK_K];
    for (int j = 0; j < PK_K; j++) {
      const bool act = i + j < n;
      const size_t k = act ? i + j : n - 1;
      pg1[j] = g1 + 24 * k;
      ppk[j] = reinterpret_cast<const u4*>(packed) + g2pack_index(k);
      pinf[j] = act ? (inf ? inf[k] : 0) : 3;
    }
    err |= miller_pk_to_slots(cx, pg1, ppk, pinf, one_line);
    f12_mul(cx, M2_ACC, M2_ACC, ML_F, M2_SCRATCH, M2_SCRATCH + 6);
  }
  f12_store_ext(cx, out, M2_ACC);
  return err;
}

int hs_fp_inv(const uint32_t* a, uint32_t* out) { return prog_fp_inv(a, out); }
int hs_fp_pow(const uint32_t* a, const uint32_t* e, int nwords, uint32_t* out) { return prog_fp_pow(a, e, nwords, out); }
int hs_fp_is_square(const uint32_t* a, uint8_t* out) { return prog_fp_is_square(a, out); }
int hs_fp_sqrt(const uint32_t* a, int sgn, uint32_t* out) { return prog_fp_sqrt(a, sgn, out); }
int hs_fp2_inv_ext(const uint32_t* a, uint32_t* out) { return prog_fp2_inv(a, out); }
int hs_fp2_sqrt(const uint32_t* a, int sgn, uint32_t* out) { return prog_fp2_sqrt(a, sgn, out); }
int hs_fp2_is_square(const uint32_t* a, uint8_t* out) { return prog_fp2_is_square(a, out); }
int hs_fp_to_digits(const uint32_t* a, uint32_t* out) { return prog_fp_to_digits(a, out); }
int hs_fp_from_digits(const uint32_t* d, uint32_t* out) { return prog_fp_from_digits(d, out); }
int hs_fp12_to_witness(const uint32_t* f, uint32_t* out) { return prog_fp12_to_witness(f, out); }
int hs_g1_deserialize(const uint8_t* in, int compressed, uint32_t* g1, uint8_t* inf) { return prog_g1_deserialize(in, compressed, g1, inf); }
int hs_g1_serialize(const uint32_t* g1, int inf, int compressed, uint8_t* out) { return prog_g1_serialize(g1, inf, compressed, out); }
int hs_g2_deserialize(const uint8_t* in, int compressed, uint32_t* g2, uint8_t* inf) { return prog_g2_deserialize(in, compressed, g2, inf); }
int hs_g2_serialize(const uint32_t* g2, int inf, int compressed, uint8_t* out) { return prog_g2_serialize(g2, inf, compressed, out); }
int hs_fp12_inv_ext(const uint32_t* a, uint32_t* out) { Ctx cx = make_ctx(); return prog_f12_inv(cx, a, out); }
int hs_fp6_inv_ext(const uint32_t* a, uint32_t* out) { Ctx cx = make_ctx(); return prog_f6_inv(cx, a, out); }

int hs_subgroup_check(const uint32_t* pt, int is_g2, int inf, uint8_t* out) {
  if (!is_g2) return prog_g1_in_subgroup(pt, inf, out);
  Ctx cx = make_ctx();
  return prog_g2_in_subgroup(cx, pt, inf, out);
}

int hs_clear_cofactor(const uint32_t* pt, int is_g2, int inf, uint32_t* out, uint8_t* out_inf) {
  if (!is_g2) return prog_g1_clear_cofactor(pt, inf, out, out_inf);
  Ctx cx = make_ctx();
  return prog_g2_clear_cofactor(cx, pt, inf, out, out_inf);
}

int hs_scalar_mul(const uint32_t* pt, int is_g2, int inf, const uint32_t* k, uint32_t* out, uint8_t* out_inf) {
  if (!is_g2) return prog_g1_scalar_mul(pt, inf, k, out, out_inf);
  Ctx cx = make_ctx();
  return prog_scalar_mul(cx, pt, is_g2, inf, k, out, out_inf);
}

int hs_g2_point_sum(const uint32_t* pts, const uint8_t* inf, size_t cnt, uint32_t* out, uint8_t* out_inf) {
  Ctx cx = make_ctx();
  return prog_g2_point_sum(cx, pts, inf, cnt, out, out_inf);
}

// the G1 bucket method exactly as the kernels of kernels.cu stage it (k_g1_to_raw, k_msm_digits x 2, k_msm_scan,
// k_msm_bucket_sums, k_msm_chunks, k_g1_jac_sums, k_msm_final), run sequentially on the host
int hs_g1_msm(const uint32_t* pts, const uint8_t* inf, const uint32_t* scalars, size_t n, int c, int CH, uint32_t* out24, uint8_t* out_inf) {
  int err = 0;
  const int W = (256 + c - 1) / c;
  const uint32_t B = 1u << c;
  std::vector<uint32_t> raw(n * G1_RAW_AFF);
  for (size_t i = 0; i < n; i++) { G1A p; err |= g1_load_ext(p, pts + 24 * i); g1_st_raw_aff(raw.data() + G1_RAW_AFF * i, p); }
  const size_t m = (size_t)W * B;
  std::vector<unsigned int> cnt(m, 0), start(m + 1, 0), cursor(m, 0);
  for (size_t i = 0; i < n; i++) {
    if (inf && (inf[i] & 1)) continue;
    for (int w = 0; w < W; w++) { uint32_t d = msm_digit(scalars + 8 * i, w, c); if (d) cnt[(size_t)w * B + d]++; }
  }
  for (size_t b = 0; b < m; b++) { start[b + 1] = start[b] + cnt[b]; cursor[b] = start[b]; }
  std::vector<uint32_t> idx(start[m] ? start[m] : 1);
  for (size_t i = 0; i < n; i++) {
    if (inf && (inf[i] & 1)) continue;
    for (int w = 0; w < W; w++) { uint32_t d = msm_digit(scalars + 8 * i, w, c); if (d) idx[cursor[(size_t)w * B + d]++] = (uint32_t)i; }
  }
  std::vector<uint32_t> buckets(m * G1_RAW_JAC);
  for (size_t b = 0; b < m; b++) { G1J acc; msm_bucket_sum(acc, raw.data(), idx.data(), start[b], start[b + 1]); g1_st_raw_jac(buckets.data() + G1_RAW_JAC * b, acc); }
  const uint32_t nchunk = (B + CH - 1) / CH;
  std::vector<uint32_t> partial((size_t)W * nchunk * G1_RAW_JAC), sums((size_t)W * G1_RAW_JAC);
  for (int w = 0; w < W; w++)
    for (uint32_t j = 0; j < nchunk; j++) {
      uint32_t lo = j * CH, hi = lo + CH < B ? lo + CH : B;
      if (lo == 0) lo = 1;
      G1J r;
      if (lo < hi) msm_chunk_weighted(r, buckets.data() + (size_t)G1_RAW_JAC * w * B, lo, hi); else g1_set_identity(r);
      g1_st_raw_jac(partial.data() + G1_RAW_JAC * ((size_t)w * nchunk + j), r);
    }
  for (int w = 0; w < W; w++) {
    G1J acc; g1_set_identity(acc);
    for (uint32_t j = 0; j < nchunk; j++) { G1J q; g1_ld_raw_jac(q, partial.data() + G1_RAW_JAC * ((size_t)w * nchunk + j)); g1_add(acc, acc, q); }
    g1_st_raw_jac(sums.data() + G1_RAW_JAC * w, acc);
  }
  G1J r;
  msm_combine_windows(r, sums.data(), W, c);
  g1_store_jac_ext(out24, out_inf, r);
  return err;
}

// the G2 bucket method exactly as the kernels stage it (k_msm_digits x 2, k_msm_scan, k_g2_msm_bucket_sums, k_g2_msm_chunks,
// k_g2_jac_sums, k_g2_msm_final), run sequentially on the host
int hs_g2_msm(const uint32_t* pts, const uint8_t* inf, const uint32_t* scalars, size_t n, int c, int CH, uint32_t* out48, uint8_t* out_inf) {
  Ctx cx = make_ctx();
  int err = 0;
  const int W = (256 + c - 1) / c;
  const uint32_t B = 1u << c;
  const size_t m = (size_t)W * B;
  std::vector<unsigned int> cnt(m, 0), start(m + 1, 0), cursor(m, 0);
  for (size_t i = 0; i < n; i++) {
    if (inf && (inf[i] & 1)) continue;
    for (int w = 0; w < W; w++) { uint32_t d = msm_digit(scalars + 8 * i, w, c); if (d) cnt[(size_t)w * B + d]++; }
  }
  for (size_t b = 0; b < m; b++) { start[b + 1] = start[b] + cnt[b]; cursor[b] = start[b]; }
  std::vector<uint32_t> idx(start[m] ? start[m] : 1);
  for (size_t i = 0; i < n; i++) {
    if (inf && (inf[i] & 1)) continue;
    for (int w = 0; w < W; w++) { uint32_t d = msm_digit(scalars + 8 * i, w, c); if (d) idx[cursor[(size_t)w * B + d]++] = (uint32_t)i; }
  }
  std::vector<uint32_t> buckets(m * G2_RAW_JAC);
  for (size_t b = 0; b < m; b++) err |= prog_g2_bucket_sum(cx, pts, idx.data(), start[b], start[b + 1], buckets.data() + G2_RAW_JAC * b);
  const uint32_t nchunk = (B + CH - 1) / CH;
  std::vector<uint32_t> partial((size_t)W * nchunk * G2_RAW_JAC), sums((size_t)W * G2_RAW_JAC);
  for (int w = 0; w < W; w++)
    for (uint32_t j = 0; j < nchunk; j++) {
      uint32_t lo = j * CH, hi = lo + CH < B ? lo + CH : B;
      if (lo == 0) lo = 1;
      prog_g2_chunk_weighted(cx, buckets.data() + (size_t)G2_RAW_JAC * w * B, lo, hi, partial.data() + G2_RAW_JAC * ((size_t)w * nchunk + j));
    }
  for (int w = 0; w < W; w++) prog_g2_jac_sum(cx, partial.data(), (size_t)w * nchunk, (size_t)(w + 1) * nchunk, sums.data() + G2_RAW_JAC * w);
  prog_g2_msm_final(cx, sums.data(), W, c, out48, out_inf);
  return err;
}

// Montgomery reduction of an arbitrary 26-word value (two's complement, |t| small enough): out = 13 stored words
int hs_redc(const uint32_t* t26, uint32_t* out13) {
  Acc t, t2, t3;
  for (int k = 0; k < NW; k++) { t.c[k] = t26[k]; t2.c[k] = t26[k]; t3.c[k] = t26[k]; }
#ifdef B381_TRACK_BOUNDS
  t.cb = t2.cb = t3.cb = 0; t.mag = t2.mag = t3.mag = 100.0;
#endif
  Fp r, r2, r3;
  acc_redc(r, t);
  acc_redc2(r2, t2, r3, t3);
  for (int k = 0; k < NL; k++) { out13[k] = (uint32_t)r.l[k]; if (r.l[k] != r2.l[k] || r.l[k] != r3.l[k]) return 1; }
  return 0;
}

// raw products: t = a * b (12- or 13-word operands) and t = a0 b0 + a1 b1 + a2 b2, 26 words out
static void set_fp_raw(Fp& r, const uint32_t* w) {
  for (int k = 0; k < NL; k++) r.l[k] = w[k];
  B381_SETRANGE(r, 0.0, 9.7);
}
int hs_mul_raw(const uint32_t* a13, const uint32_t* b13, int n12, uint32_t* out26) {
  Fp a, b;
  set_fp_raw(a, a13); set_fp_raw(b, b13);
  Acc t;
  B381_TB(t.mag = 0; t.cb = 0;)
  if (n12) acc_mul12(t, a, b); else acc_mul(t, a, b);
  for (int k = 0; k < NW; k++) out26[k] = t.c[k];
  return 0;
}
int hs_mul3_raw(const uint32_t* ops /* a0 b0 a1 b1 a2 b2, 13 words each */, int n12, uint32_t* out26) {
  Fp v[6];
  for (int i = 0; i < 6; i++) set_fp_raw(v[i], ops + 13 * i);
  Acc t;
  B381_TB(t.mag = 0; t.cb = 0;)
  if (n12) acc_mul3_12(t, v[0], v[1], v[2], v[3], v[4], v[5]); else acc_mul3(t, v[0], v[1], v[2], v[3], v[4], v[5]);
  for (int k = 0; k < NW; k++) out26[k] = t.c[k];
  return 0;
}

// weak reduction of a 13-word two's-complement value, |v| < 2^20 p
int hs_wreduce_raw(const uint32_t* in13, uint32_t* out13) {
  Fp a;
  for (int k = 0; k < NL; k++) a.l[k] = in13[k];
  B381_SETRANGE(a, -1000000.0, 1000000.0);
  fp_wreduce(a);
  for (int k = 0; k < NL; k++) out13[k] = (uint32_t)a.l[k];
  return 0;
}

int hs_tracking(void) {
#ifdef B381_TRACK_BOUNDS
  return 1;
#else
  return 0;
#endif
}

}  // extern "C"

extern "C" int hs_fp12_exp_by_x(const uint32_t* a, uint32_t* out) {
  Ctx cx = make_ctx();
  bool ok = f12_load_ext(cx, FE_F, a);
  f12_exp_by_x(cx, FE_Y1, FE_F, FE_ACC, FE_ACC2, FE_T);
  f12_store_ext(cx, out, FE_Y1);
  return ok ? 0 : 1;
}

// the plain (uncompressed, left-to-right) form: the fall-back of the compressed one
extern "C" int hs_fp12_exp_by_x_plain(const uint32_t* a, uint32_t* out) {
  Ctx cx = make_ctx();
  bool ok = f12_load_ext(cx, FE_F, a);
  f12_exp_by_x_gs(cx, FE_Y1, FE_F, FE_ACC, FE_ACC2, FE_T);
  f12_store_ext(cx, out, FE_Y1);
  return ok ? 0 : 1;
}

// n compressed (Karabina) squarings; only the coefficients +1, +2, +3, +5 of the output are meaningful
extern "C" int hs_fp12_compressed_squarings(const uint32_t* a, int n, uint32_t* out) {
  Ctx cx = make_ctx();
  bool ok = f12_load_ext(cx, FE_F, a);
  for (int i = 0; i < 6; i++) f2_set_small(slot(cx, FE_Y1 + i), 0);
  f12_csqr_run(cx, FE_Y1, FE_F, n, FE_ACC, FE_ACC2);
  f12_store_ext(cx, out, FE_Y1);
  return ok ? 0 : 1;
}
