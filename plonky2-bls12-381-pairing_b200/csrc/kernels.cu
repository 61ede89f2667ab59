// kernels.cu -- CUDA kernels (sm_100a) and the C ABI of libb381.so (include/b381.h).
//
// Execution model.  One pairing per thread.  The tower kernels are PERSISTENT: one CTA of
// B381_BLOCK threads per SM (grid = min(batches, #SM)), each CTA loops over batches of
// B381_BLOCK pairings.  Per-thread state is an arena of Fp2 slots (24 words): the hot 9 slots in shared
// memory (216 KB per CTA, word-interleaved so every LDS.128/STS.128 is conflict-free), 10 more in
// tensor memory, the cold slots in a per-CTA global scratch region that stays L2-resident because
// only #SM CTAs exist.  The arithmetic is fp32.cuh (13 x 32-bit words, IMAD.WIDE.U32.X carry chains);
// see DESIGN.md section 2.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b381.h"
#include "programs.cuh"
#include "helpers.cuh"

using namespace b381;

namespace {

constexpr int BLOCK = B381_BLOCK;
constexpr int NG_SLOTS = MAX_NSLOTS - NS;                               // cold slots per thread
constexpr size_t SMEM_BYTES = (size_t)NS * GPS * sizeof(u4) * BLOCK;    // 9 slots x 96 B x 256 threads = 216 KB
constexpr size_t GARENA_U4_PER_CTA = (size_t)NG_SLOTS * GPS * BLOCK;
constexpr int RAW_WORDS = 6 * 28;                                       // internal-format Fp12

#ifndef B381_LOCKSTEP
#define B381_LOCKSTEP 1
#endif
#ifndef B381_USE_TMEM
#define B381_USE_TMEM 1
#endif

__device__ __forceinline__ Ctx make_ctx(u4* garena, int lockstep, uint32_t tmem_base = 0, int use_tmem = 0) {
  extern __shared__ u4 smem[];
  Ctx cx;
  cx.sm = smem + threadIdx.x;
  cx.gm = garena + (size_t)blockIdx.x * GARENA_U4_PER_CTA + threadIdx.x;
  cx.sync = lockstep;
  const uint32_t warp = threadIdx.x >> 5;
  cx.tm = tmem_base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 256u;
  cx.nt = use_tmem ? NT_MAX : 0;
  return cx;
}

// the whole tensor memory of the SM (512 columns) for this CTA; called by all threads
__device__ __forceinline__ uint32_t tmem_alloc_all() {
  __shared__ uint32_t s_tmem_base;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  return s_tmem_base;
}

__device__ __forceinline__ void tmem_free_all(uint32_t base) {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}

#if B381_USE_TMEM && (B381_BLOCK == 256)
#define B381_TMEM_BEGIN() const uint32_t tmem_base_ = tmem_alloc_all()
#define B381_TMEM_CTX(garena) make_ctx(garena, B381_LOCKSTEP, tmem_base_, 1)
#define B381_TMEM_END() tmem_free_all(tmem_base_)
#else
#define B381_TMEM_BEGIN()
#define B381_TMEM_CTX(garena) make_ctx(garena, B381_LOCKSTEP)
#define B381_TMEM_END()
#endif

__device__ __forceinline__ void report(int e, int* err) {
  if (e) atomicOr(err, e);
}

// ---- tower kernels (persistent, one CTA per SM) ---------------------------------------------------
// Lock-step kernels: every thread of the CTA executes the same program (threads past the end of the
// batch recompute the last element and drop the result), so sync_point() barriers are legal.
__global__ void __launch_bounds__(BLOCK, 1)
k_miller(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out, size_t n, int mode, u4* garena, int* err, uint32_t* dump) {
  B381_TMEM_BEGIN();
  Ctx cx = B381_TMEM_CTX(garena);
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < n; base += (size_t)gridDim.x * BLOCK) {
    __syncthreads();
    size_t i = base + threadIdx.x;
    const bool active = i < n;
    if (!active) i = n - 1;
    int e = prog_miller(cx, g1 + 24 * i, g2 + 48 * i, inf ? inf[i] : 0, active ? out + 144 * i : dump + 144 * threadIdx.x, mode);
    if (active) report(e, err);
  }
  B381_TMEM_END();
}

__global__ void __launch_bounds__(BLOCK, 1)
k_final_exp(const uint32_t* in, uint32_t* out, size_t n, u4* garena, int* err, uint32_t* dump) {
  B381_TMEM_BEGIN();
  Ctx cx = B381_TMEM_CTX(garena);
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < n; base += (size_t)gridDim.x * BLOCK) {
    __syncthreads();
    size_t i = base + threadIdx.x;
    const bool active = i < n;
    if (!active) i = n - 1;
    int e = prog_final_exp(cx, in + 144 * i, active ? out + 144 * i : dump + 144 * threadIdx.x);
    if (active) report(e, err);
  }
  B381_TMEM_END();
}

__global__ void __launch_bounds__(BLOCK, 1)
k_pairing(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out, size_t n, int mode, u4* garena, int* err, uint32_t* dump) {
  B381_TMEM_BEGIN();
  Ctx cx = B381_TMEM_CTX(garena);
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < n; base += (size_t)gridDim.x * BLOCK) {
    __syncthreads();
    size_t i = base + threadIdx.x;
    const bool active = i < n;
    if (!active) i = n - 1;
    int e = prog_pairing(cx, g1 + 24 * i, g2 + 48 * i, inf ? inf[i] : 0, active ? out + 144 * i : dump + 144 * threadIdx.x, mode);
    if (active) report(e, err);
  }
  B381_TMEM_END();
}

// G2Prepared stage: 68 line-coefficient triples per Q (external format, 4896 words per point)
__global__ void __launch_bounds__(BLOCK, 1)
k_g2_prepare(const uint32_t* g2, uint32_t* coeffs, size_t n, int mode, u4* garena, int* err, uint32_t* dump_coeffs) {
  B381_TMEM_BEGIN();
  Ctx cx = B381_TMEM_CTX(garena);
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < n; base += (size_t)gridDim.x * BLOCK) {
    __syncthreads();
    size_t i = base + threadIdx.x;
    const bool active = i < n;
    if (!active) i = n - 1;
    int e = prog_g2_prepare(cx, g2 + 48 * i, active ? coeffs + (size_t)G2PREP_WORDS * i : dump_coeffs + (size_t)G2PREP_WORDS * threadIdx.x, mode);
    if (active) report(e, err);
  }
  B381_TMEM_END();
}

// Miller loop (optionally + final exponentiation) of (P, prepared Q)
__global__ void __launch_bounds__(BLOCK, 1)
k_miller_prepared(const uint32_t* g1, const uint32_t* coeffs, const uint8_t* inf, uint32_t* out, size_t n, int mode, int do_fe, u4* garena, int* err, uint32_t* dump) {
  B381_TMEM_BEGIN();
  Ctx cx = B381_TMEM_CTX(garena);
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < n; base += (size_t)gridDim.x * BLOCK) {
    __syncthreads();
    size_t i = base + threadIdx.x;
    const bool active = i < n;
    if (!active) i = n - 1;
    int e = prog_miller_prepared(cx, g1 + 24 * i, coeffs + (size_t)G2PREP_WORDS * i, inf ? inf[i] : 0, active ? out + 144 * i : dump + 144 * threadIdx.x, mode, do_fe);
    if (active) report(e, err);
  }
  B381_TMEM_END();
}

// every thread runs the Miller loops of TWO pairs per round with shared squarings, multiplies the
// result into a private accumulator and dumps it (internal format) to partial[global thread id]
__global__ void __launch_bounds__(BLOCK, 1)
k_multi_miller(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, size_t n, int mode, uint32_t* partial, int accumulate, u4* garena, int* err) {
  B381_TMEM_BEGIN();
  Ctx cx = B381_TMEM_CTX(garena);
  if (accumulate) f12_load_raw(cx, M2_ACC, partial + (size_t)RAW_WORDS * ((size_t)blockIdx.x * BLOCK + threadIdx.x));
  else f12_set_one(cx, M2_ACC);
  for (size_t base = (size_t)blockIdx.x * BLOCK * 2; base < n; base += (size_t)gridDim.x * BLOCK * 2) {
    __syncthreads();
    size_t i0 = base + 2 * (size_t)threadIdx.x, i1 = i0 + 1;
    const bool act0 = i0 < n, act1 = i1 < n;        // inactive slots contribute 1
    if (!act0) i0 = n - 1;
    if (!act1) i1 = n - 1;
    int e = miller2_to_slots(cx, g1 + 24 * i0, g2 + 48 * i0, act0 ? (inf ? inf[i0] : 0) : 3,
                             g1 + 24 * i1, g2 + 48 * i1, act1 ? (inf ? inf[i1] : 0) : 3, mode);
    if (act0) report(e, err);
    f12_mul(cx, M2_ACC, M2_ACC, ML_F, M2_SCRATCH, M2_SCRATCH + 6);   // scratch: slots that are dead once the loop has left f in ML_F
  }
  f12_store_raw(cx, partial + (size_t)RAW_WORDS * ((size_t)blockIdx.x * BLOCK + threadIdx.x), M2_ACC);
  B381_TMEM_END();
}

// out[j] = product of in[j*K .. min((j+1)K, n_in))   (internal format; trip counts differ -> no lock step)
__global__ void __launch_bounds__(BLOCK, 1)
k_f12_reduce_raw(const uint32_t* in, size_t n_in, uint32_t* out, size_t n_out, int K, u4* garena) {
  Ctx cx = make_ctx(garena, 0);
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < n_out; base += (size_t)gridDim.x * BLOCK) {
    size_t j = base + threadIdx.x;
    if (j < n_out) {
      size_t lo = j * (size_t)K, hi = lo + K < n_in ? lo + K : n_in;
      prog_f12_product_raw(cx, in + lo * RAW_WORDS, hi - lo, RAW_WORDS, out + j * RAW_WORDS);
    }
  }
}

// external -> internal dump (for b381_fp12_product) and internal -> external (optionally via final exp)
__global__ void __launch_bounds__(BLOCK, 1)
k_ext_to_raw(const uint32_t* in, uint32_t* out, size_t n, u4* garena, int* err) {
  Ctx cx = make_ctx(garena, 0);
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < n; base += (size_t)gridDim.x * BLOCK) {
    size_t i = base + threadIdx.x;
    if (i < n) {
      if (!f12_load_ext(cx, 0, in + 144 * i)) report(ERR_NOT_CANONICAL, err);
      f12_store_raw(cx, out + i * RAW_WORDS, 0);
    }
  }
}

__global__ void __launch_bounds__(BLOCK, 1)
k_raw_finish(const uint32_t* in_raw, uint32_t* out_ext, int do_final_exp, u4* garena, int* err) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Ctx cx = make_ctx(garena, 0);
  f12_load_raw(cx, FE_F, in_raw);
  if (do_final_exp) report(final_exp_slots(cx, FE_F), err);
  f12_store_ext(cx, out_ext, FE_F);
}

__global__ void __launch_bounds__(BLOCK, 1)
k_f12_mul(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, int wbasis, u4* garena, int* err, uint32_t* dump) {
  B381_TMEM_BEGIN();
  Ctx cx = B381_TMEM_CTX(garena);
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < n; base += (size_t)gridDim.x * BLOCK) {
    __syncthreads();
    size_t i = base + threadIdx.x;
    const bool active = i < n;
    if (!active) i = n - 1;
    uint32_t* o = active ? out + 144 * i : dump + 144 * threadIdx.x;
    int e = wbasis ? prog_wbasis_mul(cx, a + 144 * i, b + 144 * i, o) : prog_f12_mul(cx, a + 144 * i, b + 144 * i, o);
    if (active) report(e, err);
  }
  B381_TMEM_END();
}

// LITERAL loop: data-dependent branches (the reference's three line-function cases) -> no lock step
__global__ void __launch_bounds__(BLOCK, 1)
k_literal(const uint32_t* g1p, const uint32_t* g2p, uint32_t* out, size_t n, u4* garena, int* err) {
  Ctx cx = make_ctx(garena, 0);
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < n; base += (size_t)gridDim.x * BLOCK) {
    size_t i = base + threadIdx.x;
    if (i < n) report(prog_literal(cx, g1p + 36 * i, g2p + 72 * i, out + 144 * i), err);
  }
}

// subgroup membership [r] P == infinity: uniform ladder, data-dependent cases inside the group law -> no lock step
__global__ void __launch_bounds__(BLOCK, 1)
k_subgroup(const uint32_t* pts, const uint8_t* inf, int is_g2, uint32_t* out_words, size_t n, u4* garena, int* err) {
  Ctx cx = make_ctx(garena, 0);
  const int w = is_g2 ? 48 : 24;
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < n; base += (size_t)gridDim.x * BLOCK) {
    size_t i = base + threadIdx.x;
    if (i < n) {
      uint8_t r = 0;
      report(prog_subgroup_check(cx, pts + (size_t)w * i, is_g2, inf ? inf[i] : 0, &r), err);
      out_words[i] = r;
    }
  }
}

// [k_i] P_i with per-element 256-bit scalars: divergent control flow -> no lock step
__global__ void __launch_bounds__(BLOCK, 1)
k_scalar_mul(const uint32_t* pts, const uint32_t* scalars, const uint8_t* inf, int is_g2, uint32_t* out, size_t n, u4* garena, int* err) {
  Ctx cx = make_ctx(garena, 0);
  const int w = is_g2 ? 48 : 24;
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < n; base += (size_t)gridDim.x * BLOCK) {
    size_t i = base + threadIdx.x;
    if (i < n) {
      uint8_t f = 0;
      report(prog_scalar_mul(cx, pts + (size_t)w * i, is_g2, inf ? inf[i] : 0, scalars + 8 * i, out + (size_t)(w + 1) * i, &f), err);
      out[(size_t)(w + 1) * i + w] = f;
    }
  }
}

// out[j] = sum of in[j*K .. min((j+1)K, n_in)) in the packed point layout (trip counts differ -> no lock step)
__global__ void __launch_bounds__(BLOCK, 1)
k_point_sum(const uint32_t* in, size_t n_in, uint32_t* out, size_t n_out, int K, int is_g2, u4* garena, int* err) {
  Ctx cx = make_ctx(garena, 0);
  const size_t w1 = (is_g2 ? 48 : 24) + 1;
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < n_out; base += (size_t)gridDim.x * BLOCK) {
    size_t j = base + threadIdx.x;
    if (j < n_out) {
      size_t lo = j * (size_t)K, hi = lo + K < n_in ? lo + K : n_in;
      report(prog_point_sum(cx, in + lo * w1, hi - lo, is_g2, out + j * w1), err);
    }
  }
}

__global__ void k_fill_one_ext(uint32_t* out144) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const uint32_t one[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                              0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};  // 2^384 mod p
    for (int k = 0; k < 144; k++) out144[k] = k < 12 ? one[k] : 0;
  }
}

// ---- register-only element-wise kernels (no arena) -------------------------------------------------
__device__ __forceinline__ bool load_canon(Fp& x, const uint32_t* src) {
  uint32_t w[12];
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4 v0 = s4[0], v1 = s4[1], v2 = s4[2];
  w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
  w[8] = v2.x; w[9] = v2.y; w[10] = v2.z; w[11] = v2.w;
  fp_unpack32(x, w);
  return fp_below_p(x);
}

__device__ __forceinline__ void store_canon(uint32_t* dst, Fp& x) {   // x in (-p, 2p)
  fp_canon_small(x);
  uint32_t w[12];
  fp_pack32(w, x);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  d4[0] = make_uint4(w[0], w[1], w[2], w[3]);
  d4[1] = make_uint4(w[4], w[5], w[6], w[7]);
  d4[2] = make_uint4(w[8], w[9], w[10], w[11]);
}

__global__ void __launch_bounds__(256)
k_fp_mul(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, int* err) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fp x, y, r;
    bool ok = load_canon(x, a + 12 * i);
    ok &= load_canon(y, b + 12 * i);
    if (!ok) atomicOr(err, ERR_NOT_CANONICAL);
    Acc t;
    acc_zero(t);
    acc_mac(t, x, y);
    acc_redc384(r, t);
    store_canon(out + 12 * i, r);
  }
}

__global__ void __launch_bounds__(256)
k_fp_mul_chain(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, int k, int* err) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t wa[12], wb[12], wo[12];
    for (int j = 0; j < 12; j++) { wa[j] = a[12 * i + j]; wb[j] = b[12 * i + j]; }
    Fp x, y;
    bool ok = fp_from_ext(x, wa);
    ok &= fp_from_ext(y, wb);
    if (!ok) atomicOr(err, ERR_NOT_CANONICAL);
    for (int s = 0; s < k; s++) {
      Fp t;
      fp_mul(t, x, y);
      x = t;
    }
    fp_to_ext(wo, x);
    for (int j = 0; j < 12; j++) out[12 * i + j] = wo[j];
  }
}

__global__ void __launch_bounds__(128)
k_fp2_mul(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, int* err) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    Fp a0, a1, b0, b1, r0, r1;
    bool ok = load_canon(a0, a + 24 * i);
    ok &= load_canon(a1, a + 24 * i + 12);
    ok &= load_canon(b0, b + 24 * i);
    ok &= load_canon(b1, b + 24 * i + 12);
    if (!ok) atomicOr(err, ERR_NOT_CANONICAL);
    Acc A, B, T;
    acc_zero(A); acc_mac(A, a0, b0);
    acc_zero(B); acc_mac(B, a1, b1);
    acc_sub(T, A, B);
    acc_redc384(r0, T);
    Fp sa, sb;
    fp_add(sa, a0, a1);
    fp_add(sb, b0, b1);
    acc_add(T, A, B);
    acc_neg(T, T);
    acc_mac(T, sa, sb);
    acc_redc384(r1, T);
    store_canon(out + 24 * i, r0);
    store_canon(out + 24 * i + 12, r1);
  }
}

// ---- witness helpers (helpers.cuh): one element per thread, registers only ---------------------------
enum HelperOp { H_FP_INV = 0, H_FP_SQRT, H_FP_IS_SQUARE, H_FP_POW, H_FP2_INV, H_FP2_SQRT, H_FP2_IS_SQUARE };

__global__ void __launch_bounds__(128)
k_helper(int op, const uint32_t* a, const uint8_t* sgn, const uint32_t* e, int nwords, uint32_t* out, uint8_t* out8, size_t n, int* err) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int r = 0;
    const int s = sgn ? (sgn[i] & 1) : 0;
    switch (op) {
      case H_FP_INV: r = prog_fp_inv(a + 12 * i, out + 12 * i); break;
      case H_FP_SQRT: r = prog_fp_sqrt(a + 12 * i, s, out + 12 * i); break;
      case H_FP_IS_SQUARE: r = prog_fp_is_square(a + 12 * i, out8 + i); break;
      case H_FP_POW: r = prog_fp_pow(a + 12 * i, e, nwords, out + 12 * i); break;
      case H_FP2_INV: r = prog_fp2_inv(a + 24 * i, out + 24 * i); break;
      case H_FP2_SQRT: r = prog_fp2_sqrt(a + 24 * i, s, out + 24 * i); break;
      default: r = prog_fp2_is_square(a + 24 * i, out8 + i); break;
    }
    if (r) atomicOr(err, r);
  }
}

// wire formats (helpers.cuh): one element per thread
enum WireOp { W_FP_TO_DIGITS = 0, W_FP_FROM_DIGITS, W_FP12_TO_WITNESS, W_G1_DESER, W_G1_SER, W_G2_DESER, W_G2_SER };

__global__ void __launch_bounds__(128)
k_wire(int op, const uint32_t* in, const uint8_t* inf, int compressed, uint32_t* out, size_t n, int* err) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int r = 0;
    const uint8_t* inb = reinterpret_cast<const uint8_t*>(in);
    uint8_t* outb = reinterpret_cast<uint8_t*>(out);
    switch (op) {
      case W_FP_TO_DIGITS: r = prog_fp_to_digits(in + 12 * i, out + 12 * i); break;
      case W_FP_FROM_DIGITS: r = prog_fp_from_digits(in + 12 * i, out + 12 * i); break;
      case W_FP12_TO_WITNESS: r = prog_fp12_to_witness(in + 144 * i, out + 144 * i); break;
      case W_G1_DESER: {
        uint8_t f = 0;
        r = prog_g1_deserialize(inb + (compressed ? 48 : 96) * i, compressed, out + 25 * i, &f);
        out[25 * i + 24] = f;
      } break;
      case W_G1_SER: r = prog_g1_serialize(in + 24 * i, inf ? inf[i] : 0, compressed, outb + (compressed ? 48 : 96) * i); break;
      case W_G2_DESER: {
        uint8_t f = 0;
        r = prog_g2_deserialize(inb + (compressed ? 96 : 192) * i, compressed, out + 49 * i, &f);
        out[49 * i + 48] = f;
      } break;
      default: r = prog_g2_serialize(in + 48 * i, inf ? inf[i] : 0, compressed, outb + (compressed ? 96 : 192) * i); break;
    }
    if (r) atomicOr(err, r);
  }
}

// Fq12 (deg = 12) / Fq6 (deg = 6) inverse through the slot arena
__global__ void __launch_bounds__(BLOCK, 1)
k_tower_inv(const uint32_t* in, uint32_t* out, size_t n, int deg, u4* garena, int* err, uint32_t* dump) {
  Ctx cx = make_ctx(garena, B381_LOCKSTEP);
  const int w = deg == 12 ? 144 : 72;
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < n; base += (size_t)gridDim.x * BLOCK) {
    __syncthreads();
    size_t i = base + threadIdx.x;
    const bool active = i < n;
    if (!active) i = n - 1;
    uint32_t* o = active ? out + (size_t)w * i : dump + 144 * threadIdx.x;
    int e = deg == 12 ? prog_f12_inv(cx, in + (size_t)w * i, o) : prog_f6_inv(cx, in + (size_t)w * i, o);
    if (active) report(e, err);
  }
}

// integer-pipe roofline probe: what the multiplier pipe sustains on 32 x 32 -> 64-bit multiply-accumulates.
// Eight accumulator chains; the multiplicand of every multiply-accumulate is the low word of the
// neighbouring chain's accumulator, so it changes with every instruction and ptxas cannot hoist or
// strength-reduce anything: the loop body is 128 fused IMAD.WIDE.U32 and nothing else (checked in the
// SASS by tests/test_capi_symbols.py).  The first version multiplied loop-invariant operands; ptxas
// computed a*b once and turned the 128 "multiply-accumulates" into IADD3 chains, i.e. it measured the
// ALU pipe (64 adds/clk/SM) and overstated the multiplier peak twofold.  Measured: 32.0 IMAD.WIDE/clk/SM
// (a warp-wide IMAD.WIDE occupies the FMA-heavy pipe of its sub-partition for 4 cycles; plain 32-bit
// IMAD: 64/clk/SM, IMAD.HI: 25.6/clk/SM -- tools/imad_probe5.cu, profiles/imad_probe5_r01.jsonl).
// VARIANT 1 keeps the common multiplier in the HIGH half of a 64-bit register pair (an odd register):
// with all three source operands in even registers (variant 0, as ptxas happens to allocate it) the
// instruction loses a cycle to a register-bank conflict and the probe reads 25.6 instead of 32.0
// IMAD.WIDE/clk/SM.  The host takes the faster of the two.
template <int VARIANT>
__global__ void __launch_bounds__(1024)
k_imad_peak(uint32_t* out, const uint32_t* in, unsigned long long* cyc, int iters) {
  unsigned long long bb = ((unsigned long long)(in[32 + (threadIdx.x & 31)] | 1u) << 32) | in[33 + (threadIdx.x & 31)];
  uint32_t b;
  if (VARIANT == 0) b = (uint32_t)bb | 1u;
  else asm volatile("{ .reg .b32 lo_; mov.b64 {lo_, %0}, %1; }" : "=r"(b) : "l"(bb));
  unsigned long long d[8];
#pragma unroll
  for (int j = 0; j < 8; j++) d[j] = ((unsigned long long)in[64 + j] << 20) + threadIdx.x;
  __syncthreads();
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        unsigned long long p;
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"((uint32_t)d[(j + 1) & 7]), "r"(b));
        asm volatile("add.u64 %0, %0, %1;" : "+l"(d[j]) : "l"(p));
      }
    }
  }
  unsigned long long t1 = clock64();
  unsigned long long s = bb;
#pragma unroll
  for (int j = 0; j < 8; j++) s ^= d[j];
  if (s == 0x12345678ull) out[threadIdx.x] = (uint32_t)s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- host side ---------------------------------------------------------------------------------------
struct State {
  bool init = false;
  int device = -1;
  int sm_count = 0, cc_major = 0, cc_minor = 0;
  cudaStream_t stream[2] = {nullptr, nullptr};   // stream[0]: all kernels of the host-pointer API; stream[1]: spare lane
  cudaStream_t s_in = nullptr, s_out = nullptr;   // H2D / D2H copy streams of the pipelined host-pointer API
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  u4* garena[2] = {nullptr, nullptr};
  int* d_err = nullptr;
  uint32_t* d_in1[2] = {nullptr, nullptr};      // staging (device) per lane
  uint32_t* d_in2[2] = {nullptr, nullptr};
  uint8_t* d_inf[2] = {nullptr, nullptr};
  uint32_t* d_out[2] = {nullptr, nullptr};
  size_t cap_in1 = 0, cap_in2 = 0, cap_inf = 0, cap_out = 0;   // bytes per lane
  uint32_t* d_partial[2] = {nullptr, nullptr};  // raw partial products for multi_miller
  uint32_t* d_dump[2] = {nullptr, nullptr};     // sink for the outputs of padding threads (BLOCK x 144 words)
  const uint8_t* cur_inf = nullptr;             // identity flags of the chunk host_binary is launching
  uint32_t* d_dump_coeffs = nullptr;            // same for k_g2_prepare (BLOCK x 4896 words), allocated on first use
  unsigned long long launches = 0;
  std::string last_error;
  std::mutex mu;
};
State g;

int fail_cuda(cudaError_t e, const char* what) {
  g.last_error = std::string(what) + ": " + cudaGetErrorString(e);
  return B381_E_CUDA;
}
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail_cuda(e_, #x); } while (0)

int fail_arg(const char* what) {
  g.last_error = what;
  return B381_E_ARG;
}

int map_err(int bits) {
  if (bits & ERR_NOT_CANONICAL) { g.last_error = "input limbs not canonical (>= p)"; return B381_E_NOT_CANONICAL; }
  if (bits & ERR_ZERO_DIVISION) { g.last_error = "division by zero (final_exponentiation(0), f_den == 0 or inverse of zero)"; return B381_E_ZERO_DIVISION; }
  if (bits & 8) { g.last_error = "invalid point encoding (flag bits)"; return B381_E_BAD_ENCODING; }
  if (bits & 4) { g.last_error = "square root of a non-residue (or of zero with sgn0 = 1; point not on the curve)"; return B381_E_NOT_SQUARE; }
  return B381_OK;
}

int grid_for(size_t n) {
  size_t batches = (n + BLOCK - 1) / BLOCK;
  return (int)(batches < (size_t)g.sm_count ? batches : (size_t)g.sm_count);
}

template <typename K>
int set_smem(K kernel) {
  CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  return 0;
}


int grow(uint32_t** p0, uint32_t** p1, size_t* cap, size_t need) {
  if (need <= *cap) return 0;
  if (*p0) cudaFree(*p0);
  if (*p1) cudaFree(*p1);
  *p0 = *p1 = nullptr; *cap = 0;
  CU(cudaMalloc((void**)p0, need));
  CU(cudaMalloc((void**)p1, need));
  *cap = need;
  return 0;
}

int read_err(cudaStream_t s) {
  int h = 0;
  CU(cudaMemcpyAsync(&h, g.d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  if (h) CU(cudaMemsetAsync(g.d_err, 0, sizeof(int), s));
  return map_err(h);
}

constexpr size_t CHUNK = 1u << 17;    // elements per pipelined chunk of the element-wise host-pointer API
// pair kernels: chunks are whole rounds (4 x #SM x 256 pairs) so that no launch runs a partly filled round
size_t pair_chunk() { return 4 * (size_t)g.sm_count * BLOCK; }

// launch helpers (device pointers) ----------------------------------------------------------------------
// A launch is limited to a few rounds per CTA: CTAs of different SMs are only aligned at launch
// start, and per-round time creeps up by ~8 % once they have drifted apart (tools/gpu_exp3.py);
// back-to-back launches of one round each keep the whole chip on one instruction stream
// (34.2 ms/round against 34.8 at four rounds and 37.6 at 28 rounds per launch).
constexpr size_t MAX_ROUNDS_PER_LAUNCH = 1;
size_t pairs_per_launch() { return MAX_ROUNDS_PER_LAUNCH * (size_t)g.sm_count * BLOCK; }

int launch_miller(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out, size_t n, int mode, cudaStream_t s, int lane) {
  for (size_t off = 0; off < n; off += pairs_per_launch()) {
    size_t m = n - off < pairs_per_launch() ? n - off : pairs_per_launch();
    k_miller<<<grid_for(m), BLOCK, SMEM_BYTES, s>>>(g1 + 24 * off, g2 + 48 * off, inf ? inf + off : nullptr, out + 144 * off, m, mode, g.garena[lane], g.d_err, g.d_dump[lane]);
    g.launches++;
  }
  CU(cudaGetLastError());
  return 0;
}
int launch_final_exp(const uint32_t* in, uint32_t* out, size_t n, cudaStream_t s, int lane) {
  for (size_t off = 0; off < n; off += pairs_per_launch()) {
    size_t m = n - off < pairs_per_launch() ? n - off : pairs_per_launch();
    k_final_exp<<<grid_for(m), BLOCK, SMEM_BYTES, s>>>(in + 144 * off, out + 144 * off, m, g.garena[lane], g.d_err, g.d_dump[lane]);
    g.launches++;
  }
  CU(cudaGetLastError());
  return 0;
}
int launch_pairing(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out, size_t n, int mode, cudaStream_t s, int lane) {
  for (size_t off = 0; off < n; off += pairs_per_launch()) {
    size_t m = n - off < pairs_per_launch() ? n - off : pairs_per_launch();
    k_pairing<<<grid_for(m), BLOCK, SMEM_BYTES, s>>>(g1 + 24 * off, g2 + 48 * off, inf ? inf + off : nullptr, out + 144 * off, m, mode, g.garena[lane], g.d_err, g.d_dump[lane]);
    g.launches++;
  }
  CU(cudaGetLastError());
  return 0;
}

int launch_g2_prepare(const uint32_t* g2, uint32_t* coeffs, size_t n, int mode, cudaStream_t s, int lane) {
  if (!g.d_dump_coeffs) CU(cudaMalloc((void**)&g.d_dump_coeffs, (size_t)BLOCK * G2PREP_WORDS * sizeof(uint32_t)));
  for (size_t off = 0; off < n; off += pairs_per_launch()) {
    size_t m = n - off < pairs_per_launch() ? n - off : pairs_per_launch();
    k_g2_prepare<<<grid_for(m), BLOCK, SMEM_BYTES, s>>>(g2 + 48 * off, coeffs + (size_t)G2PREP_WORDS * off, m, mode, g.garena[lane], g.d_err, g.d_dump_coeffs);
    g.launches++;
  }
  CU(cudaGetLastError());
  return 0;
}
int launch_miller_prepared(const uint32_t* g1, const uint32_t* coeffs, const uint8_t* inf, uint32_t* out, size_t n, int mode, int do_fe, cudaStream_t s, int lane) {
  for (size_t off = 0; off < n; off += pairs_per_launch()) {
    size_t m = n - off < pairs_per_launch() ? n - off : pairs_per_launch();
    k_miller_prepared<<<grid_for(m), BLOCK, SMEM_BYTES, s>>>(g1 + 24 * off, coeffs + (size_t)G2PREP_WORDS * off, inf ? inf + off : nullptr, out + 144 * off, m, mode, do_fe, g.garena[lane], g.d_err, g.d_dump[lane]);
    g.launches++;
  }
  CU(cudaGetLastError());
  return 0;
}

// raw tree reduction of `cnt` partials living in buf (ping) using pong; returns pointer to the single result
int reduce_raw(uint32_t* ping, uint32_t* pong, size_t cnt, cudaStream_t s, int lane, uint32_t** result) {
  const int K = 16;
  while (cnt > 1) {
    size_t n_out = (cnt + K - 1) / K;
    k_f12_reduce_raw<<<grid_for(n_out), BLOCK, SMEM_BYTES, s>>>(ping, cnt, pong, n_out, K, g.garena[lane]);
    g.launches++;
    CU(cudaGetLastError());
    uint32_t* t = ping; ping = pong; pong = t;
    cnt = n_out;
  }
  *result = ping;
  return 0;
}

// accumulate the Miller values of `n` device-resident pairs into the per-thread partial products
// (every launch uses the full grid so that partial[] always has sm_count * BLOCK entries)
int launch_multi_accumulate(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, size_t n, int mode, bool first, cudaStream_t s, int lane) {
  uint32_t* ping = g.d_partial[lane];
  const size_t per = 2 * pairs_per_launch();        // two pairs per thread per round
  for (size_t off = 0; off < n; off += per) {
    size_t m = n - off < per ? n - off : per;
    k_multi_miller<<<g.sm_count, BLOCK, SMEM_BYTES, s>>>(g1 + 24 * off, g2 + 48 * off, inf ? inf + off : nullptr, m, mode, ping, !(first && off == 0), g.garena[lane], g.d_err);
    g.launches++;
  }
  CU(cudaGetLastError());
  return 0;
}

// tree-reduce the partial products, optional final exponentiation, result (external format) -> out144
int launch_multi_finish(uint32_t* out144, int do_fe, cudaStream_t s, int lane) {
  uint32_t* ping = g.d_partial[lane];
  uint32_t* pong = ping + (size_t)g.sm_count * BLOCK * RAW_WORDS;
  uint32_t* res = nullptr;
  int rc = reduce_raw(ping, pong, (size_t)g.sm_count * BLOCK, s, lane, &res);
  if (rc) return rc;
  k_raw_finish<<<1, BLOCK, SMEM_BYTES, s>>>(res, out144, do_fe, g.garena[lane], g.d_err);
  g.launches++;
  CU(cudaGetLastError());
  return 0;
}

// product of all Miller values of device-resident pairs -> out144 (device, external format), optional final exp
int launch_multi(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out144, size_t n, int mode, int do_fe, cudaStream_t s, int lane) {
  if (mode == B381_MODE_LITERAL) {                 // the reference's multi_miller_loop as written returns 1
    k_fill_one_ext<<<1, 32, 0, s>>>(out144);
    g.launches++;
    CU(cudaGetLastError());
    if (do_fe) { int rc = launch_final_exp(out144, out144, 1, s, lane); if (rc) return rc; }
    return 0;
  }
  int rc = launch_multi_accumulate(g1, g2, inf, n, mode, true, s, lane);
  if (rc) return rc;
  return launch_multi_finish(out144, do_fe, s, lane);
}

bool bad_mode(int mode) { return mode != B381_MODE_ARK && mode != B381_MODE_ZK && mode != B381_MODE_LITERAL; }

#define REQUIRE_INIT() do { if (!g.init) { g.last_error = "b381_init not called"; return B381_E_NOT_INIT; } } while (0)

// host-pointer pipelines --------------------------------------------------------------------------------
enum PairKind { PK_MILLER, PK_PAIRING };

int host_pairs(PairKind kind, const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out, size_t n, int mode) {
  const size_t CHUNK = pair_chunk();
  size_t c = n < CHUNK ? n : CHUNK;
  int rc;
  if ((rc = grow(&g.d_in1[0], &g.d_in1[1], &g.cap_in1, c * 24 * 4))) return rc;
  if ((rc = grow(&g.d_in2[0], &g.d_in2[1], &g.cap_in2, c * 48 * 4))) return rc;
  if ((rc = grow(&g.d_out[0], &g.d_out[1], &g.cap_out, c * 144 * 4))) return rc;
  if (inf) {
    if (c > g.cap_inf) {
      for (int l = 0; l < 2; l++) { if (g.d_inf[l]) cudaFree(g.d_inf[l]); g.d_inf[l] = nullptr; }
      g.cap_inf = 0;
      for (int l = 0; l < 2; l++) CU(cudaMalloc((void**)&g.d_inf[l], c));
      g.cap_inf = c;
    }
  }
  // three-stream pipeline over double-buffered staging: copies of chunk c+1 / c-1 overlap the kernels
  // of chunk c, and ALL kernels run on one stream (two kernels sharing the chip would break the
  // chip-wide lock step the launches rely on).
  cudaStream_t sk = g.stream[0];
  int lane = 0;
  size_t nchunks = 0;
  for (size_t off = 0; off < n; off += CHUNK, lane ^= 1, nchunks++) {
    size_t m = n - off < CHUNK ? n - off : CHUNK;
    if (nchunks >= 2) CU(cudaStreamWaitEvent(g.s_in, g.ev_k[lane], 0));        // staging inputs free again
    CU(cudaMemcpyAsync(g.d_in1[lane], g1 + off * 24, m * 24 * 4, cudaMemcpyHostToDevice, g.s_in));
    CU(cudaMemcpyAsync(g.d_in2[lane], g2 + off * 48, m * 48 * 4, cudaMemcpyHostToDevice, g.s_in));
    if (inf) CU(cudaMemcpyAsync(g.d_inf[lane], inf + off, m, cudaMemcpyHostToDevice, g.s_in));
    CU(cudaEventRecord(g.ev_in[lane], g.s_in));
    CU(cudaStreamWaitEvent(sk, g.ev_in[lane], 0));
    if (nchunks >= 2) CU(cudaStreamWaitEvent(sk, g.ev_out[lane], 0));          // staging output drained
    const uint8_t* dinf = inf ? g.d_inf[lane] : nullptr;
    if (kind == PK_MILLER) rc = launch_miller(g.d_in1[lane], g.d_in2[lane], dinf, g.d_out[lane], m, mode, sk, 0);
    else rc = launch_pairing(g.d_in1[lane], g.d_in2[lane], dinf, g.d_out[lane], m, mode, sk, 0);
    if (rc) return rc;
    CU(cudaEventRecord(g.ev_k[lane], sk));
    CU(cudaStreamWaitEvent(g.s_out, g.ev_k[lane], 0));
    CU(cudaMemcpyAsync(out + off * 144, g.d_out[lane], m * 144 * 4, cudaMemcpyDeviceToHost, g.s_out));
    CU(cudaEventRecord(g.ev_out[lane], g.s_out));
  }
  CU(cudaStreamSynchronize(g.s_out));
  return read_err(sk);
}

// generic element-wise host pipeline: two inputs of wi words, one output of wo words per element
int ensure_inf_staging(size_t c) {
  if (c > g.cap_inf) {
    for (int l = 0; l < 2; l++) { if (g.d_inf[l]) cudaFree(g.d_inf[l]); g.d_inf[l] = nullptr; }
    g.cap_inf = 0;
    for (int l = 0; l < 2; l++) CU(cudaMalloc((void**)&g.d_inf[l], c));
    g.cap_inf = c;
  }
  return 0;
}

template <typename L>
int host_binary(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, size_t wa, size_t wb, size_t wo, size_t chunk, L launch, const uint8_t* inf = nullptr) {
  size_t c = n < chunk ? n : chunk;
  int rc;
  if (inf && (rc = ensure_inf_staging(c))) return rc;
  if ((rc = grow(&g.d_in1[0], &g.d_in1[1], &g.cap_in1, c * wa * 4))) return rc;
  if (b && (rc = grow(&g.d_in2[0], &g.d_in2[1], &g.cap_in2, c * wb * 4))) return rc;
  if ((rc = grow(&g.d_out[0], &g.d_out[1], &g.cap_out, c * wo * 4))) return rc;
  cudaStream_t sk = g.stream[0];                    // same three-stream pipeline as host_pairs
  int lane = 0;
  size_t nchunks = 0;
  for (size_t off = 0; off < n; off += chunk, lane ^= 1, nchunks++) {
    size_t m = n - off < chunk ? n - off : chunk;
    if (nchunks >= 2) CU(cudaStreamWaitEvent(g.s_in, g.ev_k[lane], 0));
    CU(cudaMemcpyAsync(g.d_in1[lane], a + off * wa, m * wa * 4, cudaMemcpyHostToDevice, g.s_in));
    if (b) CU(cudaMemcpyAsync(g.d_in2[lane], b + off * wb, m * wb * 4, cudaMemcpyHostToDevice, g.s_in));
    if (inf) CU(cudaMemcpyAsync(g.d_inf[lane], inf + off, m, cudaMemcpyHostToDevice, g.s_in));
    CU(cudaEventRecord(g.ev_in[lane], g.s_in));
    CU(cudaStreamWaitEvent(sk, g.ev_in[lane], 0));
    if (nchunks >= 2) CU(cudaStreamWaitEvent(sk, g.ev_out[lane], 0));
    g.cur_inf = inf ? g.d_inf[lane] : nullptr;
    if ((rc = launch(g.d_in1[lane], g.d_in2[lane], g.d_out[lane], m, sk, 0))) return rc;
    CU(cudaEventRecord(g.ev_k[lane], sk));
    CU(cudaStreamWaitEvent(g.s_out, g.ev_k[lane], 0));
    CU(cudaMemcpyAsync(out + off * wo, g.d_out[lane], m * wo * 4, cudaMemcpyDeviceToHost, g.s_out));
    CU(cudaEventRecord(g.ev_out[lane], g.s_out));
  }
  CU(cudaStreamSynchronize(g.s_out));
  return read_err(sk);
}

int elem_grid(size_t n, int threads, int per_sm) {
  size_t blocks = (n + threads - 1) / threads;
  size_t cap = (size_t)g.sm_count * per_sm;
  return (int)(blocks < cap ? blocks : cap);
}

}  // namespace

// ======================================================================================================
// C ABI
// ======================================================================================================
extern "C" {

int b381_init(int device) {
  std::lock_guard<std::mutex> lk(g.mu);
  if (g.init) {
    if (g.device == device) return B381_OK;
    g.last_error = "already initialised on another device";
    return B381_E_ARG;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g.last_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libb381 has no CPU fallback)";
    return B381_E_CUDA;
  }
  if (device < 0 || device >= count) return fail_arg("device index out of range");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  g.device = device;
  g.sm_count = prop.multiProcessorCount;
  g.cc_major = prop.major;
  g.cc_minor = prop.minor;
  if ((size_t)prop.sharedMemPerBlockOptin < SMEM_BYTES) {
    g.last_error = "device lacks 216 KB opt-in shared memory per block (built for sm_100a)";
    return B381_E_CUDA;
  }
  CU(cudaStreamCreateWithFlags(&g.s_in, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&g.s_out, cudaStreamNonBlocking));
  for (int l = 0; l < 2; l++) {
    CU(cudaEventCreateWithFlags(&g.ev_in[l], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&g.ev_k[l], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&g.ev_out[l], cudaEventDisableTiming));
    CU(cudaStreamCreateWithFlags(&g.stream[l], cudaStreamNonBlocking));
    CU(cudaMalloc((void**)&g.garena[l], GARENA_U4_PER_CTA * sizeof(u4) * g.sm_count));
    CU(cudaMalloc((void**)&g.d_partial[l], 2 * (size_t)g.sm_count * BLOCK * RAW_WORDS * sizeof(uint32_t)));
    CU(cudaMalloc((void**)&g.d_dump[l], (size_t)BLOCK * 144 * sizeof(uint32_t)));
  }
  CU(cudaMalloc((void**)&g.d_err, sizeof(int)));
  CU(cudaMemset(g.d_err, 0, sizeof(int)));
  int rc;
  if ((rc = set_smem(k_miller)) || (rc = set_smem(k_final_exp)) || (rc = set_smem(k_pairing)) || (rc = set_smem(k_multi_miller)) ||
      (rc = set_smem(k_f12_reduce_raw)) || (rc = set_smem(k_ext_to_raw)) || (rc = set_smem(k_raw_finish)) || (rc = set_smem(k_f12_mul)) ||
      (rc = set_smem(k_literal)) || (rc = set_smem(k_g2_prepare)) || (rc = set_smem(k_miller_prepared)) || (rc = set_smem(k_tower_inv)) || (rc = set_smem(k_subgroup)) || (rc = set_smem(k_scalar_mul)) || (rc = set_smem(k_point_sum)))
    return rc;
  CU(cudaDeviceSynchronize());
  g.launches = 0;
  g.init = true;
  return B381_OK;
}

int b381_shutdown(void) {
  std::lock_guard<std::mutex> lk(g.mu);
  if (!g.init) return B381_OK;
  cudaDeviceSynchronize();
  for (int l = 0; l < 2; l++) {
    if (g.stream[l]) cudaStreamDestroy(g.stream[l]);
    if (g.ev_in[l]) { cudaEventDestroy(g.ev_in[l]); cudaEventDestroy(g.ev_k[l]); cudaEventDestroy(g.ev_out[l]); g.ev_in[l] = g.ev_k[l] = g.ev_out[l] = nullptr; }
    cudaFree(g.garena[l]); cudaFree(g.d_partial[l]); cudaFree(g.d_dump[l]); g.d_dump[l] = nullptr;
    cudaFree(g.d_in1[l]); cudaFree(g.d_in2[l]); cudaFree(g.d_inf[l]); cudaFree(g.d_out[l]);
    g.stream[l] = nullptr; g.garena[l] = nullptr; g.d_partial[l] = nullptr;
    g.d_in1[l] = g.d_in2[l] = g.d_out[l] = nullptr; g.d_inf[l] = nullptr;
  }
  cudaFree(g.d_err);
  g.d_err = nullptr;
  if (g.s_in) { cudaStreamDestroy(g.s_in); cudaStreamDestroy(g.s_out); g.s_in = g.s_out = nullptr; }
  g.cap_in1 = g.cap_in2 = g.cap_inf = g.cap_out = 0;
  g.init = false;
  return B381_OK;
}

const char* b381_last_error(void) { return g.last_error.c_str(); }

int b381_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* scratch_bytes) {
  REQUIRE_INIT();
  if (sm_count) *sm_count = g.sm_count;
  if (cc_major) *cc_major = g.cc_major;
  if (cc_minor) *cc_minor = g.cc_minor;
  if (scratch_bytes) *scratch_bytes = 2 * GARENA_U4_PER_CTA * sizeof(u4) * g.sm_count;
  return B381_OK;
}

unsigned long long b381_kernel_launches(void) { return g.launches; }

// ---- device-pointer API ---------------------------------------------------------------------------
int b381_miller_loop_dev(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out, size_t n, int mode, void* stream) {
  REQUIRE_INIT();
  if (!g1 || !g2 || !out || n == 0 || bad_mode(mode)) return fail_arg("b381_miller_loop_dev: bad argument");
  if (mode == B381_MODE_LITERAL) return fail_arg("LITERAL per-pair values: use b381_literal_optimized / b381_multi_miller_loop");
  std::lock_guard<std::mutex> lk(g.mu);
  return launch_miller(g1, g2, inf, out, n, mode, (cudaStream_t)stream, 0);
}

int b381_final_exp_dev(const uint32_t* f, uint32_t* out, size_t n, void* stream) {
  REQUIRE_INIT();
  if (!f || !out || n == 0) return fail_arg("b381_final_exp_dev: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return launch_final_exp(f, out, n, (cudaStream_t)stream, 0);
}

int b381_pairing_dev(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out, size_t n, int mode, void* stream) {
  REQUIRE_INIT();
  if (!g1 || !g2 || !out || n == 0 || bad_mode(mode) || mode == B381_MODE_LITERAL) return fail_arg("b381_pairing_dev: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return launch_pairing(g1, g2, inf, out, n, mode, (cudaStream_t)stream, 0);
}

int b381_multi_miller_loop_dev(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out144, size_t n, int mode, void* stream) {
  REQUIRE_INIT();
  if (!g1 || !g2 || !out144 || n == 0 || bad_mode(mode)) return fail_arg("b381_multi_miller_loop_dev: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return launch_multi(g1, g2, inf, out144, n, mode, 0, (cudaStream_t)stream, 0);
}

int b381_fp_mul_dev(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, void* stream) {
  REQUIRE_INIT();
  if (!a || !b || !out || n == 0) return fail_arg("b381_fp_mul_dev: bad argument");
  k_fp_mul<<<elem_grid(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(a, b, out, n, g.d_err);
  g.launches++;
  CU(cudaGetLastError());
  return 0;
}

int b381_fp_mul_chain_dev(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, int k, void* stream) {
  REQUIRE_INIT();
  if (!a || !b || !out || n == 0 || k < 0) return fail_arg("b381_fp_mul_chain_dev: bad argument");
  k_fp_mul_chain<<<elem_grid(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(a, b, out, n, k, g.d_err);
  g.launches++;
  CU(cudaGetLastError());
  return 0;
}

int b381_fp2_mul_dev(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, void* stream) {
  REQUIRE_INIT();
  if (!a || !b || !out || n == 0) return fail_arg("b381_fp2_mul_dev: bad argument");
  k_fp2_mul<<<elem_grid(n, 128, 8), 128, 0, (cudaStream_t)stream>>>(a, b, out, n, g.d_err);
  g.launches++;
  CU(cudaGetLastError());
  return 0;
}

int b381_fp12_mul_dev(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, void* stream) {
  REQUIRE_INIT();
  if (!a || !b || !out || n == 0) return fail_arg("b381_fp12_mul_dev: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  k_f12_mul<<<grid_for(n), BLOCK, SMEM_BYTES, (cudaStream_t)stream>>>(a, b, out, n, 0, g.garena[0], g.d_err, g.d_dump[0]);
  g.launches++;
  CU(cudaGetLastError());
  return 0;
}

int b381_check_dev(void* stream) {
  REQUIRE_INIT();
  return read_err((cudaStream_t)stream);
}

// ---- host-pointer API -----------------------------------------------------------------------------
int b381_miller_loop(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out, size_t n, int mode) {
  REQUIRE_INIT();
  if (!g1 || !g2 || !out || n == 0 || bad_mode(mode)) return fail_arg("b381_miller_loop: bad argument");
  if (mode == B381_MODE_LITERAL) return fail_arg("LITERAL per-pair values: use b381_literal_optimized / b381_multi_miller_loop");
  std::lock_guard<std::mutex> lk(g.mu);
  return host_pairs(PK_MILLER, g1, g2, inf, out, n, mode);
}

int b381_pairing(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out, size_t n, int mode) {
  REQUIRE_INIT();
  if (!g1 || !g2 || !out || n == 0 || bad_mode(mode) || mode == B381_MODE_LITERAL) return fail_arg("b381_pairing: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return host_pairs(PK_PAIRING, g1, g2, inf, out, n, mode);
}

int b381_final_exp(const uint32_t* f, uint32_t* out, size_t n) {
  REQUIRE_INIT();
  if (!f || !out || n == 0) return fail_arg("b381_final_exp: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return host_binary(f, nullptr, out, n, 144, 0, 144, pair_chunk(),
                     [](uint32_t* a, uint32_t*, uint32_t* o, size_t m, cudaStream_t s, int lane) { return launch_final_exp(a, o, m, s, lane); });
}

static int multi_host(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out144, size_t n, int mode, int do_fe) {
  // chunked like host_pairs: H2D of chunk c+1 overlaps the Miller kernels of chunk c; the per-thread
  // partial products stay on the device and are reduced once at the end.
  const size_t CHUNK = pair_chunk();
  size_t c = n < CHUNK ? n : CHUNK;
  int rc;
  if ((rc = grow(&g.d_in1[0], &g.d_in1[1], &g.cap_in1, c * 24 * 4))) return rc;
  if ((rc = grow(&g.d_in2[0], &g.d_in2[1], &g.cap_in2, c * 48 * 4))) return rc;
  if ((rc = grow(&g.d_out[0], &g.d_out[1], &g.cap_out, 144 * 4))) return rc;
  if (inf && c > g.cap_inf) {
    for (int l = 0; l < 2; l++) { if (g.d_inf[l]) cudaFree(g.d_inf[l]); g.d_inf[l] = nullptr; }
    g.cap_inf = 0;
    for (int l = 0; l < 2; l++) CU(cudaMalloc((void**)&g.d_inf[l], c));
    g.cap_inf = c;
  }
  cudaStream_t sk = g.stream[0];
  if (mode == B381_MODE_LITERAL) {
    if ((rc = launch_multi(nullptr, nullptr, nullptr, g.d_out[0], n, mode, do_fe, sk, 0))) return rc;
  } else {
    int lane = 0;
    size_t nchunks = 0;
    for (size_t off = 0; off < n; off += CHUNK, lane ^= 1, nchunks++) {
      size_t m = n - off < CHUNK ? n - off : CHUNK;
      if (nchunks >= 2) CU(cudaStreamWaitEvent(g.s_in, g.ev_k[lane], 0));
      CU(cudaMemcpyAsync(g.d_in1[lane], g1 + off * 24, m * 24 * 4, cudaMemcpyHostToDevice, g.s_in));
      CU(cudaMemcpyAsync(g.d_in2[lane], g2 + off * 48, m * 48 * 4, cudaMemcpyHostToDevice, g.s_in));
      if (inf) CU(cudaMemcpyAsync(g.d_inf[lane], inf + off, m, cudaMemcpyHostToDevice, g.s_in));
      CU(cudaEventRecord(g.ev_in[lane], g.s_in));
      CU(cudaStreamWaitEvent(sk, g.ev_in[lane], 0));
      if ((rc = launch_multi_accumulate(g.d_in1[lane], g.d_in2[lane], inf ? g.d_inf[lane] : nullptr, m, mode, off == 0, sk, 0))) return rc;
      CU(cudaEventRecord(g.ev_k[lane], sk));
    }
    if ((rc = launch_multi_finish(g.d_out[0], do_fe, sk, 0))) return rc;
  }
  CU(cudaMemcpyAsync(out144, g.d_out[0], 144 * 4, cudaMemcpyDeviceToHost, sk));
  return read_err(sk);
}

int b381_multi_miller_loop(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out144, size_t n, int mode) {
  REQUIRE_INIT();
  if (!g1 || !g2 || !out144 || n == 0 || bad_mode(mode)) return fail_arg("b381_multi_miller_loop: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return multi_host(g1, g2, inf, out144, n, mode, 0);
}

int b381_multi_pairing(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out144, size_t n, int mode) {
  REQUIRE_INIT();
  if (!g1 || !g2 || !out144 || n == 0 || bad_mode(mode)) return fail_arg("b381_multi_pairing: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return multi_host(g1, g2, inf, out144, n, mode, 1);
}

int b381_fp12_product(const uint32_t* in, uint32_t* out144, size_t n) {
  REQUIRE_INIT();
  if (!in || !out144 || n == 0) return fail_arg("b381_fp12_product: bad argument");
  if (n > (size_t)g.sm_count * BLOCK) return fail_arg("b381_fp12_product: n too large (max sm_count * 256)");
  std::lock_guard<std::mutex> lk(g.mu);
  int rc;
  if ((rc = grow(&g.d_in1[0], &g.d_in1[1], &g.cap_in1, n * 144 * 4))) return rc;
  if ((rc = grow(&g.d_out[0], &g.d_out[1], &g.cap_out, 144 * 4))) return rc;
  cudaStream_t s = g.stream[0];
  CU(cudaMemcpyAsync(g.d_in1[0], in, n * 144 * 4, cudaMemcpyHostToDevice, s));
  uint32_t* ping = g.d_partial[0];
  uint32_t* pong = ping + (size_t)g.sm_count * BLOCK * RAW_WORDS;
  k_ext_to_raw<<<grid_for(n), BLOCK, SMEM_BYTES, s>>>(g.d_in1[0], ping, n, g.garena[0], g.d_err);
  g.launches++;
  CU(cudaGetLastError());
  uint32_t* res = nullptr;
  if ((rc = reduce_raw(ping, pong, n, s, 0, &res))) return rc;
  k_raw_finish<<<1, BLOCK, SMEM_BYTES, s>>>(res, g.d_out[0], 0, g.garena[0], g.d_err);
  g.launches++;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out144, g.d_out[0], 144 * 4, cudaMemcpyDeviceToHost, s));
  return read_err(s);
}

int b381_literal_optimized(const uint32_t* g1proj, const uint32_t* g2proj, uint32_t* out, size_t n) {
  REQUIRE_INIT();
  if (!g1proj || !g2proj || !out || n == 0) return fail_arg("b381_literal_optimized: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return host_binary(g1proj, g2proj, out, n, 36, 72, 144, CHUNK,
                     [](uint32_t* a, uint32_t* b, uint32_t* o, size_t m, cudaStream_t s, int lane) {
                       k_literal<<<grid_for(m), BLOCK, SMEM_BYTES, s>>>(a, b, o, m, g.garena[lane], g.d_err);
                       g.launches++;
                       return cudaGetLastError() == cudaSuccess ? 0 : fail_cuda(cudaGetLastError(), "k_literal");
                     });
}

int b381_fp_mul(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  REQUIRE_INIT();
  if (!a || !b || !out || n == 0) return fail_arg("b381_fp_mul: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return host_binary(a, b, out, n, 12, 12, 12, (size_t)1 << 22,
                     [](uint32_t* x, uint32_t* y, uint32_t* o, size_t m, cudaStream_t s, int) { return b381_fp_mul_dev(x, y, o, m, s); });
}

int b381_fp_mul_chain(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, int k) {
  REQUIRE_INIT();
  if (!a || !b || !out || n == 0 || k < 0) return fail_arg("b381_fp_mul_chain: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return host_binary(a, b, out, n, 12, 12, 12, (size_t)1 << 22,
                     [k](uint32_t* x, uint32_t* y, uint32_t* o, size_t m, cudaStream_t s, int) { return b381_fp_mul_chain_dev(x, y, o, m, k, s); });
}

int b381_fp2_mul(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  REQUIRE_INIT();
  if (!a || !b || !out || n == 0) return fail_arg("b381_fp2_mul: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return host_binary(a, b, out, n, 24, 24, 24, (size_t)1 << 21,
                     [](uint32_t* x, uint32_t* y, uint32_t* o, size_t m, cudaStream_t s, int) { return b381_fp2_mul_dev(x, y, o, m, s); });
}

static int f12_mul_host(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, int wbasis) {
  return host_binary(a, b, out, n, 144, 144, 144, CHUNK,
                     [wbasis](uint32_t* x, uint32_t* y, uint32_t* o, size_t m, cudaStream_t s, int lane) {
                       k_f12_mul<<<grid_for(m), BLOCK, SMEM_BYTES, s>>>(x, y, o, m, wbasis, g.garena[lane], g.d_err, g.d_dump[lane]);
                       g.launches++;
                       return cudaGetLastError() == cudaSuccess ? 0 : fail_cuda(cudaGetLastError(), "k_f12_mul");
                     });
}

int b381_fp12_mul(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  REQUIRE_INIT();
  if (!a || !b || !out || n == 0) return fail_arg("b381_fp12_mul: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return f12_mul_host(a, b, out, n, 0);
}

int b381_fp12_mul_wbasis(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  REQUIRE_INIT();
  if (!a || !b || !out || n == 0) return fail_arg("b381_fp12_mul_wbasis: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return f12_mul_host(a, b, out, n, 1);
}

// ---- G2Prepared (cached line coefficients) ---------------------------------------------------------
int b381_g2_prepare(const uint32_t* g2, uint32_t* coeffs, size_t n, int mode) {
  REQUIRE_INIT();
  if (!g2 || !coeffs || n == 0 || bad_mode(mode) || mode == B381_MODE_LITERAL) return fail_arg("b381_g2_prepare: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return host_binary(g2, (const uint32_t*)nullptr, coeffs, n, 48, 0, G2PREP_WORDS, pairs_per_launch(),
                     [mode](uint32_t* x, uint32_t*, uint32_t* o, size_t m, cudaStream_t s, int lane) { return launch_g2_prepare(x, o, m, mode, s, lane); });
}

static int miller_prepared_host(const uint32_t* g1, const uint32_t* coeffs, const uint8_t* inf, uint32_t* out, size_t n, int mode, int do_fe) {
  return host_binary(g1, coeffs, out, n, 24, G2PREP_WORDS, 144, pairs_per_launch(),
                     [mode, do_fe](uint32_t* x, uint32_t* y, uint32_t* o, size_t m, cudaStream_t s, int lane) {
                       return launch_miller_prepared(x, y, g.cur_inf, o, m, mode, do_fe, s, lane);
                     }, inf);
}

int b381_miller_loop_prepared(const uint32_t* g1, const uint32_t* coeffs, const uint8_t* inf, uint32_t* out, size_t n, int mode) {
  REQUIRE_INIT();
  if (!g1 || !coeffs || !out || n == 0 || bad_mode(mode) || mode == B381_MODE_LITERAL) return fail_arg("b381_miller_loop_prepared: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return miller_prepared_host(g1, coeffs, inf, out, n, mode, 0);
}

int b381_pairing_prepared(const uint32_t* g1, const uint32_t* coeffs, const uint8_t* inf, uint32_t* out, size_t n, int mode) {
  REQUIRE_INIT();
  if (!g1 || !coeffs || !out || n == 0 || bad_mode(mode) || mode == B381_MODE_LITERAL) return fail_arg("b381_pairing_prepared: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return miller_prepared_host(g1, coeffs, inf, out, n, mode, 1);
}

int b381_g2_prepare_dev(const uint32_t* g2, uint32_t* coeffs, size_t n, int mode, void* stream) {
  REQUIRE_INIT();
  if (!g2 || !coeffs || n == 0 || bad_mode(mode) || mode == B381_MODE_LITERAL) return fail_arg("b381_g2_prepare_dev: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return launch_g2_prepare(g2, coeffs, n, mode, (cudaStream_t)stream, 0);
}

int b381_miller_loop_prepared_dev(const uint32_t* g1, const uint32_t* coeffs, const uint8_t* inf, uint32_t* out, size_t n, int mode, int final_exp, void* stream) {
  REQUIRE_INIT();
  if (!g1 || !coeffs || !out || n == 0 || bad_mode(mode) || mode == B381_MODE_LITERAL) return fail_arg("b381_miller_loop_prepared_dev: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return launch_miller_prepared(g1, coeffs, inf, out, n, mode, final_exp ? 1 : 0, (cudaStream_t)stream, 0);
}

// ---- witness helpers -------------------------------------------------------------------
// element-wise pipelines over host buffers; wi / wo = words per element in / out (wo = 0: byte output)
static int helper_host(int op, const uint32_t* a, const uint8_t* sgn, const uint32_t* e_words, int nwords, uint32_t* out, uint8_t* out8, size_t n, size_t w) {
  // exponent (shared by the batch) to the device once
  uint32_t* d_e = nullptr;
  if (e_words) {
    CU(cudaMalloc((void**)&d_e, (size_t)nwords * 4));
    CU(cudaMemcpy(d_e, e_words, (size_t)nwords * 4, cudaMemcpyHostToDevice));
  }
  const bool bytes_out = out8 != nullptr;
  std::vector<uint32_t> tmp;                         // byte results: each chunk writes its m bytes at the start of an m-word window
  uint32_t* hout = out;
  if (bytes_out) { tmp.resize(n); hout = tmp.data(); }
  int rc = host_binary(a, (const uint32_t*)nullptr, hout, n, w, 0, bytes_out ? 1 : w, CHUNK,
                       [op, d_e, nwords, bytes_out](uint32_t* x, uint32_t*, uint32_t* o, size_t m, cudaStream_t s, int) {
                         k_helper<<<elem_grid(m, 128, 8), 128, 0, s>>>(op, x, g.cur_inf, d_e, nwords, bytes_out ? nullptr : o, bytes_out ? reinterpret_cast<uint8_t*>(o) : nullptr, m, g.d_err);
                         g.launches++;
                         return cudaGetLastError() == cudaSuccess ? 0 : fail_cuda(cudaGetLastError(), "k_helper");
                       }, sgn);
  if (d_e) cudaFree(d_e);
  if (rc) return rc;
  if (bytes_out) {
    // each chunk wrote m bytes at the start of its m-word window of hout
    for (size_t off = 0; off < n; off += CHUNK) {
      size_t m = n - off < CHUNK ? n - off : CHUNK;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(tmp.data() + off);
      for (size_t i = 0; i < m; i++) out8[off + i] = src[i];
    }
  }
  return 0;
}

int b381_fp_inv(const uint32_t* a, uint32_t* out, size_t n) {
  REQUIRE_INIT();
  if (!a || !out || n == 0) return fail_arg("b381_fp_inv: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return helper_host(H_FP_INV, a, nullptr, nullptr, 0, out, nullptr, n, 12);
}
int b381_fp_sqrt(const uint32_t* a, const uint8_t* sgn, uint32_t* out, size_t n) {
  REQUIRE_INIT();
  if (!a || !out || n == 0) return fail_arg("b381_fp_sqrt: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return helper_host(H_FP_SQRT, a, sgn, nullptr, 0, out, nullptr, n, 12);
}
int b381_fp_is_square(const uint32_t* a, uint8_t* out, size_t n) {
  REQUIRE_INIT();
  if (!a || !out || n == 0) return fail_arg("b381_fp_is_square: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return helper_host(H_FP_IS_SQUARE, a, nullptr, nullptr, 0, nullptr, out, n, 12);
}
int b381_fp_pow(const uint32_t* a, const uint64_t* exp, size_t exp_limbs, uint32_t* out, size_t n) {
  REQUIRE_INIT();
  if (!a || !exp || exp_limbs == 0 || exp_limbs > 64 || !out || n == 0) return fail_arg("b381_fp_pow: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  std::vector<uint32_t> e(2 * exp_limbs);
  for (size_t i = 0; i < exp_limbs; i++) { e[2 * i] = (uint32_t)exp[i]; e[2 * i + 1] = (uint32_t)(exp[i] >> 32); }
  return helper_host(H_FP_POW, a, nullptr, e.data(), (int)(2 * exp_limbs), out, nullptr, n, 12);
}
int b381_fp2_inv(const uint32_t* a, uint32_t* out, size_t n) {
  REQUIRE_INIT();
  if (!a || !out || n == 0) return fail_arg("b381_fp2_inv: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return helper_host(H_FP2_INV, a, nullptr, nullptr, 0, out, nullptr, n, 24);
}
int b381_fp2_sqrt(const uint32_t* a, const uint8_t* sgn, uint32_t* out, size_t n) {
  REQUIRE_INIT();
  if (!a || !out || n == 0) return fail_arg("b381_fp2_sqrt: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return helper_host(H_FP2_SQRT, a, sgn, nullptr, 0, out, nullptr, n, 24);
}
int b381_fp2_is_square(const uint32_t* a, uint8_t* out, size_t n) {
  REQUIRE_INIT();
  if (!a || !out || n == 0) return fail_arg("b381_fp2_is_square: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return helper_host(H_FP2_IS_SQUARE, a, nullptr, nullptr, 0, nullptr, out, n, 24);
}
static int tower_inv_host(const uint32_t* a, uint32_t* out, size_t n, int deg) {
  const size_t w = deg == 12 ? 144 : 72;
  return host_binary(a, (const uint32_t*)nullptr, out, n, w, 0, w, CHUNK,
                     [deg](uint32_t* x, uint32_t*, uint32_t* o, size_t m, cudaStream_t s, int lane) {
                       k_tower_inv<<<grid_for(m), BLOCK, SMEM_BYTES, s>>>(x, o, m, deg, g.garena[lane], g.d_err, g.d_dump[lane]);
                       g.launches++;
                       return cudaGetLastError() == cudaSuccess ? 0 : fail_cuda(cudaGetLastError(), "k_tower_inv");
                     });
}
int b381_fp6_inv(const uint32_t* a, uint32_t* out, size_t n) {
  REQUIRE_INIT();
  if (!a || !out || n == 0) return fail_arg("b381_fp6_inv: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return tower_inv_host(a, out, n, 6);
}
int b381_fp12_inv(const uint32_t* a, uint32_t* out, size_t n) {
  REQUIRE_INIT();
  if (!a || !out || n == 0) return fail_arg("b381_fp12_inv: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return tower_inv_host(a, out, n, 12);
}

// ---- wire formats -------------------------------------------------------------------------
static int wire_host(int op, const uint32_t* in, const uint8_t* inf, int compressed, uint32_t* out, size_t n, size_t wi, size_t wo) {
  return host_binary(in, (const uint32_t*)nullptr, out, n, wi, 0, wo, CHUNK,
                     [op, compressed](uint32_t* x, uint32_t*, uint32_t* o, size_t m, cudaStream_t s, int) {
                       k_wire<<<elem_grid(m, 128, 8), 128, 0, s>>>(op, x, g.cur_inf, compressed, o, m, g.d_err);
                       g.launches++;
                       return cudaGetLastError() == cudaSuccess ? 0 : fail_cuda(cudaGetLastError(), "k_wire");
                     }, inf);
}
int b381_fp_to_u32_digits(const uint32_t* a, uint32_t* out, size_t n) {
  REQUIRE_INIT();
  if (!a || !out || n == 0) return fail_arg("b381_fp_to_u32_digits: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return wire_host(W_FP_TO_DIGITS, a, nullptr, 0, out, n, 12, 12);
}
int b381_fp_from_u32_digits(const uint32_t* digits, uint32_t* out, size_t n) {
  REQUIRE_INIT();
  if (!digits || !out || n == 0) return fail_arg("b381_fp_from_u32_digits: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return wire_host(W_FP_FROM_DIGITS, digits, nullptr, 0, out, n, 12, 12);
}
int b381_fp12_to_witness_limbs(const uint32_t* f, uint32_t* out, size_t n) {
  REQUIRE_INIT();
  if (!f || !out || n == 0) return fail_arg("b381_fp12_to_witness_limbs: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return wire_host(W_FP12_TO_WITNESS, f, nullptr, 0, out, n, 144, 144);
}
static int deser_host(int op, const uint8_t* in, int compressed, uint32_t* pts, uint8_t* inf, size_t n, size_t in_bytes, size_t pt_words) {
  std::vector<uint32_t> tmp(n * (pt_words + 1));
  int rc = wire_host(op, reinterpret_cast<const uint32_t*>(in), nullptr, compressed, tmp.data(), n, in_bytes / 4, pt_words + 1);
  for (size_t i = 0; i < n; i++) {                  // results are written even when an error is reported
    memcpy(pts + pt_words * i, tmp.data() + (pt_words + 1) * i, pt_words * 4);
    if (inf) inf[i] = (uint8_t)tmp[(pt_words + 1) * i + pt_words];
  }
  return rc;
}
int b381_g1_deserialize(const uint8_t* in, int compressed, uint32_t* g1, uint8_t* inf, size_t n) {
  REQUIRE_INIT();
  if (!in || !g1 || !inf || n == 0) return fail_arg("b381_g1_deserialize: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return deser_host(W_G1_DESER, in, compressed ? 1 : 0, g1, inf, n, compressed ? 48 : 96, 24);
}
int b381_g2_deserialize(const uint8_t* in, int compressed, uint32_t* g2, uint8_t* inf, size_t n) {
  REQUIRE_INIT();
  if (!in || !g2 || !inf || n == 0) return fail_arg("b381_g2_deserialize: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return deser_host(W_G2_DESER, in, compressed ? 1 : 0, g2, inf, n, compressed ? 96 : 192, 48);
}
int b381_g1_serialize(const uint32_t* g1, const uint8_t* inf, int compressed, uint8_t* out, size_t n) {
  REQUIRE_INIT();
  if (!g1 || !out || n == 0) return fail_arg("b381_g1_serialize: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return wire_host(W_G1_SER, g1, inf, compressed ? 1 : 0, reinterpret_cast<uint32_t*>(out), n, 24, compressed ? 12 : 24);
}
int b381_g2_serialize(const uint32_t* g2, const uint8_t* inf, int compressed, uint8_t* out, size_t n) {
  REQUIRE_INIT();
  if (!g2 || !out || n == 0) return fail_arg("b381_g2_serialize: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return wire_host(W_G2_SER, g2, inf, compressed ? 1 : 0, reinterpret_cast<uint32_t*>(out), n, 48, compressed ? 24 : 48);
}

static int subgroup_host(const uint32_t* pts, const uint8_t* inf, int is_g2, uint8_t* out, size_t n) {
  std::vector<uint32_t> tmp(n);
  int rc = host_binary(pts, (const uint32_t*)nullptr, tmp.data(), n, is_g2 ? 48 : 24, 0, 1, CHUNK,
                       [is_g2](uint32_t* x, uint32_t*, uint32_t* o, size_t m, cudaStream_t s, int lane) {
                         k_subgroup<<<grid_for(m), BLOCK, SMEM_BYTES, s>>>(x, g.cur_inf, is_g2, o, m, g.garena[lane], g.d_err);
                         g.launches++;
                         return cudaGetLastError() == cudaSuccess ? 0 : fail_cuda(cudaGetLastError(), "k_subgroup");
                       }, inf);
  for (size_t i = 0; i < n; i++) out[i] = (uint8_t)tmp[i];
  return rc;
}
int b381_g1_in_subgroup(const uint32_t* g1, const uint8_t* inf, uint8_t* out, size_t n) {
  REQUIRE_INIT();
  if (!g1 || !out || n == 0) return fail_arg("b381_g1_in_subgroup: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return subgroup_host(g1, inf, 0, out, n);
}
int b381_g2_in_subgroup(const uint32_t* g2, const uint8_t* inf, uint8_t* out, size_t n) {
  REQUIRE_INIT();
  if (!g2 || !out || n == 0) return fail_arg("b381_g2_in_subgroup: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return subgroup_host(g2, inf, 1, out, n);
}

static int scalar_mul_host(const uint32_t* pts, const uint32_t* scalars, const uint8_t* inf, int is_g2, uint32_t* out, uint8_t* out_inf, size_t n) {
  const size_t w = is_g2 ? 48 : 24;
  std::vector<uint32_t> tmp(n * (w + 1));
  int rc = host_binary(pts, scalars, tmp.data(), n, w, 8, w + 1, CHUNK,
                       [is_g2](uint32_t* x, uint32_t* y, uint32_t* o, size_t m, cudaStream_t s, int lane) {
                         k_scalar_mul<<<grid_for(m), BLOCK, SMEM_BYTES, s>>>(x, y, g.cur_inf, is_g2, o, m, g.garena[lane], g.d_err);
                         g.launches++;
                         return cudaGetLastError() == cudaSuccess ? 0 : fail_cuda(cudaGetLastError(), "k_scalar_mul");
                       }, inf);
  for (size_t i = 0; i < n; i++) {
    memcpy(out + w * i, tmp.data() + (w + 1) * i, w * 4);
    if (out_inf) out_inf[i] = (uint8_t)tmp[(w + 1) * i + w];
  }
  return rc;
}
int b381_g1_scalar_mul(const uint32_t* g1, const uint8_t* inf, const uint32_t* scalars, uint32_t* out, uint8_t* out_inf, size_t n) {
  REQUIRE_INIT();
  if (!g1 || !scalars || !out || !out_inf || n == 0) return fail_arg("b381_g1_scalar_mul: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return scalar_mul_host(g1, scalars, inf, 0, out, out_inf, n);
}
int b381_g2_scalar_mul(const uint32_t* g2, const uint8_t* inf, const uint32_t* scalars, uint32_t* out, uint8_t* out_inf, size_t n) {
  REQUIRE_INIT();
  if (!g2 || !scalars || !out || !out_inf || n == 0) return fail_arg("b381_g2_scalar_mul: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return scalar_mul_host(g2, scalars, inf, 1, out, out_inf, n);
}

// ---- point sums and the (naive) multi-scalar multiplication ---------------------------------------------
// tree-reduce n packed points on the device (16-ary), result -> host
static int point_tree(uint32_t* d_a, uint32_t* d_b, size_t n, int is_g2, uint32_t* out, uint8_t* out_inf) {
  const size_t w = is_g2 ? 48 : 24;
  cudaStream_t s = g.stream[0];
  const int K = 16;
  size_t cnt = n;
  do {                                              // at least one level: it also normalises a single point
    size_t n_out = (cnt + K - 1) / K;
    k_point_sum<<<grid_for(n_out), BLOCK, SMEM_BYTES, s>>>(d_a, cnt, d_b, n_out, K, is_g2, g.garena[0], g.d_err);
    g.launches++;
    CU(cudaGetLastError());
    uint32_t* t = d_a; d_a = d_b; d_b = t;
    cnt = n_out;
  } while (cnt > 1);
  std::vector<uint32_t> res(w + 1);
  CU(cudaMemcpyAsync(res.data(), d_a, (w + 1) * 4, cudaMemcpyDeviceToHost, s));
  int rc = read_err(s);
  memcpy(out, res.data(), w * 4);
  *out_inf = (uint8_t)res[w];
  return rc;
}

static int point_sum_host(const uint32_t* pts, const uint8_t* inf, const uint32_t* scalars, int is_g2, uint32_t* out, uint8_t* out_inf, size_t n) {
  const size_t w = is_g2 ? 48 : 24;
  cudaStream_t s = g.stream[0];
  int rc;
  // the library's persistent staging buffers (no allocation per call): d_out[0] / d_out[1] ping-pong the packed points
  if ((rc = grow(&g.d_out[0], &g.d_out[1], &g.cap_out, n * (w + 1) * 4))) return rc;
  uint32_t *d_a = g.d_out[0], *d_b = g.d_out[1];
  if (scalars) {                                    // MSM: [k_i] P_i on the device, straight into the packed layout
    if ((rc = grow(&g.d_in1[0], &g.d_in1[1], &g.cap_in1, n * w * 4))) return rc;
    if ((rc = grow(&g.d_in2[0], &g.d_in2[1], &g.cap_in2, n * 8 * 4))) return rc;
    if (inf && (rc = ensure_inf_staging(n))) return rc;
    CU(cudaMemcpyAsync(g.d_in1[0], pts, n * w * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(g.d_in2[0], scalars, n * 8 * 4, cudaMemcpyHostToDevice, s));
    if (inf) CU(cudaMemcpyAsync(g.d_inf[0], inf, n, cudaMemcpyHostToDevice, s));
    k_scalar_mul<<<grid_for(n), BLOCK, SMEM_BYTES, s>>>(g.d_in1[0], g.d_in2[0], inf ? g.d_inf[0] : nullptr, is_g2, d_a, n, g.garena[0], g.d_err);
    g.launches++;
    CU(cudaGetLastError());
  } else {
    std::vector<uint32_t> packed(n * (w + 1));
    for (size_t i = 0; i < n; i++) {
      memcpy(packed.data() + (w + 1) * i, pts + w * i, w * 4);
      packed[(w + 1) * i + w] = inf ? (inf[i] & 1) : 0;
    }
    CU(cudaMemcpyAsync(d_a, packed.data(), n * (w + 1) * 4, cudaMemcpyHostToDevice, s));
    CU(cudaStreamSynchronize(s));
  }
  return point_tree(d_a, d_b, n, is_g2, out, out_inf);
}

int b381_g1_sum(const uint32_t* g1, const uint8_t* inf, uint32_t* out, uint8_t* out_inf, size_t n) {
  REQUIRE_INIT();
  if (!g1 || !out || !out_inf || n == 0) return fail_arg("b381_g1_sum: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return point_sum_host(g1, inf, nullptr, 0, out, out_inf, n);
}
int b381_g2_sum(const uint32_t* g2, const uint8_t* inf, uint32_t* out, uint8_t* out_inf, size_t n) {
  REQUIRE_INIT();
  if (!g2 || !out || !out_inf || n == 0) return fail_arg("b381_g2_sum: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return point_sum_host(g2, inf, nullptr, 1, out, out_inf, n);
}
int b381_g1_msm(const uint32_t* g1, const uint8_t* inf, const uint32_t* scalars, uint32_t* out, uint8_t* out_inf, size_t n) {
  REQUIRE_INIT();
  if (!g1 || !scalars || !out || !out_inf || n == 0) return fail_arg("b381_g1_msm: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return point_sum_host(g1, inf, scalars, 0, out, out_inf, n);
}
int b381_g2_msm(const uint32_t* g2, const uint8_t* inf, const uint32_t* scalars, uint32_t* out, uint8_t* out_inf, size_t n) {
  REQUIRE_INIT();
  if (!g2 || !scalars || !out || !out_inf || n == 0) return fail_arg("b381_g2_msm: bad argument");
  std::lock_guard<std::mutex> lk(g.mu);
  return point_sum_host(g2, inf, scalars, 1, out, out_inf, n);
}

int b381_imad_peak(double* imad_wide_ginst_per_s, double* sm_mhz) {
  REQUIRE_INIT();
  std::lock_guard<std::mutex> lk(g.mu);
  const int threads = 1024, iters = 4096;
  uint32_t *d_in = nullptr, *d_out = nullptr;
  unsigned long long* d_cyc = nullptr;
  CU(cudaMalloc((void**)&d_in, 4096 * 4));
  CU(cudaMalloc((void**)&d_out, 4096 * 4));
  CU(cudaMalloc((void**)&d_cyc, 1024 * 8));
  std::vector<uint32_t> h(4096);
  for (int i = 0; i < 4096; i++) h[i] = (0x9e3779b9u * (uint32_t)(i + 1)) | 1u;
  CU(cudaMemcpy(d_in, h.data(), 4096 * 4, cudaMemcpyHostToDevice));
  cudaStream_t s = g.stream[0];
  double best_rate = 0, best_mhz = 0;
  for (int variant = 0; variant < 2; variant++) {
    if (variant == 0) k_imad_peak<0><<<g.sm_count, threads, 0, s>>>(d_out, d_in, d_cyc, 256);
    else k_imad_peak<1><<<g.sm_count, threads, 0, s>>>(d_out, d_in, d_cyc, 256);
    g.launches++;
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0, s));
    if (variant == 0) k_imad_peak<0><<<g.sm_count, threads, 0, s>>>(d_out, d_in, d_cyc, iters);
    else k_imad_peak<1><<<g.sm_count, threads, 0, s>>>(d_out, d_in, d_cyc, iters);
    g.launches++;
    CU(cudaEventRecord(e1, s));
    CU(cudaStreamSynchronize(s));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<unsigned long long> hc(g.sm_count);
    CU(cudaMemcpy(hc.data(), d_cyc, g.sm_count * 8, cudaMemcpyDeviceToHost));
    double cavg = 0;
    for (int i = 0; i < g.sm_count; i++) cavg += (double)hc[i];
    cavg /= g.sm_count;
    const double inst = 16.0 * 8.0 * iters * threads * (double)g.sm_count;
    const double rate = inst / (ms * 1e-3) / 1e9;
    if (rate > best_rate) { best_rate = rate; best_mhz = cavg / (ms * 1e-3) / 1e6; }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
  }
  if (imad_wide_ginst_per_s) *imad_wide_ginst_per_s = best_rate;
  if (sm_mhz) *sm_mhz = best_mhz;
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_cyc);
  return B381_OK;
}

}  // extern "C"
