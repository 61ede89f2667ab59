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
#include "g1_launch.h"

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
// They run with 256 threads per CTA (two warps per SM sub-partition), or with 128 for the tail of a batch
// (host_api.inc pair_cfg): the slot arena is indexed by threadIdx.x with a fixed stride of BLOCK, the barriers are
// per group of 128 threads, and the tensor-memory columns of the second warp group simply stay unused.
__global__ void __launch_bounds__(BLOCK, 1)
k_miller(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out, size_t n, int mode, u4* garena, int* err, uint32_t* dump) {
  B381_TMEM_BEGIN();
  Ctx cx = B381_TMEM_CTX(garena);
  for (size_t base = (size_t)blockIdx.x * blockDim.x; base < n; base += (size_t)gridDim.x * blockDim.x) {
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
  for (size_t base = (size_t)blockIdx.x * blockDim.x; base < n; base += (size_t)gridDim.x * blockDim.x) {
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

  for (size_t base = (size_t)blockIdx.x * blockDim.x; base < n; base += (size_t)gridDim.x * blockDim.x) {
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
  for (size_t base = (size_t)blockIdx.x * blockDim.x; base < n; base += (size_t)gridDim.x * blockDim.x) {
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
  for (size_t base = (size_t)blockIdx.x * blockDim.x; base < n; base += (size_t)gridDim.x * blockDim.x) {
    __syncthreads();
    size_t i = base + threadIdx.x;
    const bool active = i < n;
    if (!active) i = n - 1;
    int e = prog_miller_prepared(cx, g1 + 24 * i, coeffs + (size_t)G2PREP_WORDS * i, inf ? inf[i] : 0, active ? out + 144 * i : dump + 144 * threadIdx.x, mode, do_fe);
    if (active) report(e, err);
  }
  B381_TMEM_END();
}

// G2Prepared in the packed layout (programs.cuh): `packed` is this launch's first tile; the launch offset is a multiple
// of BLOCK, the CTA base a multiple of 128 (tail launches use CTAs of 128 threads), so point i of the launch sits at
// g2pack_index(i).  Threads past the end of the batch work on the last point and write to a dump tile.
__global__ void __launch_bounds__(BLOCK, 1)
k_g2_prepare_packed(const uint32_t* g2, u4* packed, size_t n, int mode, u4* garena, int* err, u4* dump_tile) {
  B381_TMEM_BEGIN();
  Ctx cx = B381_TMEM_CTX(garena);
  for (size_t base = (size_t)blockIdx.x * blockDim.x; base < n; base += (size_t)gridDim.x * blockDim.x) {
    __syncthreads();
    size_t i = base + threadIdx.x;
    const bool active = i < n;
    u4* dst = active ? packed + g2pack_index(i) : dump_tile + threadIdx.x;
    if (!active) i = n - 1;
    int e = prog_g2_prepare_packed(cx, g2 + 48 * i, dst, mode);
    if (active) report(e, err);
  }
  B381_TMEM_END();
}

// Miller loop (optionally + final exponentiation) of (P, packed prepared Q)
__global__ void __launch_bounds__(BLOCK, 1)
k_miller_packed(const uint32_t* g1, const u4* packed, const uint8_t* inf, uint32_t* out, size_t n, int mode, int do_fe, int one_q, u4* garena, int* err, uint32_t* dump) {
  B381_TMEM_BEGIN();
  Ctx cx = B381_TMEM_CTX(garena);
  for (size_t base = (size_t)blockIdx.x * blockDim.x; base < n; base += (size_t)gridDim.x * blockDim.x) {
    __syncthreads();
    size_t i = base + threadIdx.x;
    const bool active = i < n;
    if (!active) i = n - 1;
    // one_q: every pair reads point 0 of the buffer (one cached Q against many P): warp-uniform addresses, broadcast loads
    int e = prog_miller_packed(cx, g1 + 24 * i, one_q ? packed : packed + g2pack_index(i), inf ? inf[i] : 0, active ? out + 144 * i : dump + 144 * threadIdx.x, mode, do_fe);
    if (active) report(e, err);
  }
  B381_TMEM_END();
}

// every thread runs the Miller loops of MK = FOUR pairs per round with shared squarings, multiplies the
// result into a private accumulator and dumps it (internal format) to partial[global thread id]
__global__ void __launch_bounds__(BLOCK, 1)
k_multi_miller(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, size_t n, int mode, uint32_t* partial, int accumulate, u4* garena, int* err) {
  B381_TMEM_BEGIN();
  Ctx cx = B381_TMEM_CTX(garena);
  if (accumulate) f12_load_raw(cx, M2_ACC, partial + (size_t)RAW_WORDS * ((size_t)blockIdx.x * BLOCK + threadIdx.x));
  else f12_set_one(cx, M2_ACC);
  for (size_t base = (size_t)blockIdx.x * BLOCK * MK; base < n; base += (size_t)gridDim.x * BLOCK * MK) {
    __syncthreads();
    const uint32_t* pg1[MK];
    const uint32_t* pg2[MK];
    int pinf[MK];
    bool any = false;
    for (int j = 0; j < MK; j++) {
      size_t i = base + MK * (size_t)threadIdx.x + j;
      const bool act = i < n;                        // inactive slots contribute 1
      if (!act) i = n - 1;
      pg1[j] = g1 + 24 * i;
      pg2[j] = g2 + 48 * i;
      pinf[j] = act ? (inf ? inf[i] : 0) : 3;
      any = any || act;
    }
    int e = miller_multi_to_slots(cx, pg1, pg2, pinf, mode);
    if (any) report(e, err);
    f12_mul(cx, M2_ACC, M2_ACC, ML_F, M2_SCRATCH, M2_SCRATCH + 6);   // scratch: slots that are dead once the loop has left f in ML_F
  }
  f12_store_raw(cx, partial + (size_t)RAW_WORDS * ((size_t)blockIdx.x * BLOCK + threadIdx.x), M2_ACC);
  B381_TMEM_END();
}

// every thread runs the Miller loops of FOUR pairs per round against packed prepared Q's with shared squarings and
// multiplies the result into its private accumulator (the partial products are reduced like k_multi_miller's).
// A round covers PK_K consecutive tiles: thread t takes point t of each, so the line reads stay coalesced.
__global__ void __launch_bounds__(BLOCK, 1)
k_multi_miller_packed(const uint32_t* g1, const u4* packed, const uint8_t* inf, size_t n, uint32_t* partial, int accumulate, u4* garena, int* err, const u4* one_tile) {
  B381_TMEM_BEGIN();
  Ctx cx = B381_TMEM_CTX(garena);
  if (accumulate) f12_load_raw(cx, M2_ACC, partial + (size_t)RAW_WORDS * ((size_t)blockIdx.x * BLOCK + threadIdx.x));
  else f12_set_one(cx, M2_ACC);
  for (size_t base = (size_t)blockIdx.x * BLOCK * PK_K; base < n; base += (size_t)gridDim.x * BLOCK * PK_K) {
    __syncthreads();
    const uint32_t* pg1[PK_K];
    const u4* ppk[PK_K];
    int pinf[PK_K];
    bool any = false;
    for (int j = 0; j < PK_K; j++) {
      size_t i = base + (size_t)j * BLOCK + threadIdx.x;
      const bool act = i < n;
      if (!act) i = n - 1;
      pg1[j] = g1 + 24 * i;
      ppk[j] = packed + g2pack_index(i);
      pinf[j] = act ? (inf ? inf[i] : 0) : 3;
      any = any || act;
    }
    int e = miller_pk_to_slots(cx, pg1, ppk, pinf, one_tile + threadIdx.x);
    if (any) report(e, err);
    f12_mul(cx, M2_ACC, M2_ACC, ML_F, M2_SCRATCH, M2_SCRATCH + 6);
  }
  f12_store_raw(cx, partial + (size_t)RAW_WORDS * ((size_t)blockIdx.x * BLOCK + threadIdx.x), M2_ACC);
  B381_TMEM_END();
}

__global__ void __launch_bounds__(BLOCK, 1) k_fill_one_line(u4* tile) { fill_one_line(tile + threadIdx.x); }

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
  for (size_t base = (size_t)blockIdx.x * blockDim.x; base < n; base += (size_t)gridDim.x * blockDim.x) {
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

// ---- group kernels.  Points and identity flags are separate arrays (points: 24 / 48 words, flags: one byte).
// G2 runs on the slot arena (divergent special cases inside the group law -> no lock step); G1 runs in registers
// over the base field (g1.cuh) and is compiled as its own translation unit (g1_kernels.cu, launchers in g1_launch.h).

__global__ void __launch_bounds__(BLOCK, 1)
k_g2_point_op(int op, const uint32_t* pts, const uint8_t* inf, const uint32_t* scalars, uint32_t* out, uint8_t* out8, size_t n, u4* garena, int* err) {
  Ctx cx = make_ctx(garena, 0);
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < n; base += (size_t)gridDim.x * BLOCK) {
    size_t i = base + threadIdx.x;
    if (i < n) {
      uint8_t f = 0;
      const int in_f = inf ? inf[i] : 0;
      if (op == PO_SUBGROUP) report(prog_g2_in_subgroup(cx, pts + 48 * i, in_f, &f), err);
      else if (op == PO_CLEAR_COFACTOR) report(prog_g2_clear_cofactor(cx, pts + 48 * i, in_f, out + 48 * i, &f), err);
      else report(prog_scalar_mul(cx, pts + 48 * i, 1, in_f, scalars + 8 * i, out + 48 * i, &f), err);
      out8[i] = f;
    }
  }
}

// G2: out[j] = sum of the points j*K .. min((j+1)K, n_in) (trip counts differ -> no lock step)
__global__ void __launch_bounds__(BLOCK, 1)
k_g2_point_sum(const uint32_t* in, const uint8_t* in_inf, size_t n_in, uint32_t* out, uint8_t* out_inf, size_t n_out, int K, u4* garena, int* err) {
  Ctx cx = make_ctx(garena, 0);
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < n_out; base += (size_t)gridDim.x * BLOCK) {
    size_t j = base + threadIdx.x;
    if (j < n_out) {
      size_t lo = j * (size_t)K, hi = lo + K < n_in ? lo + K : n_in;
      report(prog_g2_point_sum(cx, in + lo * 48, in_inf ? in_inf + lo : nullptr, hi - lo, out + j * 48, out_inf + j), err);
    }
  }
}

// ---- G2 bucket method (programs.cuh): per-thread trip counts differ -> no lock step, no tensor memory ----------
__global__ void __launch_bounds__(BLOCK, 1)
k_g2_msm_bucket_sums(const uint32_t* pts, const uint32_t* idx, const unsigned int* start, const uint32_t* order, size_t m, uint32_t* buckets, u4* garena, int* err) {
  Ctx cx = make_ctx(garena, 0);
  for (size_t t = (size_t)blockIdx.x * BLOCK + threadIdx.x; t < m; t += (size_t)gridDim.x * BLOCK) {
    const size_t b = order[t];                      // buckets by decreasing size: a warp's 32 buckets are equally full
    report(prog_g2_bucket_sum(cx, pts, idx, start[b], start[b + 1], buckets + (size_t)G2_RAW_JAC * b), err);
  }
}
__global__ void __launch_bounds__(BLOCK, 1)
k_g2_msm_chunks(const uint32_t* buckets, int W, int c, int CH, uint32_t* partial, u4* garena) {
  Ctx cx = make_ctx(garena, 0);
  const uint32_t B = 1u << c, nchunk = (B + CH - 1) / CH;
  const size_t total = (size_t)W * nchunk;
  for (size_t t = (size_t)blockIdx.x * BLOCK + threadIdx.x; t < total; t += (size_t)gridDim.x * BLOCK) {
    const uint32_t w = (uint32_t)(t / nchunk), j = (uint32_t)(t % nchunk);
    uint32_t lo = j * CH, hi = lo + CH < B ? lo + CH : B;
    if (lo == 0) lo = 1;                              // digit 0 contributes nothing
    prog_g2_chunk_weighted(cx, buckets + (size_t)G2_RAW_JAC * w * B, lo, hi, partial + (size_t)G2_RAW_JAC * t);
  }
}
__global__ void __launch_bounds__(BLOCK, 1)
k_g2_jac_sums(const uint32_t* in, size_t n_in, size_t per, uint32_t* sums, size_t n_out, u4* garena) {
  Ctx cx = make_ctx(garena, 0);
  for (size_t j = (size_t)blockIdx.x * BLOCK + threadIdx.x; j < n_out; j += (size_t)gridDim.x * BLOCK) {
    const size_t lo = j * per, hi = lo + per < n_in ? lo + per : n_in;
    prog_g2_jac_sum(cx, in, lo, hi, sums + (size_t)G2_RAW_JAC * j);
  }
}
__global__ void __launch_bounds__(BLOCK, 1)
k_g2_msm_final(const uint32_t* sums, int W, int c, uint32_t* out48, uint8_t* out_inf, u4* garena) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  Ctx cx = make_ctx(garena, 0);
  prog_g2_msm_final(cx, sums, W, c, out48, out_inf);
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

// Streaming element-wise kernels (BASELINE config #2): a warp handles 32 consecutive elements.  The 32 x W words of
// a tile are contiguous in HBM: the warp moves them with fully coalesced 128-bit loads / stores (lane l takes the
// uint4 groups l, l + 32, ...: 512 contiguous bytes per instruction) through a shared-memory tile; each thread then
// picks up ITS element with 128-bit shared-memory loads.  Row strides of 12 and 28 words make those conflict-free
// (a quarter-warp's eight 16-byte accesses fall into eight different groups of four banks).
template <int W, int RS>
__device__ __forceinline__ void tile_load(uint32_t* tile, const uint32_t* gsrc, int valid, int lane) {
  const uint4* g4 = reinterpret_cast<const uint4*>(gsrc);
#pragma unroll
  for (int k = 0; k < W / 4; k++) {
    const int idx = lane + 32 * k, e = (idx * 4) / W, w = (idx * 4) % W;
    if (e < valid) *reinterpret_cast<uint4*>(tile + e * RS + w) = g4[idx];
  }
}
template <int W, int RS>
__device__ __forceinline__ void tile_store(uint32_t* gdst, const uint32_t* tile, int valid, int lane) {
  uint4* g4 = reinterpret_cast<uint4*>(gdst);
#pragma unroll
  for (int k = 0; k < W / 4; k++) {
    const int idx = lane + 32 * k, e = (idx * 4) / W, w = (idx * 4) % W;
    if (e < valid) g4[idx] = *reinterpret_cast<const uint4*>(tile + e * RS + w);
  }
}
__device__ __forceinline__ bool load_canon_s(Fp& x, const uint32_t* src) {       // 12 words from shared memory
  uint32_t w[12];
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4 v0 = s4[0], v1 = s4[1], v2 = s4[2];
  w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w;
  w[8] = v2.x; w[9] = v2.y; w[10] = v2.z; w[11] = v2.w;
  fp_unpack32(x, w);
  return fp_below_p(x);
}

__global__ void __launch_bounds__(256)
k_fp_mul(const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n, int* err) {
  constexpr int W = 12, RS = 12, WARPS = 8;
  __shared__ __align__(16) uint32_t sm[WARPS][3][32 * RS];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const size_t ntiles = (n + 31) / 32, stride = (size_t)gridDim.x * WARPS;
  for (size_t t = (size_t)blockIdx.x * WARPS + wi; t < ntiles; t += stride) {
    const size_t e0 = t * 32;
    const int valid = (int)(n - e0 < 32 ? n - e0 : 32);
    tile_load<W, RS>(sm[wi][0], a + e0 * W, valid, lane);
    tile_load<W, RS>(sm[wi][1], b + e0 * W, valid, lane);
    __syncwarp();
    if (lane < valid) {
      Fp x, y, r;
      bool ok = load_canon_s(x, sm[wi][0] + lane * RS);
      ok &= load_canon_s(y, sm[wi][1] + lane * RS);
      if (!ok) atomicOr(err, ERR_NOT_CANONICAL);
      Acc acc;
      acc_zero(acc);
      acc_mac(acc, x, y);
      acc_redc384(r, acc);
      store_canon(sm[wi][2] + lane * RS, r);
    }
    __syncwarp();
    tile_store<W, RS>(out + e0 * W, sm[wi][2], valid, lane);
    __syncwarp();
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
  constexpr int W = 24, RS = 28, WARPS = 4;
  __shared__ __align__(16) uint32_t sm[WARPS][3][32 * RS];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const size_t ntiles = (n + 31) / 32, stride = (size_t)gridDim.x * WARPS;
  for (size_t t = (size_t)blockIdx.x * WARPS + wi; t < ntiles; t += stride) {
    const size_t e0 = t * 32;
    const int valid = (int)(n - e0 < 32 ? n - e0 : 32);
    tile_load<W, RS>(sm[wi][0], a + e0 * W, valid, lane);
    tile_load<W, RS>(sm[wi][1], b + e0 * W, valid, lane);
    __syncwarp();
    if (lane < valid) {
      Fp a0, a1, b0, b1, r0, r1;
      bool ok = load_canon_s(a0, sm[wi][0] + lane * RS);
      ok &= load_canon_s(a1, sm[wi][0] + lane * RS + 12);
      ok &= load_canon_s(b0, sm[wi][1] + lane * RS);
      ok &= load_canon_s(b1, sm[wi][1] + lane * RS + 12);
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
      store_canon(sm[wi][2] + lane * RS, r0);
      store_canon(sm[wi][2] + lane * RS + 12, r1);
    }
    __syncwarp();
    tile_store<W, RS>(out + e0 * W, sm[wi][2], valid, lane);
    __syncwarp();
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
k_wire(int op, const uint32_t* in, const uint8_t* inf, int compressed, uint32_t* out, uint8_t* out8, size_t n, int* err) {
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
        r = prog_g1_deserialize(inb + (compressed ? 48 : 96) * i, compressed, out + 24 * i, &f);
        out8[i] = f;
      } break;
      case W_G1_SER: r = prog_g1_serialize(in + 24 * i, inf ? inf[i] : 0, compressed, outb + (compressed ? 48 : 96) * i); break;
      case W_G2_DESER: {
        uint8_t f = 0;
        r = prog_g2_deserialize(inb + (compressed ? 96 : 192) * i, compressed, out + 48 * i, &f);
        out8[i] = f;
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
  for (size_t base = (size_t)blockIdx.x * blockDim.x; base < n; base += (size_t)gridDim.x * blockDim.x) {
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

}  // namespace

#include "host_api.inc"
