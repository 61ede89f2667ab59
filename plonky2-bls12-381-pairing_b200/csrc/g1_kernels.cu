// g1_kernels.cu -- G1 kernels over the base field (g1.cuh): subgroup test, cofactor clearing, scalar multiplication,
// point sums and the bucket method (Pippenger).  Own translation unit of libb381.so (see g1_launch.h).
#include <cuda_runtime.h>
#include "g1_launch.h"
namespace b381_launch = b381;
// the arithmetic headers define non-inline device functions (host stubs with external linkage): this unit gets its
// own copy of them in its own namespace, so that the two objects link
#define b381 b381_g1unit
#include "g1.cuh"

using namespace b381;
using namespace b381_launch;
static_assert(G1L_RAW_AFF == G1_RAW_AFF && G1L_RAW_JAC == G1_RAW_JAC, "raw point sizes");

namespace {

__global__ void __launch_bounds__(128)
k_g1_point_op(int op, const uint32_t* pts, const uint8_t* inf, const uint32_t* scalars, uint32_t* out, uint8_t* out8, size_t n, int* err) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint8_t f = 0;
    const int in_f = inf ? inf[i] : 0;
    int r;
    if (op == PO_SUBGROUP) r = prog_g1_in_subgroup(pts + 24 * i, in_f, &f);
    else if (op == PO_CLEAR_COFACTOR) r = prog_g1_clear_cofactor(pts + 24 * i, in_f, out + 24 * i, &f);
    else r = prog_g1_scalar_mul(pts + 24 * i, in_f, scalars + 8 * i, out + 24 * i, &f);
    out8[i] = f;
    if (r) atomicOr(err, r);
  }
}

// ---- G1 bucket method (Pippenger).  Stages: points -> internal format; per-window digit histogram; exclusive scan;
// scatter of point indices by (window, digit); one thread per bucket sums its points (mixed additions); one thread
// per chunk of buckets forms the weighted chunk sum (running sums); one block per window adds the chunk sums; one
// thread combines the windows (Horner) and converts to affine.
__global__ void __launch_bounds__(128)
k_g1_to_raw(const uint32_t* pts, uint32_t* raw, size_t n, int* err) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    G1A p;
    int r = g1_load_ext(p, pts + 24 * i);
    g1_st_raw_aff(raw + (size_t)G1_RAW_AFF * i, p);
    if (r) atomicOr(err, r);
  }
}

// mode 0: count digits into cnt[w * B + d]; mode 1: scatter point indices to idx[cursor[w * B + d]++]
__global__ void __launch_bounds__(256)
k_msm_digits(const uint32_t* scalars, const uint8_t* inf, size_t n, int W, int c, unsigned int* cnt_or_cursor, uint32_t* idx, int mode) {
  const uint32_t B = 1u << c;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (inf && (inf[i] & 1)) continue;
    uint32_t k[8];
#pragma unroll
    for (int j = 0; j < 8; j++) k[j] = scalars[8 * i + j];
    for (int w = 0; w < W; w++) {
      const uint32_t d = msm_digit(k, w, c);
      if (d == 0) continue;
      if (mode == 0) atomicAdd(&cnt_or_cursor[(size_t)w * B + d], 1u);
      else idx[atomicAdd(&cnt_or_cursor[(size_t)w * B + d], 1u)] = (uint32_t)i;
    }
  }
}

// exclusive scan of cnt[0 .. m) -> start[0 .. m] (start[m] = total) and cursor = start; one block, m up to a few million
__global__ void __launch_bounds__(1024)
k_msm_scan(const unsigned int* cnt, unsigned int* start, unsigned int* cursor, size_t m) {
  __shared__ unsigned int s_part[1024];
  __shared__ unsigned int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (size_t base = 0; base < m; base += 1024) {
    const size_t i = base + threadIdx.x;
    const unsigned int v = i < m ? cnt[i] : 0u;
    s_part[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {        // Hillis-Steele inclusive scan
      unsigned int t = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0u;
      __syncthreads();
      s_part[threadIdx.x] += t;
      __syncthreads();
    }
    const unsigned int excl = s_carry + s_part[threadIdx.x] - v;
    if (i < m) { start[i] = excl; cursor[i] = excl; }
    __syncthreads();
    if (threadIdx.x == 1023) s_carry += s_part[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) start[m] = s_carry;
}

// Buckets ordered by decreasing size (counting sort over min(size, 255)): the threads of a warp then sum buckets of
// (nearly) equal size instead of waiting for the fullest of 32 random ones (sizes are Poisson: mean 16 at 2^20 points
// and 16-bit windows, the maximum of 32 is about 27).  mode 0: hist[size]++; mode 1: order[cursor[size]++] = bucket.
// Per-block histograms in shared memory keep the global atomics at 256 per block.
__global__ void __launch_bounds__(256)
k_msm_size_sort(const unsigned int* start, size_t m, unsigned int* hist_or_cursor, uint32_t* order, int mode) {
  __shared__ unsigned int s_cnt[256], s_base[256];
  const size_t per = (m + gridDim.x - 1) / gridDim.x, lo = (size_t)blockIdx.x * per, hi = lo + per < m ? lo + per : m;
  s_cnt[threadIdx.x] = 0;
  __syncthreads();
  for (size_t b = lo + threadIdx.x; b < hi; b += blockDim.x) {
    const unsigned int sz = start[b + 1] - start[b];
    atomicAdd(&s_cnt[sz < 255u ? sz : 255u], 1u);
  }
  __syncthreads();
  if (s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&hist_or_cursor[threadIdx.x], s_cnt[threadIdx.x]);
  if (mode == 0) return;
  __syncthreads();
  s_cnt[threadIdx.x] = 0;
  __syncthreads();
  for (size_t b = lo + threadIdx.x; b < hi; b += blockDim.x) {
    const unsigned int sz = start[b + 1] - start[b], bin = sz < 255u ? sz : 255u;
    order[s_base[bin] + atomicAdd(&s_cnt[bin], 1u)] = (uint32_t)b;
  }
}
// hist (256 bins) -> first rank of every size, largest sizes first
__global__ void k_msm_size_offsets(unsigned int* hist) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  unsigned int off = 0;
  for (int c = 255; c >= 0; c--) { const unsigned int t = hist[c]; hist[c] = off; off += t; }
}

// The same scan over many blocks for large m: every block scans 1024 counters (exclusive, block-relative) and reports
// its total; the totals are scanned by k_msm_scan; the offsets are added back.  1.7 ms -> 0.05 ms at 2^20 counters.
__global__ void __launch_bounds__(1024)
k_msm_scan_blocks(const unsigned int* cnt, unsigned int* start, unsigned int* blk_total, size_t m) {
  __shared__ unsigned int s_part[1024];
  const size_t i = (size_t)blockIdx.x * 1024 + threadIdx.x;
  const unsigned int v = i < m ? cnt[i] : 0u;
  s_part[threadIdx.x] = v;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    unsigned int t = threadIdx.x >= off ? s_part[threadIdx.x - off] : 0u;
    __syncthreads();
    s_part[threadIdx.x] += t;
    __syncthreads();
  }
  if (i < m) start[i] = s_part[threadIdx.x] - v;
  if (threadIdx.x == 1023) blk_total[blockIdx.x] = s_part[1023];
}
__global__ void __launch_bounds__(1024)
k_msm_scan_add(unsigned int* start, unsigned int* cursor, const unsigned int* blk_start, size_t m, size_t nblk) {
  const size_t i = (size_t)blockIdx.x * 1024 + threadIdx.x;
  if (i < m) {
    const unsigned int v = start[i] + blk_start[blockIdx.x];
    start[i] = v;
    cursor[i] = v;
  }
  if (i == 0) start[m] = blk_start[nblk];
}

// one thread per bucket (w, d), taken in the order of decreasing size: sum of its points -> buckets[(w * B + d)] (raw Jacobian)
__global__ void __launch_bounds__(128)
k_msm_bucket_sums(const uint32_t* pts_raw, const uint32_t* idx, const unsigned int* start, const uint32_t* order, size_t m, uint32_t* buckets) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < m; t += (size_t)gridDim.x * blockDim.x) {
    const size_t b = order[t];
    G1J acc;
    msm_bucket_sum(acc, pts_raw, idx, start[b], start[b + 1]);
    g1_st_raw_jac(buckets + (size_t)G1_RAW_JAC * b, acc);
  }
}

// one thread per chunk of CH buckets of a window: partial[w * nchunk + j] = sum_{d in chunk} d B_{w,d}
__global__ void __launch_bounds__(128)
k_msm_chunks(const uint32_t* buckets, int W, int c, int CH, uint32_t* partial) {
  const uint32_t B = 1u << c, nchunk = (B + CH - 1) / CH;
  const size_t total = (size_t)W * nchunk;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const uint32_t w = (uint32_t)(t / nchunk), j = (uint32_t)(t % nchunk);
    uint32_t lo = j * CH, hi = lo + CH < B ? lo + CH : B;
    if (lo == 0) lo = 1;                              // digit 0 contributes nothing
    G1J r;
    if (lo < hi) msm_chunk_weighted(r, buckets + (size_t)G1_RAW_JAC * w * B, lo, hi);
    else g1_set_identity(r);
    g1_st_raw_jac(partial + (size_t)G1_RAW_JAC * t, r);
  }
}

// sums[j] = sum of in[j * per .. min((j+1) per, n_in)) (raw Jacobian points); one thread per output
__global__ void __launch_bounds__(128)
k_g1_jac_sums(const uint32_t* in, size_t n_in, size_t per, uint32_t* sums, size_t n_out) {
  for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_out; j += (size_t)gridDim.x * blockDim.x) {
    G1J acc;
    g1_set_identity(acc);
    const size_t lo = j * per, hi = lo + per < n_in ? lo + per : n_in;
    for (size_t t = lo; t < hi; t++) {
      G1J q;
      g1_ld_raw_jac(q, in + (size_t)G1_RAW_JAC * t);
      g1_add(acc, acc, q);
    }
    g1_st_raw_jac(sums + (size_t)G1_RAW_JAC * j, acc);
  }
}

// sums[w] = sum of the nchunk chunk results of window w: one block per window; every thread adds its strided share,
// then a shuffle tree inside the warps and one across them: 8 + 5 + 3 sequential additions for 2048 chunks instead of
// the 32 + 64 of two levels of one-thread sums
__device__ __forceinline__ void g1_shfl_down(G1J& r, const G1J& p, int delta) {
#pragma unroll
  for (int k = 0; k < NL; k++) {
    r.x.l[k] = __shfl_down_sync(0xffffffffu, p.x.l[k], delta);
    r.y.l[k] = __shfl_down_sync(0xffffffffu, p.y.l[k], delta);
    r.z.l[k] = __shfl_down_sync(0xffffffffu, p.z.l[k], delta);
  }
}
__global__ void __launch_bounds__(256)
k_g1_window_tree(const uint32_t* partial, size_t nchunk, uint32_t* sums) {
  __shared__ uint32_t s_w[8][G1_RAW_JAC];
  const uint32_t* in = partial + (size_t)G1_RAW_JAC * blockIdx.x * nchunk;
  G1J acc;
  g1_set_identity(acc);
  for (size_t t = threadIdx.x; t < nchunk; t += 256) {
    G1J q;
    g1_ld_raw_jac(q, in + (size_t)G1_RAW_JAC * t);
    g1_add(acc, acc, q);
  }
  for (int d = 16; d >= 1; d >>= 1) {
    G1J q;
    g1_shfl_down(q, acc, d);
    g1_add(acc, acc, q);
  }
  if ((threadIdx.x & 31) == 0) g1_st_raw_jac(s_w[threadIdx.x >> 5], acc);
  __syncthreads();
  if (threadIdx.x < 32) {
    if (threadIdx.x < 8) g1_ld_raw_jac(acc, s_w[threadIdx.x]);
    else g1_set_identity(acc);
    for (int d = 4; d >= 1; d >>= 1) {
      G1J q;
      g1_shfl_down(q, acc, d);
      g1_add(acc, acc, q);
    }
    if (threadIdx.x == 0) g1_st_raw_jac(sums + (size_t)G1_RAW_JAC * blockIdx.x, acc);
  }
}

// Horner over the W window sums (c doublings per window) and conversion to affine; W = 1, c = 0: plain conversion
__global__ void k_msm_final(const uint32_t* sums, int W, int c, uint32_t* out24, uint8_t* out_inf) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  G1J r;
  msm_combine_windows(r, sums, W, c);
  g1_store_jac_ext(out24, out_inf, r);
}

// partial sums of affine external points (for b381_g1_sum): thread t adds points t, t + T, ... -> raw Jacobian
__global__ void __launch_bounds__(128)
k_g1_sum_strided(const uint32_t* pts, const uint8_t* inf, size_t n, uint32_t* partial, size_t T, int* err) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (size_t)gridDim.x * blockDim.x) {
    G1J acc;
    g1_set_identity(acc);
    int e = 0;
    for (size_t i = t; i < n; i += T) {
      if (inf && (inf[i] & 1)) continue;
      G1A q;
      e |= g1_load_ext(q, pts + 24 * i);
      g1_add_mixed(acc, acc, q);
    }
    g1_st_raw_jac(partial + (size_t)G1_RAW_JAC * t, acc);
    if (e) atomicOr(err, e);
  }
}

}  // namespace

#undef b381
namespace b381 {

int g1l_point_op(int grid, cudaStream_t s, int op, const uint32_t* pts, const uint8_t* inf, const uint32_t* scalars, uint32_t* out, uint8_t* out8, size_t n, int* err) {
  k_g1_point_op<<<grid, 128, 0, s>>>(op, pts, inf, scalars, out, out8, n, err);
  return 1;
}
int g1l_to_raw(int grid, cudaStream_t s, const uint32_t* pts, uint32_t* raw, size_t n, int* err) {
  k_g1_to_raw<<<grid, 128, 0, s>>>(pts, raw, n, err);
  return 1;
}
int g1l_msm_digits(int grid, cudaStream_t s, const uint32_t* scalars, const uint8_t* inf, size_t n, int W, int c, unsigned int* cnt_or_cursor, uint32_t* idx, int mode) {
  k_msm_digits<<<grid, 256, 0, s>>>(scalars, inf, n, W, c, cnt_or_cursor, idx, mode);
  return 1;
}
int g1l_msm_scan(cudaStream_t s, const unsigned int* cnt, unsigned int* start, unsigned int* cursor, size_t m, unsigned int* scratch) {
  if (m <= 16384) {
    k_msm_scan<<<1, 1024, 0, s>>>(cnt, start, cursor, m);
    return 1;
  }
  const size_t nblk = (m + 1023) / 1024;             // scratch: 3 x (nblk + 1) words
  unsigned int* blk_total = scratch;
  unsigned int* blk_start = scratch + (nblk + 1);
  unsigned int* blk_cursor = scratch + 2 * (nblk + 1);
  k_msm_scan_blocks<<<(unsigned)nblk, 1024, 0, s>>>(cnt, start, blk_total, m);
  k_msm_scan<<<1, 1024, 0, s>>>(blk_total, blk_start, blk_cursor, nblk);
  k_msm_scan_add<<<(unsigned)nblk, 1024, 0, s>>>(start, cursor, blk_start, m, nblk);
  return 3;
}
int g1l_msm_size_order(int grid, cudaStream_t s, const unsigned int* start, size_t m, unsigned int* hist256, uint32_t* order) {
  cudaMemsetAsync(hist256, 0, 256 * sizeof(unsigned int), s);
  k_msm_size_sort<<<grid, 256, 0, s>>>(start, m, hist256, order, 0);
  k_msm_size_offsets<<<1, 32, 0, s>>>(hist256);
  k_msm_size_sort<<<grid, 256, 0, s>>>(start, m, hist256, order, 1);
  return 3;
}
int g1l_msm_bucket_sums(int grid, cudaStream_t s, const uint32_t* pts_raw, const uint32_t* idx, const unsigned int* start, const uint32_t* order, size_t m, uint32_t* buckets) {
  k_msm_bucket_sums<<<grid, 128, 0, s>>>(pts_raw, idx, start, order, m, buckets);
  return 1;
}
int g1l_msm_chunks(int grid, cudaStream_t s, const uint32_t* buckets, int W, int c, int CH, uint32_t* partial) {
  k_msm_chunks<<<grid, 128, 0, s>>>(buckets, W, c, CH, partial);
  return 1;
}
int g1l_jac_sums(int grid, cudaStream_t s, const uint32_t* in, size_t n_in, size_t per, uint32_t* sums, size_t n_out) {
  k_g1_jac_sums<<<grid, 128, 0, s>>>(in, n_in, per, sums, n_out);
  return 1;
}
int g1l_window_tree(cudaStream_t s, const uint32_t* partial, int W, size_t nchunk, uint32_t* sums) {
  k_g1_window_tree<<<W, 256, 0, s>>>(partial, nchunk, sums);
  return 1;
}
int g1l_msm_final(cudaStream_t s, const uint32_t* sums, int W, int c, uint32_t* out24, uint8_t* out_inf) {
  k_msm_final<<<1, 32, 0, s>>>(sums, W, c, out24, out_inf);
  return 1;
}
int g1l_sum_strided(int grid, cudaStream_t s, const uint32_t* pts, const uint8_t* inf, size_t n, uint32_t* partial, size_t T, int* err) {
  k_g1_sum_strided<<<grid, 128, 0, s>>>(pts, inf, n, partial, T, err);
  return 1;
}

}  // namespace b381
