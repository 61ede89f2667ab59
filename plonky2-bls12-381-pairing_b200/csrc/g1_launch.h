// g1_launch.h -- launchers of the G1 / bucket-method kernels (g1_kernels.cu), called by the host side in host_api.inc.
// The G1 kernels are a separate translation unit: their fully inlined group law dominates the build time, and
// compiling them with -split-compile (which costs the pairing kernels 2.6 %) keeps the whole build at a few minutes.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b381 {

enum PointOp { PO_SUBGROUP = 0, PO_CLEAR_COFACTOR, PO_SCALAR_MUL };
constexpr int G1L_RAW_AFF = 24, G1L_RAW_JAC = 36;      // words of a raw affine / Jacobian point (= g1.cuh G1_RAW_*)

// every launcher enqueues on `s` and returns the number of kernels it launched
int g1l_point_op(int grid, cudaStream_t s, int op, const uint32_t* pts, const uint8_t* inf, const uint32_t* scalars, uint32_t* out, uint8_t* out8, size_t n, int* err);
int g1l_to_raw(int grid, cudaStream_t s, const uint32_t* pts, uint32_t* raw, size_t n, int* err);
int g1l_msm_digits(int grid, cudaStream_t s, const uint32_t* scalars, const uint8_t* inf, size_t n, int W, int c, unsigned int* cnt_or_cursor, uint32_t* idx, int mode);
// scratch: 3 x (ceil(m / 1024) + 1) words (used when m > 16384)
int g1l_msm_scan(cudaStream_t s, const unsigned int* cnt, unsigned int* start, unsigned int* cursor, size_t m, unsigned int* scratch);
// order[0 .. m) = the buckets by decreasing size (hist256: 256 words of scratch); shared by the G1 and G2 bucket sums
int g1l_msm_size_order(int grid, cudaStream_t s, const unsigned int* start, size_t m, unsigned int* hist256, uint32_t* order);
int g1l_msm_bucket_sums(int grid, cudaStream_t s, const uint32_t* pts_raw, const uint32_t* idx, const unsigned int* start, const uint32_t* order, size_t m, uint32_t* buckets);
int g1l_msm_chunks(int grid, cudaStream_t s, const uint32_t* buckets, int W, int c, int CH, uint32_t* partial);
int g1l_jac_sums(int grid, cudaStream_t s, const uint32_t* in, size_t n_in, size_t per, uint32_t* sums, size_t n_out);
int g1l_window_tree(cudaStream_t s, const uint32_t* partial, int W, size_t nchunk, uint32_t* sums);
int g1l_msm_final(cudaStream_t s, const uint32_t* sums, int W, int c, uint32_t* out24, uint8_t* out_inf);
int g1l_sum_strided(int grid, cudaStream_t s, const uint32_t* pts, const uint8_t* inf, size_t n, uint32_t* partial, size_t T, int* err);

}  // namespace b381
