// mul_probe.cu -- register-resident chains of the hot multiplication primitives (f2_mul_reg, f2_sqr_reg,
// the three-product sum of f2_sop) on all SMs: executed IMAD.WIDE per clock per SM, and a checksum so that
// two builds (-DB381_NO_PAIRMUL: generic products; default: register-bank-aware paired products) can be
// compared for identical values.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-DB381_NO_PAIRMUL] -o build/mul_probe tools/mul_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../plonky2-bls12-381-pairing_b200/csrc/tower.cuh"
using namespace b381;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

// OP 0: f2_mul_reg (3 x 144 + 2 x 156), 1: f2_sqr_reg (2 x 144 + 2 x 156), 2: sum of three Fp2 products as in f2_sop (9 x 144 + 2 x 156)
template <int OP>
__global__ void __launch_bounds__(256, 1) probe(uint32_t* out, const uint32_t* in, unsigned long long* cyc, int iters) {
  Fp a0, a1, b0, b1, c0, c1;
  for (int k = 0; k < NL; k++) {
    a0.l[k] = in[k] + threadIdx.x; a1.l[k] = in[20 + k] ^ threadIdx.x;
    b0.l[k] = in[40 + k] + 3 * threadIdx.x; b1.l[k] = in[60 + k] + 7 * threadIdx.x;
    c0.l[k] = in[80 + k] + 5 * threadIdx.x; c1.l[k] = in[100 + k] + 11 * threadIdx.x;
  }
  // operands like stored values: below 2^381 (top word zero, word 11 small)
  a0.l[12] = a1.l[12] = b0.l[12] = b1.l[12] = c0.l[12] = c1.l[12] = 0;
  a0.l[11] &= 0x0fffffffu; a1.l[11] &= 0x0fffffffu; b0.l[11] &= 0x0fffffffu; b1.l[11] &= 0x0fffffffu; c0.l[11] &= 0x0fffffffu; c1.l[11] &= 0x0fffffffu;
  __syncthreads();
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    Fp r0, r1;
    if (OP == 0) f2_mul_reg(r0, r1, a0, a1, b0, b1);
    if (OP == 1) f2_sqr_reg(r0, r1, a0, a1);
    if (OP == 2) {
      Acc P, Q, X;
      HOT_MUL3(P, a0, b0, b0, c0, c0, a0);
      HOT_MUL3(Q, a1, b1, b1, c1, c1, a1);
      Fp s0, s1, s2;
      fp_add(s0, a0, a1); fp_add(s1, b0, b1); fp_add(s2, c0, c1);
      HOT_MUL3(X, s0, s1, s1, s2, s2, s0);
      acc_sub(X, X, P); acc_sub(X, X, Q);
      acc_sub(P, P, Q);
      acc_redc2(r0, P, r1, X);
      fp_add_p(r0, r0);
    }
    a0 = r0; a1 = r1;
    a0.l[12] = 0; a1.l[12] = 0; a0.l[11] &= 0x0fffffffu; a1.l[11] &= 0x0fffffffu;     // keep the chain inside the 12-word range
  }
  unsigned long long t1 = clock64();
  uint32_t s = 0;
  for (int k = 0; k < NL; k++) s = s * 31 + (a0.l[k] ^ (a1.l[k] * 7u));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, double imads, int warps, int nsm, uint32_t* dout, uint32_t* din, unsigned long long* dcyc) {
  const int iters = 1000;
  probe<OP><<<nsm, warps * 32>>>(dout, din, dcyc, 10);
  CK(cudaDeviceSynchronize());
  probe<OP><<<nsm, warps * 32>>>(dout, din, dcyc, iters);
  CK(cudaDeviceSynchronize());
  static unsigned long long h[1024];
  static uint32_t ho[148 * 256];
  CK(cudaMemcpy(h, dcyc, nsm * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ho, dout, (size_t)nsm * warps * 32 * 4, cudaMemcpyDeviceToHost));
  unsigned long long sum = 0;
  for (int i = 0; i < nsm * warps * 32; i++) sum = sum * 1000003ull + ho[i];
  double cavg = 0; for (int i = 0; i < nsm; i++) cavg += (double)h[i]; cavg /= nsm;
  printf("{\"probe\":\"%s\",\"pairmul\":%d,\"warps_per_sm\":%d,\"cycles_per_op\":%.0f,\"imad_wide_per_clk_per_sm\":%.2f,\"checksum\":\"%016llx\"}\n", name,
#ifdef B381_NO_PAIRMUL
         0,
#else
         1,
#endif
         warps, cavg / iters, imads * 32 * warps * iters / cavg, sum);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  uint32_t *dout, *din; unsigned long long* dcyc;
  CK(cudaMalloc(&dout, 148 * 1024 * 4)); CK(cudaMalloc(&din, 4096 * 4)); CK(cudaMalloc(&dcyc, 1024 * 8));
  uint32_t h[4096]; for (int i = 0; i < 4096; i++) h[i] = 0x9e3779b9u * (i + 1) | 1;
  CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
  for (int w : {4, 8}) run<0>("f2_mul", 3.0 * 144 + 2 * 156, w, nsm, dout, din, dcyc);
  for (int w : {4, 8}) run<1>("f2_sqr", 2.0 * 144 + 2 * 156, w, nsm, dout, din, dcyc);
  for (int w : {4, 8}) run<2>("sop3", 9.0 * 144 + 2 * 156, w, nsm, dout, din, dcyc);
  return 0;
}
