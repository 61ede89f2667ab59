// phase_probe.cu -- cycles of the phases of an Fp2 multiply in isolation (8 warps/SM): MAC blocks vs reductions
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../plonky2-bls12-381-pairing_b200/csrc/tower.cuh"
using namespace b381;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int MODE>
__global__ void __launch_bounds__(256, 1) probe(uint32_t* out, const uint32_t* in, unsigned long long* cyc, int iters) {
  Fp a0, a1, b0, b1;
  for (int k = 0; k < NL; k++) {
    a0.l[k] = (in[k] + threadIdx.x) & MASK; a1.l[k] = (in[20 + k] ^ threadIdx.x) & MASK;
    b0.l[k] = (in[40 + k] + 3 * threadIdx.x) & MASK; b1.l[k] = (in[60 + k] + 7 * threadIdx.x) & MASK;
  }
  Acc A, B;
  acc_zero(A); acc_zero(B);
  __syncthreads();
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {          // three MAC blocks, operands perturbed by the result (no hoisting)
      acc_mac(A, a0, b0); acc_mac(B, a1, b1); acc_mac(A, a1, b0);
#pragma unroll
      for (int k = 0; k < NL; k++) { a0.l[k] = (a0.l[k] + (int)(A.c[k] & 1)) & MASK; a1.l[k] = (a1.l[k] + (int)(B.c[k] & 1)) & MASK; }
    } else if (MODE == 1) {   // two interleaved reductions of data-dependent accumulators
      Fp r0, r1;
#pragma unroll
      for (int k = 0; k < 2 * NL - 1; k++) { A.c[k] = (int64_t)a0.l[k % NL] * 0x7654321 + k; B.c[k] = (int64_t)a1.l[k % NL] * 0x1234567 + k; }
#pragma unroll
      for (int k = 2 * NL - 1; k < NCOL; k++) { A.c[k] = 0; B.c[k] = 0; }
      acc_redc2(r0, A, r1, B);
      a0 = r0; a1 = r1;
    } else {                  // single reduction
      Fp r0;
#pragma unroll
      for (int k = 0; k < 2 * NL - 1; k++) A.c[k] = (int64_t)a0.l[k % NL] * 0x7654321 + k;
#pragma unroll
      for (int k = 2 * NL - 1; k < NCOL; k++) A.c[k] = 0;
      acc_redc(r0, A);
      a0 = r0;
    }
  }
  unsigned long long t1 = clock64();
  uint32_t s = 0;
  for (int k = 0; k < NL; k++) s ^= a0.l[k] ^ a1.l[k];
  for (int k = 0; k < NCOL; k++) s ^= (uint32_t)A.c[k] ^ (uint32_t)B.c[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template<int MODE>
void run(const char* name, int imads, int nsm, uint32_t* dout, uint32_t* din, unsigned long long* dcyc) {
  const int iters = 2000;
  for (int threads : {128, 256}) {
    probe<MODE><<<nsm, threads>>>(dout, din, dcyc, 10); CK(cudaDeviceSynchronize());
    probe<MODE><<<nsm, threads>>>(dout, din, dcyc, iters); CK(cudaDeviceSynchronize());
    static unsigned long long h[1024]; CK(cudaMemcpy(h, dcyc, nsm * 8, cudaMemcpyDeviceToHost));
    double cavg = 0; for (int i = 0; i < nsm; i++) cavg += (double)h[i]; cavg /= nsm;
    double per = cavg / iters / (threads / 128.0);
    printf("{\"phase\":\"%s\",\"threads_per_sm\":%d,\"cycles_per_iter_per_warp_slot\":%.0f,\"imad_per_clk_per_sm\":%.1f}\n", name, threads, per, imads * 128.0 / per);
  }
}
int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); int nsm = prop.multiProcessorCount;
  uint32_t *dout, *din; unsigned long long* dcyc;
  CK(cudaMalloc(&dout, 148 * 1024 * 4)); CK(cudaMalloc(&din, 4096 * 4)); CK(cudaMalloc(&dcyc, 1024 * 8));
  uint32_t h[4096]; for (int i = 0; i < 4096; i++) h[i] = 0x9e3779b9u * (i + 1) | 1;
  CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
  run<0>("3_mac_blocks", 588, nsm, dout, din, dcyc);
  run<1>("redc2", 450 + 55, nsm, dout, din, dcyc);
  run<2>("redc1", 225 + 27, nsm, dout, din, dcyc);
  return 0;
}
