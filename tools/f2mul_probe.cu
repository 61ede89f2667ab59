// f2mul_probe.cu -- register-resident chained Fp2 multiply / square: cycles per op per warp.
// Build variants with -DB381_MAC_STYLE=0 (plain C, ptxas free to reorder) or 1 (volatile ordered MACs).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../plonky2-bls12-381-pairing_b200/csrc/tower.cuh"
using namespace b381;
#if B381_FMT == 32
#define PMASK 0x0fffffffu       /* keeps the top word small: operands stay far below 2^17 p */
#define B381_MAC_STYLE 32
#define IMADS_MUL (3.0 * 169 + 2 * 156)
#define IMADS_SQR (2.0 * 169 + 2 * 156)
#else
#define PMASK MASK
#define IMADS_MUL 1038.0
#define IMADS_SQR 842.0
#endif
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

template<int OP>
__global__ void __launch_bounds__(512, 1) probe(uint32_t* out, const uint32_t* in, unsigned long long* cyc, int iters) {
  Fp a0, a1, b0, b1;
  for (int k = 0; k < NL; k++) {
    a0.l[k] = (in[k] + threadIdx.x) & PMASK; a1.l[k] = (in[20 + k] ^ threadIdx.x) & PMASK;
    b0.l[k] = (in[40 + k] + 3 * threadIdx.x) & PMASK; b1.l[k] = (in[60 + k] + 7 * threadIdx.x) & PMASK;
  }
  __syncthreads();
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    Fp r0, r1;
    if (OP == 0) f2_mul_reg(r0, r1, a0, a1, b0, b1);
    else f2_sqr_reg(r0, r1, a0, a1);
    a0 = r0; a1 = r1;
  }
  unsigned long long t1 = clock64();
  uint32_t s = 0;
  for (int k = 0; k < NL; k++) s ^= a0.l[k] ^ a1.l[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template<int OP>
void run(const char* name, int warps, int nsm, uint32_t* dout, uint32_t* din, unsigned long long* dcyc) {
  const int iters = 2000;
  probe<OP><<<nsm, warps * 32>>>(dout, din, dcyc, 10);
  CK(cudaDeviceSynchronize());
  probe<OP><<<nsm, warps * 32>>>(dout, din, dcyc, iters);
  CK(cudaDeviceSynchronize());
  static unsigned long long h[1024];
  CK(cudaMemcpy(h, dcyc, nsm * 8, cudaMemcpyDeviceToHost));
  double cavg = 0; for (int i = 0; i < nsm; i++) cavg += (double)h[i]; cavg /= nsm;
  double per_smsp = cavg / iters / (warps / 4.0);
  printf("{\"probe\":\"%s\",\"style\":%d,\"warps_per_sm\":%d,\"cycles_per_op_per_warp_slot\":%.0f,\"imad_per_clk_per_sm\":%.1f}\n", name, B381_MAC_STYLE, warps, per_smsp,
         (OP == 0 ? IMADS_MUL : IMADS_SQR) * 32 * 4 / per_smsp);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  uint32_t *dout, *din; unsigned long long* dcyc;
  CK(cudaMalloc(&dout, 148 * 1024 * 4)); CK(cudaMalloc(&din, 4096 * 4)); CK(cudaMalloc(&dcyc, 1024 * 8));
  uint32_t h[4096]; for (int i = 0; i < 4096; i++) h[i] = 0x9e3779b9u * (i + 1) | 1;
  CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
  for (int w : {4, 8, 12, 16}) run<0>("f2_mul", w, nsm, dout, din, dcyc);
  for (int w : {4, 8, 12, 16}) run<1>("f2_sqr", w, nsm, dout, din, dcyc);
  return 0;
}
