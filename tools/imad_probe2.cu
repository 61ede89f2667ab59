// imad_probe2.cu -- IMAD.WIDE rate for MAC-block-like operand patterns (distinct registers, signed vs unsigned)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

// 14x14 MAC block into 27 columns, like acc_mac; SIGNED selects int vs unsigned
template<bool SIGNED, int NB>
__global__ void __launch_bounds__(1024) probe(uint32_t* out, const uint32_t* in, unsigned long long* cyc, int iters) {
  int a[14], b[14];
  long long c[28];
  for (int k = 0; k < 14; k++) { a[k] = (in[k] ^ threadIdx.x) & 0x0fffffff; b[k] = (in[20 + k] + threadIdx.x) & 0x0fffffff; }
  for (int k = 0; k < 28; k++) c[k] = in[40 + k];
  __syncthreads();
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < NB; r++) {
#pragma unroll
      for (int i = 0; i < 14; i++)
#pragma unroll
        for (int j = 0; j < 14; j++) {
          if (SIGNED) c[i + j] += (long long)a[i] * (long long)b[j];
          else c[i + j] += (long long)((unsigned long long)(unsigned)a[i] * (unsigned long long)(unsigned)b[j]);
        }
      // light dependency so blocks cannot be merged: rotate a
#pragma unroll
      for (int k = 0; k < 14; k++) a[k] = (a[k] + (int)(c[k] & 1));
    }
  }
  unsigned long long t1 = clock64();
  long long s = 0;
  for (int k = 0; k < 28; k++) s ^= c[k];
  if (s == 0x123456789ll) out[threadIdx.x] = (uint32_t)s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template<bool SIGNED>
void run(const char* name, int warps, int nsm, uint32_t* dout, uint32_t* din, unsigned long long* dcyc) {
  const int iters = 2000, NB = 2;
  probe<SIGNED, NB><<<nsm, warps * 32>>>(dout, din, dcyc, 10);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<SIGNED, NB><<<nsm, warps * 32>>>(dout, din, dcyc, iters);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  static unsigned long long h[1024];
  CK(cudaMemcpy(h, dcyc, nsm * 8, cudaMemcpyDeviceToHost));
  double cavg = 0; for (int i = 0; i < nsm; i++) cavg += (double)h[i]; cavg /= nsm;
  double inst = 196.0 * NB * iters * warps * 32;
  printf("{\"probe\":\"%s\",\"warps_per_sm\":%d,\"imad_wide_per_clk_per_sm\":%.2f,\"chip_Ginst_per_s\":%.1f}\n", name, warps, inst / cavg, inst * nsm / (ms * 1e-3) / 1e9);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  uint32_t *dout, *din; unsigned long long* dcyc;
  CK(cudaMalloc(&dout, 4096 * 4)); CK(cudaMalloc(&din, 4096 * 4)); CK(cudaMalloc(&dcyc, 1024 * 8));
  uint32_t h[4096]; for (int i = 0; i < 4096; i++) h[i] = 0x9e3779b9u * (i + 1) | 1;
  CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
  for (int w : {4, 8, 16}) run<true>("mac_block_signed", w, nsm, dout, din, dcyc);
  for (int w : {4, 8, 16}) run<false>("mac_block_unsigned", w, nsm, dout, din, dcyc);
  return 0;
}
