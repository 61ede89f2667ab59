// imad_probe4.cu -- signed vs unsigned IMAD.WIDE rate (register operands and immediates)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
typedef unsigned long long u64;
#define MACU(c,a,b) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c) : "r"(a), "r"(b))
#define MACS(c,a,b) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(c) : "r"(a), "r"(b))
#define MACUI(c,a,i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c) : "r"(a), "n"(i))
#define MACSI(c,a,i) asm volatile("mad.wide.s32 %0, %1, %2, %0;" : "+l"(c) : "r"(a), "n"(i))
template<int MODE>
__global__ void __launch_bounds__(512) probe(uint32_t* out, const uint32_t* in, u64* cyc, int iters) {
  uint32_t a[14], b[14]; u64 c[28];
  for (int k = 0; k < 14; k++) { a[k] = (in[k] ^ threadIdx.x) & 0xfffffff; b[k] = (in[20 + k] + threadIdx.x) & 0xfffffff; }
  for (int k = 0; k < 28; k++) c[k] = in[40 + k];
  __syncthreads(); u64 t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 14; i++) {
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 14; j++) MACU(c[i + j], a[i], b[j]);
      } else if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 14; j++) MACS(c[i + j], a[i], b[j]);
      } else if (MODE == 2) {
        MACUI(c[i+0], a[i], 0xfffaaab); MACUI(c[i+1], a[i], 0xfefffff); MACUI(c[i+2], a[i], 0x3ffffb9); MACUI(c[i+3], a[i], 0xfffeb15);
        MACUI(c[i+4], a[i], 0x6241eab); MACUI(c[i+5], a[i], 0xa0f6b0f); MACUI(c[i+6], a[i], 0xf6730d2); MACUI(c[i+7], a[i], 0xf38512b);
        MACUI(c[i+8], a[i], 0x4774b84); MACUI(c[i+9], a[i], 0x4bacd76); MACUI(c[i+10], a[i], 0xba7b643); MACUI(c[i+11], a[i], 0xe69a4b1);
        MACUI(c[i+12], a[i], 0x1ea397f); MACUI(c[i+13], a[i], 0x001a011);
      } else {
        MACSI(c[i+0], a[i], 0xfffaaab); MACSI(c[i+1], a[i], 0xfefffff); MACSI(c[i+2], a[i], 0x3ffffb9); MACSI(c[i+3], a[i], 0xfffeb15);
        MACSI(c[i+4], a[i], 0x6241eab); MACSI(c[i+5], a[i], 0xa0f6b0f); MACSI(c[i+6], a[i], 0xf6730d2); MACSI(c[i+7], a[i], 0xf38512b);
        MACSI(c[i+8], a[i], 0x4774b84); MACSI(c[i+9], a[i], 0x4bacd76); MACSI(c[i+10], a[i], 0xba7b643); MACSI(c[i+11], a[i], 0xe69a4b1);
        MACSI(c[i+12], a[i], 0x1ea397f); MACSI(c[i+13], a[i], 0x001a011);
      }
    }
  }
  u64 t1 = clock64(), s = 0;
  for (int k = 0; k < 28; k++) s ^= c[k];
  if (s == 0x123456789ull) out[threadIdx.x] = (uint32_t)s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template<int MODE>
void run(const char* name, int warps, int nsm, uint32_t* dout, uint32_t* din, u64* dcyc) {
  const int iters = 3000;
  probe<MODE><<<nsm, warps * 32>>>(dout, din, dcyc, 10); CK(cudaDeviceSynchronize());
  probe<MODE><<<nsm, warps * 32>>>(dout, din, dcyc, iters); CK(cudaDeviceSynchronize());
  static u64 h[1024]; CK(cudaMemcpy(h, dcyc, nsm * 8, cudaMemcpyDeviceToHost));
  double cavg = 0; for (int i = 0; i < nsm; i++) cavg += (double)h[i]; cavg /= nsm;
  printf("{\"probe\":\"%s\",\"warps_per_sm\":%d,\"mac_per_clk_per_sm\":%.2f}\n", name, warps, 196.0 * iters * warps * 32 / cavg);
}
int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); int nsm = prop.multiProcessorCount;
  uint32_t *dout, *din; u64* dcyc;
  CK(cudaMalloc(&dout, 4096 * 4)); CK(cudaMalloc(&din, 4096 * 4)); CK(cudaMalloc(&dcyc, 1024 * 8));
  uint32_t h[4096]; for (int i = 0; i < 4096; i++) h[i] = 0x9e3779b9u * (i + 1) | 1;
  CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
  for (int w : {4, 8}) run<0>("block_unsigned_reg", w, nsm, dout, din, dcyc);
  for (int w : {4, 8}) run<1>("block_signed_reg", w, nsm, dout, din, dcyc);
  for (int w : {4, 8}) run<2>("rows_unsigned_imm", w, nsm, dout, din, dcyc);
  for (int w : {4, 8}) run<3>("rows_signed_imm", w, nsm, dout, din, dcyc);
  return 0;
}
