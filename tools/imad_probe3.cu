// imad_probe3.cu -- what limits IMAD.WIDE in MAC blocks: operand register parity / reuse?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
typedef unsigned long long u64;

template<int MODE>
__global__ void __launch_bounds__(512) probe(uint32_t* out, const uint32_t* in, u64* cyc, int iters) {
  u64 t0, t1, s = 0;
  if (MODE == 1) {          // a fixed, b varies, 14 accumulators
    uint32_t a = in[0] ^ threadIdx.x, b[14]; u64 c[14];
    for (int k = 0; k < 14; k++) { b[k] = in[20 + k] + threadIdx.x; c[k] = in[40 + k]; }
    __syncthreads(); t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 14; r++)
#pragma unroll
        for (int j = 0; j < 14; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[j]) : "r"(a), "r"(b[j]));
    }
    t1 = clock64(); for (int k = 0; k < 14; k++) s ^= c[k];
  } else if (MODE == 2) {   // a and b both vary (same index), 14 accumulators
    uint32_t a[14], b[14]; u64 c[14];
    for (int k = 0; k < 14; k++) { a[k] = in[k] ^ threadIdx.x; b[k] = in[20 + k] + threadIdx.x; c[k] = in[40 + k]; }
    __syncthreads(); t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int r = 0; r < 14; r++)
#pragma unroll
        for (int j = 0; j < 14; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[j]) : "r"(a[j]), "r"(b[(j + r) % 14]));
    }
    t1 = clock64(); for (int k = 0; k < 14; k++) s ^= c[k];
  } else if (MODE == 3 || MODE == 4 || MODE == 5) {   // full MAC block, asm volatile in source order (row-major / column-major)
    uint32_t a[14], b[14]; u64 c[28];
    for (int k = 0; k < 14; k++) { a[k] = in[k] ^ threadIdx.x; b[k] = in[20 + k] + threadIdx.x; }
    for (int k = 0; k < 28; k++) c[k] = in[40 + k];
    __syncthreads(); t0 = clock64();
    for (int it = 0; it < iters; it++) {
      if (MODE == 3) {
#pragma unroll
        for (int i = 0; i < 14; i++)
#pragma unroll
          for (int j = 0; j < 14; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[i + j]) : "r"(a[i]), "r"(b[j]));
      } else if (MODE == 4) {   // non-volatile C expression (compiler free to reorder)
#pragma unroll
        for (int i = 0; i < 14; i++)
#pragma unroll
          for (int j = 0; j < 14; j++) c[i + j] += (u64)a[i] * (u64)b[j];
#pragma unroll
        for (int k = 0; k < 14; k++) a[k] += (uint32_t)(c[k] & 1);
      } else {                  // packed operands: a_i in the low half, b_i in the high half of one 64-bit register
        u64 ab[14];
#pragma unroll
        for (int k = 0; k < 14; k++) ab[k] = ((u64)b[k] << 32) | a[k];
#pragma unroll
        for (int i = 0; i < 14; i++)
#pragma unroll
          for (int j = 0; j < 14; j++) {
            uint32_t x, y, z, w;
            asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(x), "=r"(y) : "l"(ab[i]));
            asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(z), "=r"(w) : "l"(ab[j]));
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[i + j]) : "r"(x), "r"(w));
          }
      }
    }
    t1 = clock64(); for (int k = 0; k < 28; k++) s ^= c[k];
  }
  if (s == 0x123456789ull) out[threadIdx.x] = (uint32_t)s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template<int MODE>
void run(const char* name, int warps, int nsm, uint32_t* dout, uint32_t* din, u64* dcyc) {
  const int iters = 3000;
  probe<MODE><<<nsm, warps * 32>>>(dout, din, dcyc, 10);
  CK(cudaDeviceSynchronize());
  probe<MODE><<<nsm, warps * 32>>>(dout, din, dcyc, iters);
  CK(cudaDeviceSynchronize());
  static u64 h[1024];
  CK(cudaMemcpy(h, dcyc, nsm * 8, cudaMemcpyDeviceToHost));
  double cavg = 0; for (int i = 0; i < nsm; i++) cavg += (double)h[i]; cavg /= nsm;
  double inst = 196.0 * iters * warps * 32;
  printf("{\"probe\":\"%s\",\"warps_per_sm\":%d,\"imad_wide_per_clk_per_sm\":%.2f}\n", name, warps, inst / cavg);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  uint32_t *dout, *din; u64* dcyc;
  CK(cudaMalloc(&dout, 4096 * 4)); CK(cudaMalloc(&din, 4096 * 4)); CK(cudaMalloc(&dcyc, 1024 * 8));
  uint32_t h[4096]; for (int i = 0; i < 4096; i++) h[i] = 0x9e3779b9u * (i + 1) | 1;
  CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
  for (int w : {4, 8, 16}) run<1>("a_fixed_b_varies", w, nsm, dout, din, dcyc);
  for (int w : {4, 8, 16}) run<2>("a_b_vary", w, nsm, dout, din, dcyc);
  for (int w : {4, 8, 16}) run<3>("mac_block_asm_rowmajor", w, nsm, dout, din, dcyc);
  for (int w : {4, 8, 16}) run<4>("mac_block_c", w, nsm, dout, din, dcyc);
  for (int w : {4, 8, 16}) run<5>("mac_block_packed_pairs", w, nsm, dout, din, dcyc);
  return 0;
}
