// imad_probe5.cu -- integer-multiplier issue rates on B200 with operands ptxas cannot strength-reduce.
//
// The earlier probes (imad_probe.cu MODE 0/1/2 and the first k_imad_peak) multiplied loop-invariant
// operands: ptxas computed a*b once and turned the "multiply-accumulates" into IADD3 chains, so they
// measured the ALU pipe (64 adds/clk/SM), not the multiplier.  Here every multiplicand is the low
// word of another chain's accumulator (it changes every round), and the SASS is checked by
// tools/check_probe_sass.sh: N mad instructions in the loop body => N IMAD* in the SASS.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/imad_probe5 tools/imad_probe5.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

constexpr int NCH = 8;   // independent accumulator chains per thread

// MODE 0: IMAD.WIDE.U32 (64-bit accumulate)      d[j] += lo(d[j+1]) * b
// MODE 1: IMAD (32-bit lo)                        e[j] += e[j+1] * b
// MODE 2: IMAD.HI.U32                             e[j] += hi(e[j+1] * b)
// MODE 3: IMAD.WIDE.U32 + one independent IADD3 per multiply (co-issue check)
// MODE 4: IMAD.WIDE.U32 with immediate multiplicand
// MODE 5: IMAD.WIDE.U32, RZ addend (pure multiply) + IADD3/IADD3.X pair (what ptxas emits when it un-fuses)
template<int MODE>
__global__ void __launch_bounds__(1024) probe(uint32_t* out, const uint32_t* in, unsigned long long* cyc, int iters) {
  uint32_t b = in[32 + (threadIdx.x & 31)] | 1u;
  unsigned long long d[NCH];
  uint32_t e[NCH], x[NCH];
#pragma unroll
  for (int j = 0; j < NCH; j++) { d[j] = ((unsigned long long)in[64 + j] << 20) + threadIdx.x; e[j] = in[80 + j] ^ threadIdx.x; x[j] = in[96 + j] + threadIdx.x; }
  __syncthreads();
  unsigned long long t0 = clock64();
  for (int it = 0; it < (MODE >= 6 ? 0 : iters); it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int j = 0; j < NCH; j++) {
        const int jn = (j + 1) % NCH;
        if (MODE == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(d[j]) : "r"((uint32_t)d[jn]), "r"(b));
        if (MODE == 1) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(e[j]) : "r"(e[jn]), "r"(b));
        if (MODE == 2) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(e[j]) : "r"(e[jn]), "r"(b));
        if (MODE == 3) {
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(d[j]) : "r"((uint32_t)d[jn]), "r"(b));
          asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(x[jn]));
        }
        if (MODE == 4) asm volatile("mad.wide.u32 %0, %1, 0x0fffaaab, %0;" : "+l"(d[j]) : "r"((uint32_t)d[jn]));
        if (MODE == 5) {
          unsigned long long p;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"((uint32_t)d[jn]), "r"(b));
          asm volatile("add.u64 %0, %0, %1;" : "+l"(d[j]) : "l"(p));
        }
      }
    }
  }
  if (MODE == 6 || MODE == 7) {
    // an 8 x 8 block of distinct products per iteration, like a real MAC block: d[j] += a[u] * bb[j]
    uint32_t a[NCH], bb[NCH];
#pragma unroll
    for (int j = 0; j < NCH; j++) { a[j] = e[j]; bb[j] = x[j]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
#pragma unroll
        for (int j = 0; j < NCH; j++) {
          if (MODE == 6) d[j] += (unsigned long long)a[u] * (unsigned long long)bb[j];
          else d[(j + u) % NCH] += (unsigned long long)a[u] * (unsigned long long)bb[j];
        }
      }
#pragma unroll
      for (int j = 0; j < NCH; j++) a[j] ^= (uint32_t)it;
    }
#pragma unroll
    for (int j = 0; j < NCH; j++) e[j] = a[j];
  }
  unsigned long long t1 = clock64();
  unsigned long long s = 0;
#pragma unroll
  for (int j = 0; j < NCH; j++) s ^= d[j] ^ e[j] ^ x[j];
  if (s == 0x12345678ull) out[threadIdx.x] = (uint32_t)s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template<int MODE>
void run(const char* name, int warps, int nsm, uint32_t* dout, uint32_t* din, unsigned long long* dcyc) {
  const int iters = 4096;
  const int threads = warps * 32;
  probe<MODE><<<nsm, threads>>>(dout, din, dcyc, 64);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  probe<MODE><<<nsm, threads>>>(dout, din, dcyc, iters);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  static unsigned long long h[1024];
  CK(cudaMemcpy(h, dcyc, nsm * 8, cudaMemcpyDeviceToHost));
  double cavg = 0; for (int i = 0; i < nsm; i++) cavg += (double)h[i]; cavg /= nsm;
  double mults = 8.0 * NCH * iters * threads;          // multiply instructions per SM (thread level)
  printf("{\"probe\":\"%s\",\"warps_per_sm\":%d,\"thread_mul_per_clk_per_sm\":%.2f,\"chip_Gmul_per_s\":%.1f,\"eff_sm_mhz\":%.0f}\n",
         name, warps, mults / cavg, mults * nsm / (ms * 1e-3) / 1e9, cavg / (ms * 1e-3) / 1e6);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  printf("{\"device\":\"%s\",\"sms\":%d}\n", prop.name, nsm);
  uint32_t *dout, *din; unsigned long long* dcyc;
  CK(cudaMalloc(&dout, 1024 * 4)); CK(cudaMalloc(&din, 4096 * 4)); CK(cudaMalloc(&dcyc, 1024 * 8));
  uint32_t h[4096]; for (int i = 0; i < 4096; i++) h[i] = 0x9e3779b9u * (i + 1) | 1;
  CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
  for (int w : {4, 8, 16, 32}) run<0>("imad_wide_u32_acc64", w, nsm, dout, din, dcyc);
  for (int w : {4, 8, 16, 32}) run<1>("imad_lo", w, nsm, dout, din, dcyc);
  for (int w : {4, 8, 16, 32}) run<2>("imad_hi", w, nsm, dout, din, dcyc);
  for (int w : {8, 32}) run<3>("imad_wide_plus_iadd", w, nsm, dout, din, dcyc);
  for (int w : {8, 32}) run<4>("imad_wide_imm", w, nsm, dout, din, dcyc);
  for (int w : {8, 32}) run<5>("mul_wide_then_add64", w, nsm, dout, din, dcyc);
  for (int w : {4, 8, 16, 32}) run<6>("mac_block_8x8_c", w, nsm, dout, din, dcyc);
  for (int w : {8, 32}) run<7>("mac_block_8x8_c_skewed", w, nsm, dout, din, dcyc);
  return 0;
}
