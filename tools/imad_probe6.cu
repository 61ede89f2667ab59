// imad_probe6.cu -- what makes an IMAD.WIDE take 4 or 5 cycles of the multiplier pipe on B200:
// register-bank parity of the two 32-bit sources, immediate / constant-bank multiplicands, carry chains.
//
// imad_probe5 showed 32.0 IMAD.WIDE/clk/SM with (a even, b odd) and 25.6 with (a even, b even), and
// 25.4 for an immediate multiplicand.  The pairing kernel's products read a[j] * b[i] with register
// parities that follow the word indices, and its Montgomery rows multiply by immediates, so the
// effective multiplier peak inside the kernel depends on these rules.  Every mode below runs 8
// independent accumulator chains per thread; the multiplicand of chain j is the LOW (even register)
// or HIGH (odd register) word of chain j+1's accumulator, so nothing can be hoisted; the fixed
// multiplier sits in the low or the high half of a 64-bit register pair (even / odd register).
// The SASS register parities are checked by tools/check_probe6_sass.py before the numbers are trusted.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/imad_probe6 tools/imad_probe6.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

constexpr int NCH = 8;
__constant__ uint32_t c_p[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                 0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};


// All modes are mad.lo.cc / madc.hi.cc carry pairs (SASS: IMAD.WIDE.U32.X, what the field code executes), eight
// chains per thread linked pairwise through the carry flag like the kernel's even / odd arrays.
// MODE 0: a even, b even    1: a even, b odd      2: a odd, b even    3: a odd, b odd   (as allocated by ptxas 12.9;
// check with: cuobjdump -sass build/imad_probe6 | grep IMAD.WIDE)
// MODE 4: a even, immediate 5: a odd, immediate   6: a even, constant bank   7: a odd, constant bank
// The halves are picked INSIDE the asm block (mov.b64 {lo, hi}, pair), so ptxas maps them onto the even / odd
// register of the aligned pair instead of choosing a parity of its own.
#define STEP(FIRST, ASEL, BEXPR)                                                                       \
  asm volatile("{ .reg .b32 alo, ahi, blo, bhi, dlo, dhi;\n\t"                                          \
               "mov.b64 {alo, ahi}, %1;\n\t mov.b64 {blo, bhi}, %2;\n\t mov.b64 {dlo, dhi}, %0;\n\t"   \
               FIRST ".lo.cc.u32 dlo, " ASEL ", " BEXPR ", dlo;\n\t"                                    \
               "madc.hi.cc.u32 dhi, " ASEL ", " BEXPR ", dhi;\n\t"                                      \
               "mov.b64 %0, {dlo, dhi}; }" : "+l"(d[j]) : "l"(d[jn]), "l"(bb))
#define STEPC(FIRST, ASEL)                                                                             \
  asm volatile("{ .reg .b32 alo, ahi, dlo, dhi;\n\t"                                                    \
               "mov.b64 {alo, ahi}, %1;\n\t mov.b64 {dlo, dhi}, %0;\n\t"                               \
               FIRST ".lo.cc.u32 dlo, " ASEL ", %2, dlo;\n\t"                                           \
               "madc.hi.cc.u32 dhi, " ASEL ", %2, dhi;\n\t"                                             \
               "mov.b64 %0, {dlo, dhi}; }" : "+l"(d[j]) : "l"(d[jn]), "r"(c_p[(u + j) % 12]))
template <int MODE>
__global__ void __launch_bounds__(1024) probe(uint32_t* out, const uint32_t* in, unsigned long long* cyc, int iters) {
  unsigned long long bb = ((unsigned long long)(in[32 + (threadIdx.x & 31)] | 1u) << 32) | (in[33 + (threadIdx.x & 31)] | 1u);
  unsigned long long d[NCH];
#pragma unroll
  for (int j = 0; j < NCH; j++) d[j] = ((unsigned long long)in[64 + j] << 20) + threadIdx.x;
  __syncthreads();
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    bb += d[3] | 0x100000001ull;                     // keeps the multiplier in its own aligned pair (not hoistable)
#pragma unroll
    for (int u = 0; u < 16; u++) {
#pragma unroll
      for (int j = 0; j < NCH; j++) {
        const int jn = (j + 1) % NCH;
        if (j < 2) {
          if (MODE == 0) STEP("mad", "alo", "bhi");
          if (MODE == 1) STEP("mad", "alo", "blo");
          if (MODE == 2) STEP("mad", "ahi", "bhi");
          if (MODE == 3) STEP("mad", "ahi", "blo");
          if (MODE == 4) STEP("mad", "alo", "0x397fe69a");
          if (MODE == 5) STEP("mad", "ahi", "0x397fe69a");
          if (MODE == 6) STEPC("mad", "alo");
          if (MODE == 7) STEPC("mad", "ahi");
        } else {
          if (MODE == 0) STEP("madc", "alo", "bhi");
          if (MODE == 1) STEP("madc", "alo", "blo");
          if (MODE == 2) STEP("madc", "ahi", "bhi");
          if (MODE == 3) STEP("madc", "ahi", "blo");
          if (MODE == 4) STEP("madc", "alo", "0x397fe69a");
          if (MODE == 5) STEP("madc", "ahi", "0x397fe69a");
          if (MODE == 6) STEPC("madc", "alo");
          if (MODE == 7) STEPC("madc", "ahi");
        }
      }
    }
  }
  unsigned long long t1 = clock64();
  unsigned long long s = bb;
#pragma unroll
  for (int j = 0; j < NCH; j++) s ^= d[j];
  if (s == 0x12345678ull) out[threadIdx.x] = (uint32_t)s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps, int nsm, uint32_t* dout, uint32_t* din, unsigned long long* dcyc) {
  const int iters = 2048;
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
  double mults = 16.0 * NCH * iters * threads;
  printf("{\"probe\":\"%s\",\"mode\":%d,\"warps_per_sm\":%d,\"imad_wide_per_clk_per_sm\":%.2f,\"chip_G_per_s\":%.1f,\"eff_sm_mhz\":%.0f}\n",
         name, MODE, warps, mults / cavg, mults * nsm / (ms * 1e-3) / 1e9, cavg / (ms * 1e-3) / 1e6);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  printf("{\"device\":\"%s\",\"sms\":%d}\n", prop.name, nsm);
  uint32_t *dout, *din; unsigned long long* dcyc;
  CK(cudaMalloc(&dout, 1024 * 4)); CK(cudaMalloc(&din, 4096 * 4)); CK(cudaMalloc(&dcyc, 1024 * 8));
  uint32_t h[4096]; for (int i = 0; i < 4096; i++) h[i] = 0x9e3779b9u * (i + 1) | 1;
  CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
  for (int w : {8, 16}) {
    run<0>("carry a_even b_even", w, nsm, dout, din, dcyc);
    run<1>("carry a_even b_odd", w, nsm, dout, din, dcyc);
    run<2>("carry a_odd b_even", w, nsm, dout, din, dcyc);
    run<3>("carry a_odd b_odd", w, nsm, dout, din, dcyc);
    run<4>("carry a_even imm", w, nsm, dout, din, dcyc);
    run<5>("carry a_odd imm", w, nsm, dout, din, dcyc);
    run<6>("carry a_even const", w, nsm, dout, din, dcyc);
    run<7>("carry a_odd const", w, nsm, dout, din, dcyc);
  }
  return 0;
}
