// imad_probe.cu -- integer-multiply pipe probe for B200 (sm_100a).
// Measures sustained issue rate of IMAD / IMAD.HI / IMAD.WIDE / carry-chained IMAD.WIDE.X
// and IADD3 co-issue, as instructions per clock per SM (clock64) and per second (CUDA events).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/imad_probe tools/imad_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

constexpr int ITER = 2048;

// one carry row: 6 x (mad.lo.cc, madc.hi.cc) -> 6 IMAD.WIDE.U32.X, + 1 addc
#define ROW(acc, x, b) asm volatile( \
   "mad.lo.cc.u32 %0, %13, %19, %0;\n\t"  "madc.hi.cc.u32 %1, %13, %19, %1;\n\t" \
   "madc.lo.cc.u32 %2, %14, %19, %2;\n\t" "madc.hi.cc.u32 %3, %14, %19, %3;\n\t" \
   "madc.lo.cc.u32 %4, %15, %19, %4;\n\t" "madc.hi.cc.u32 %5, %15, %19, %5;\n\t" \
   "madc.lo.cc.u32 %6, %16, %19, %6;\n\t" "madc.hi.cc.u32 %7, %16, %19, %7;\n\t" \
   "madc.lo.cc.u32 %8, %17, %19, %8;\n\t" "madc.hi.cc.u32 %9, %17, %19, %9;\n\t" \
   "madc.lo.cc.u32 %10, %18, %19, %10;\n\t" "madc.hi.cc.u32 %11, %18, %19, %11;\n\t" \
   "addc.u32 %12, %12, 0;\n\t" \
   : "+r"(acc[0]),"+r"(acc[1]),"+r"(acc[2]),"+r"(acc[3]),"+r"(acc[4]),"+r"(acc[5]), \
     "+r"(acc[6]),"+r"(acc[7]),"+r"(acc[8]),"+r"(acc[9]),"+r"(acc[10]),"+r"(acc[11]),"+r"(acc[12]) \
   : "r"(x[0]),"r"(x[1]),"r"(x[2]),"r"(x[3]),"r"(x[4]),"r"(x[5]),"r"(b))

template<int MODE, int ILP>
__global__ void __launch_bounds__(1024) probe(uint32_t* out, const uint32_t* in, unsigned long long* cyc, int iters) {
  uint32_t a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  unsigned long long t0 = 0, t1 = 0;
  if (MODE == 0 || MODE == 1 || MODE == 4) {
    uint32_t d[ILP];
    #pragma unroll
    for (int j = 0; j < ILP; j++) d[j] = in[64 + j] + threadIdx.x;
    __syncthreads();
    t0 = clock64();
    for (int it = 0; it < iters; it++) {
      #pragma unroll
      for (int u = 0; u < 16; u++) {
        #pragma unroll
        for (int j = 0; j < ILP; j++) {
          if (MODE == 0) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(d[j]) : "r"(a), "r"(b));
          if (MODE == 1) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(d[j]) : "r"(a), "r"(b));
          if (MODE == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(d[j]) : "r"(a));
        }
      }
    }
    t1 = clock64();
    uint32_t s = 0;
    #pragma unroll
    for (int j = 0; j < ILP; j++) s ^= d[j];
    if (s == 0x12345678u) out[threadIdx.x] = s;
  } else if (MODE == 2) {
    unsigned long long d[ILP];
    #pragma unroll
    for (int j = 0; j < ILP; j++) d[j] = in[64 + j] + threadIdx.x;
    __syncthreads();
    t0 = clock64();
    for (int it = 0; it < iters; it++) {
      #pragma unroll
      for (int u = 0; u < 16; u++) {
        #pragma unroll
        for (int j = 0; j < ILP; j++)
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(d[j]) : "r"(a), "r"(b));
      }
    }
    t1 = clock64();
    unsigned long long s = 0;
    #pragma unroll
    for (int j = 0; j < ILP; j++) s ^= d[j];
    if (s == 0x12345678ull) out[threadIdx.x] = (uint32_t)s;
  } else {  // MODE 3: carry rows; MODE 5: carry rows + independent IADD3 1:1; MODE 6: rows + 2 adds per IMAD
    uint32_t acc[ILP][13];
    uint32_t x[6];
    uint32_t e[12];
    #pragma unroll
    for (int j = 0; j < ILP; j++)
      #pragma unroll
      for (int k = 0; k < 13; k++) acc[j][k] = in[64 + k] + threadIdx.x + j;
    #pragma unroll
    for (int k = 0; k < 6; k++) x[k] = in[80 + k] ^ threadIdx.x;
    #pragma unroll
    for (int k = 0; k < 12; k++) e[k] = in[90 + k] ^ threadIdx.x;
    __syncthreads();
    t0 = clock64();
    for (int it = 0; it < iters; it++) {
      #pragma unroll
      for (int u = 0; u < 2; u++) {
        #pragma unroll
        for (int j = 0; j < ILP; j++) {
          ROW(acc[j], x, b);
          if (MODE == 5) {
            #pragma unroll
            for (int k = 0; k < 6; k++) asm volatile("add.u32 %0, %0, %1;" : "+r"(e[k]) : "r"(a));
          }
          if (MODE == 6) {
            #pragma unroll
            for (int k = 0; k < 12; k++) asm volatile("add.u32 %0, %0, %1;" : "+r"(e[k]) : "r"(a));
          }
        }
      }
    }
    t1 = clock64();
    uint32_t s = 0;
    #pragma unroll
    for (int j = 0; j < ILP; j++)
      #pragma unroll
      for (int k = 0; k < 13; k++) s ^= acc[j][k];
    #pragma unroll
    for (int k = 0; k < 12; k++) s ^= e[k];
    if (s == 0x12345678u) out[threadIdx.x] = s;
  }
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

struct Res { double ipc_sm; double ginst_s; double mhz; };

template<int MODE, int ILP>
Res run(int warps_per_sm, int nsm, uint32_t* dout, uint32_t* din, unsigned long long* dcyc) {
  int threads = warps_per_sm * 32;   // one block per SM
  int blocks = nsm;
  double per_thread_per_iter;
  if (MODE == 0 || MODE == 1 || MODE == 2 || MODE == 4) per_thread_per_iter = 16.0 * ILP;
  else per_thread_per_iter = 2.0 * ILP * 6.0;   // IMAD.WIDE.X count only
  probe<MODE, ILP><<<blocks, threads>>>(dout, din, dcyc, 64);  // warm
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<MODE, ILP><<<blocks, threads>>>(dout, din, dcyc, ITER);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  static unsigned long long h[1024];
  CK(cudaMemcpy(h, dcyc, blocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  double cavg = 0; for (int i = 0; i < blocks; i++) cavg += (double)h[i]; cavg /= blocks;
  double inst_per_sm = per_thread_per_iter * ITER * threads;
  Res r;
  r.ipc_sm = inst_per_sm / cavg;
  r.ginst_s = inst_per_sm * blocks / (ms * 1e-3) / 1e9;
  r.mhz = cavg / (ms * 1e-3) / 1e6;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return r;
}

template<int MODE, int ILP>
void sweep(const char* name, int nsm, uint32_t* dout, uint32_t* din, unsigned long long* dcyc) {
  int ws[] = {4, 8, 16, 32};
  for (int w : ws) {
    Res r = run<MODE, ILP>(w, nsm, dout, din, dcyc);
    printf("{\"probe\":\"%s\",\"ilp\":%d,\"warps_per_sm\":%d,\"thread_inst_per_clk_per_sm\":%.2f,\"chip_Ginst_per_s\":%.1f,\"eff_sm_mhz\":%.0f}\n",
           name, ILP, w, r.ipc_sm, r.ginst_s, r.mhz);
  }
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int nsm = prop.multiProcessorCount;
  printf("{\"device\":\"%s\",\"sms\":%d,\"clock_khz\":%d}\n", prop.name, nsm, prop.clockRate);
  uint32_t *dout, *din; unsigned long long* dcyc;
  CK(cudaMalloc(&dout, 4096 * 4)); CK(cudaMalloc(&din, 4096 * 4)); CK(cudaMalloc(&dcyc, 1024 * 8));
  uint32_t h[4096]; for (int i = 0; i < 4096; i++) h[i] = 0x9e3779b9u * (i + 1) | 1;
  CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
  sweep<0,1>("imad_lo", nsm, dout, din, dcyc);
  sweep<0,4>("imad_lo", nsm, dout, din, dcyc);
  sweep<0,8>("imad_lo", nsm, dout, din, dcyc);
  sweep<1,8>("imad_hi", nsm, dout, din, dcyc);
  sweep<2,1>("imad_wide", nsm, dout, din, dcyc);
  sweep<2,4>("imad_wide", nsm, dout, din, dcyc);
  sweep<2,8>("imad_wide", nsm, dout, din, dcyc);
  sweep<3,1>("imad_wide_x_row", nsm, dout, din, dcyc);
  sweep<3,2>("imad_wide_x_row", nsm, dout, din, dcyc);
  sweep<3,4>("imad_wide_x_row", nsm, dout, din, dcyc);
  sweep<4,8>("iadd", nsm, dout, din, dcyc);
  sweep<5,2>("row_plus_1add_per_imad", nsm, dout, din, dcyc);
  sweep<5,4>("row_plus_1add_per_imad", nsm, dout, din, dcyc);
  sweep<6,2>("row_plus_2add_per_imad", nsm, dout, din, dcyc);
  sweep<6,4>("row_plus_2add_per_imad", nsm, dout, din, dcyc);
  return 0;
}
