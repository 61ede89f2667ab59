// f2mul_probe_sb.cu -- schoolbook Fp2 multiply with ONE live column accumulator (low register
// footprint), to test whether 4 warps per sub-partition at <=128 registers beat 2 warps at 255.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../plonky2-bls12-381-pairing_b200/csrc/tower.cuh"
using namespace b381;
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void f2_mul_sb(Fp& r0, Fp& r1, const Fp& a0, const Fp& a1, const Fp& b0, const Fp& b1) {
  {
    Fp na1;
    fp_neg(na1, a1); fp_norm(na1);
    Acc T;
    acc_zero(T); acc_mac(T, a0, b0); acc_mac(T, na1, b1);
    acc_redc(r0, T);
  }
  {
    Acc T;
    acc_zero(T); acc_mac(T, a0, b1); acc_mac(T, a1, b0);
    acc_redc(r1, T);
  }
}

template<int THREADS>
__global__ void __launch_bounds__(THREADS, 1) probe(uint32_t* out, const uint32_t* in, unsigned long long* cyc, int iters) {
  Fp a0, a1, b0, b1;
  for (int k = 0; k < NL; k++) {
    a0.l[k] = (in[k] + threadIdx.x) & MASK; a1.l[k] = (in[20 + k] ^ threadIdx.x) & MASK;
    b0.l[k] = (in[40 + k] + 3 * threadIdx.x) & MASK; b1.l[k] = (in[60 + k] + 7 * threadIdx.x) & MASK;
  }
  __syncthreads();
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    Fp r0, r1;
    f2_mul_sb(r0, r1, a0, a1, b0, b1);
    a0 = r0; a1 = r1;
  }
  unsigned long long t1 = clock64();
  uint32_t s = 0;
  for (int k = 0; k < NL; k++) s ^= a0.l[k] ^ a1.l[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template<int THREADS>
void run(int nsm, uint32_t* dout, uint32_t* din, unsigned long long* dcyc) {
  const int iters = 2000;
  probe<THREADS><<<nsm, THREADS>>>(dout, din, dcyc, 10); CK(cudaDeviceSynchronize());
  probe<THREADS><<<nsm, THREADS>>>(dout, din, dcyc, iters); CK(cudaDeviceSynchronize());
  static unsigned long long h[1024]; CK(cudaMemcpy(h, dcyc, nsm * 8, cudaMemcpyDeviceToHost));
  double cavg = 0; for (int i = 0; i < nsm; i++) cavg += (double)h[i]; cavg /= nsm;
  double per_op_slot = cavg / iters / (THREADS / 128.0);
  printf("{\"probe\":\"f2_mul_schoolbook_1acc\",\"threads_per_sm\":%d,\"cycles_per_op_per_warp_slot\":%.0f}\n", THREADS, per_op_slot);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); int nsm = prop.multiProcessorCount;
  uint32_t *dout, *din; unsigned long long* dcyc;
  CK(cudaMalloc(&dout, 148 * 1024 * 4)); CK(cudaMalloc(&din, 4096 * 4)); CK(cudaMalloc(&dcyc, 1024 * 8));
  uint32_t h[4096]; for (int i = 0; i < 4096; i++) h[i] = 0x9e3779b9u * (i + 1) | 1;
  CK(cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice));
  run<128>(nsm, dout, din, dcyc); run<256>(nsm, dout, din, dcyc); run<384>(nsm, dout, din, dcyc); run<512>(nsm, dout, din, dcyc);
  return 0;
}
