// hot_kernel.cu -- the fused pairing kernel alone (same body as k_pairing in kernels.cu), for quick SASS-level
// experiments on the hot primitives: compile in ~30 s, then
//   cuobjdump -sass build/hot_kernel.cubin | python tools/sass_parity.py      (register-bank conflicts)
//   nvdisasm -hex -c build/hot_kernel.cubin > build/hot.sass ; python tools/sass_sched.py build/hot.sass hot_pairing
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cubin -o build/hot_kernel.cubin tools/hot_kernel.cu
#include <cuda_runtime.h>
#include "../plonky2-bls12-381-pairing_b200/csrc/programs.cuh"
using namespace b381;
constexpr int BLOCK = B381_BLOCK;
constexpr size_t GARENA_U4_PER_CTA = (size_t)(MAX_NSLOTS - NS) * GPS * BLOCK;

extern "C" __global__ void __launch_bounds__(BLOCK, 1)
hot_pairing(const uint32_t* g1, const uint32_t* g2, const uint8_t* inf, uint32_t* out, size_t n, int mode, u4* garena, int* err, uint32_t* dump, uint32_t tmem_base) {
  extern __shared__ u4 smem[];
  Ctx cx;
  cx.sm = smem + threadIdx.x;
  cx.gm = garena + (size_t)blockIdx.x * GARENA_U4_PER_CTA + threadIdx.x;
  cx.sync = 1;
  const uint32_t warp = threadIdx.x >> 5;
  cx.tm = tmem_base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 256u;
  cx.nt = NT_MAX;
  for (size_t base = (size_t)blockIdx.x * BLOCK; base < n; base += (size_t)gridDim.x * BLOCK) {
    __syncthreads();
    size_t i = base + threadIdx.x;
    const bool active = i < n;
    if (!active) i = n - 1;
    int e = prog_pairing(cx, g1 + 24 * i, g2 + 48 * i, inf ? inf[i] : 0, active ? out + 144 * i : dump + 144 * threadIdx.x, mode);
    if (active && e) atomicOr(err, e);
  }
}
