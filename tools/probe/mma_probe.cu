// mma_probe.cu -- microbenchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128) for different shared-memory
// operand layouts (SWIZZLE_NONE / 32B / 64B / 128B, K-major) and N.  Decides the operand layout of the conv kernels.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../predict_pv_yield_b200/csrc/tc_common.cuh"
using namespace pvb;

template <bool kTf32>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (kTf32)
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc));
  else
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc));
}

// layout: 0 none, 1 = 128B_base32B, 2 = 128B, 4 = 64B, 6 = 32B
// kTf32: kind::tf32 (K = 8 elements of 4 bytes per MMA: the same 32 bytes / two 16-byte core-matrix columns per row as
// kind::f16 with K = 16, so the descriptors of the f16 configurations are reused unchanged) -- added at the end of round 1
// for the 3xTF32 sizing of DESIGN.md section 7, not yet run
template <bool kTf32>
__global__ void probe(int layout, int N, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, int iters, int nacc,
                      long long* out, int M = 128, int nissuers = 1, int always_overwrite = 0, int amaj = 0, int bmaj = 0, int a_step = 8, int b_step = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, nissuers); tc::fence_barrier_init(); }
  tc::fence_proxy_async();
  if (warp == 0) tc::tmem_alloc(&tptr, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tptr;
  if ((threadIdx.x & 31) == 0 && warp < nissuers) {
    const uint32_t idesc = tc::umma_idesc(M, N, kTf32 ? 2 : 1, amaj, bmaj);
    const uint32_t a_addr = tc::smem_u32(smem), b_addr = a_addr + 96 * 1024;
    uint64_t ad = tc::umma_desc(a_addr, a_lbo, a_sbo) | (static_cast<uint64_t>(layout) << 61);
    uint64_t bd = tc::umma_desc(b_addr, b_lbo, b_sbo) | (static_cast<uint64_t>(layout) << 61);
    long long t0 = clock64();
    const uint32_t tbase = tmem + warp * (512 / nissuers);
    const uint32_t accflag = always_overwrite ? 0u : 1u;
    const uint32_t nmask = static_cast<uint32_t>(nacc - 1);  // nacc is a power of two
    for (int i = 0; i < nacc; ++i) mma<kTf32>(tbase + i * N, ad, bd, idesc, 0);
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) mma<kTf32>(tbase + ((j & nmask) * N), ad + j * a_step, bd + j * b_step, idesc, accflag);
    }
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (warp == 0) out[blockIdx.x] = t1 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  cudaFuncSetAttribute(probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  struct Cfg { const char* name; int layout, N; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; int nacc, M, nissuers, ow, amaj, bmaj; int a_step = 8, b_step = 0; int tf32 = 0; };
  Cfg cfgs[] = {
      // name, layout, N, a_lbo, a_sbo, b_lbo, b_sbo, nacc, M, issuers, overwrite, amaj, bmaj
      {"K/K   none  N=32", 0, 32, 6144, 128, 512, 128, 1, 128, 1, 0, 0, 0},
      {"MN/MN none  N=32", 0, 32, 128, 4096, 128, 2048, 1, 128, 1, 0, 1, 1},
      {"MN/K  none  N=32", 0, 32, 128, 4096, 512, 128, 1, 128, 1, 0, 1, 0},
      {"K/MN  none  N=32", 0, 32, 6144, 128, 128, 2048, 1, 128, 1, 0, 0, 1},
      {"MN/MN sw64  N=32", 4, 32, 8192, 512, 8192, 512, 1, 128, 1, 0, 1, 1},
      {"MN/MN sw128 N=32", 2, 32, 8192, 1024, 8192, 1024, 1, 128, 1, 0, 1, 1},
      {"MN/MN sw128 N=64", 2, 64, 8192, 1024, 8192, 1024, 1, 128, 1, 0, 1, 1},
      {"MN/MN sw32  N=32", 6, 32, 8192, 256, 8192, 256, 1, 128, 1, 0, 1, 1},
      {"MN/MN none  N=128", 0, 128, 128, 4096, 128, 2048, 1, 128, 1, 0, 1, 1},
      {"MN/MN sw128 N=128", 2, 128, 8192, 1024, 8192, 1024, 1, 128, 1, 0, 1, 1},
      {"MN/MN sw128 N=256", 2, 256, 8192, 1024, 8192, 1024, 1, 128, 1, 0, 1, 1},
      {"K/K   none  N=96", 0, 96, 6144, 128, 1536, 128, 1, 128, 1, 0, 0, 0},
      {"K/K   none  N=128", 0, 128, 6144, 128, 2048, 128, 1, 128, 1, 0, 0, 0},
      {"K/K   none  N=256", 0, 256, 6144, 128, 4096, 128, 1, 128, 1, 0, 0, 0},
      {"K/K   none  N=32 x2 issuers", 0, 32, 6144, 128, 512, 128, 1, 128, 2, 0, 0, 0},
      {"K/K   none  N=32 x4 issuers", 0, 32, 6144, 128, 512, 128, 1, 128, 4, 0, 0, 0},
      {"K/K   none  N=32 4 acc", 0, 32, 6144, 128, 512, 128, 4, 128, 1, 0, 0, 0},
      {"K/K   none  N=96 4 acc", 0, 96, 6144, 128, 1536, 128, 4, 128, 1, 0, 0, 0},
      {"K/K   none  N=32 M=64", 0, 32, 6144, 128, 512, 128, 1, 64, 1, 0, 0, 0},
      {"K/K   none  N=96 M=64", 0, 96, 6144, 128, 1536, 128, 1, 64, 1, 0, 0, 0},
      {"MN/MN none  N=96", 0, 96, 128, 4096, 128, 2048, 1, 128, 1, 0, 1, 1},
      {"MN/MN none  N=96 4 acc", 0, 96, 128, 4096, 128, 2048, 4, 128, 1, 0, 1, 1},
      {"K/K none N=96 A fixed B fixed", 0, 96, 6144, 128, 1536, 128, 1, 128, 1, 0, 0, 0, 0, 0},
      {"K/K none N=96 A var(2KB) B fixed", 0, 96, 6144, 128, 1536, 128, 1, 128, 1, 0, 0, 0, 128, 0},
      {"K/K none N=96 A fixed B var(3KB)", 0, 96, 6144, 128, 1536, 128, 1, 128, 1, 0, 0, 0, 0, 192},
      {"K/K none N=96 A var B var", 0, 96, 6144, 128, 1536, 128, 1, 128, 1, 0, 0, 0, 128, 192},
      {"K/K none N=96 A var misaligned B var", 0, 96, 6144, 128, 1536, 128, 1, 128, 1, 0, 0, 0, 62, 192},
      {"K/K none N=32 A var B var", 0, 32, 6144, 128, 512, 128, 1, 128, 1, 0, 0, 0, 128, 64},
      {"K/K none N=32 A var B var x2 issuers", 0, 32, 6144, 128, 512, 128, 1, 128, 2, 0, 0, 0, 128, 64},
      {"K/K none N=128 A var B var", 0, 128, 6144, 128, 2048, 128, 1, 128, 1, 0, 0, 0, 128, 256},
      {"K/K none N=256 A var B var", 0, 256, 6144, 128, 4096, 128, 1, 128, 1, 0, 0, 0, 128, 512},
      // kind::tf32 (K = 8): the MAC/cycle column counts 16 per MMA row x column, so halve it for these lines
      {"tf32 K/K none N=32", 0, 32, 6144, 128, 512, 128, 1, 128, 1, 0, 0, 0, 128, 64, 1},
      {"tf32 K/K none N=96", 0, 96, 6144, 128, 1536, 128, 1, 128, 1, 0, 0, 0, 128, 192, 1},
      {"tf32 K/K none N=96 x2 issuers", 0, 96, 6144, 128, 1536, 128, 1, 128, 2, 0, 0, 0, 128, 192, 1},
      {"tf32 K/K none N=48", 0, 48, 6144, 128, 768, 128, 1, 128, 1, 0, 0, 0, 128, 96, 1},
      {"tf32 K/K none N=48 x2 issuers", 0, 48, 6144, 128, 768, 128, 1, 128, 2, 0, 0, 0, 128, 96, 1},
      {"tf32 K/K none N=48 x4 issuers", 0, 48, 6144, 128, 768, 128, 1, 128, 4, 0, 0, 0, 128, 96, 1},
      {"tf32 K/K none N=64", 0, 64, 6144, 128, 1024, 128, 1, 128, 1, 0, 0, 0, 128, 128, 1},
      {"tf32 K/K none N=64 x2 issuers", 0, 64, 6144, 128, 1024, 128, 1, 128, 2, 0, 0, 0, 128, 128, 1},
      {"tf32 MN/MN none N=32", 0, 32, 128, 4096, 128, 2048, 1, 128, 1, 0, 1, 1, 8, 0, 1},
      {"tf32 MN/MN none N=96", 0, 96, 128, 4096, 128, 2048, 1, 128, 1, 0, 1, 1, 8, 0, 1},
      {"tf32 K/K none N=128", 0, 128, 6144, 128, 2048, 128, 1, 128, 1, 0, 0, 0, 128, 256, 1},
      {"tf32 K/K none N=256", 0, 256, 6144, 128, 4096, 128, 1, 128, 1, 0, 0, 0, 128, 512, 1},
  };
  const int iters = 2000;
  for (auto& c : cfgs) {
    const int grid = 148;
    if (c.tf32)
      probe<true><<<grid, 128, 160 * 1024>>>(c.layout, c.N, c.a_lbo, c.a_sbo, c.b_lbo, c.b_sbo, iters, c.nacc, out, c.M, c.nissuers, c.ow, c.amaj, c.bmaj, c.a_step, c.b_step);
    else
      probe<false><<<grid, 128, 160 * 1024>>>(c.layout, c.N, c.a_lbo, c.a_sbo, c.b_lbo, c.b_sbo, iters, c.nacc, out, c.M, c.nissuers, c.ow, c.amaj, c.bmaj, c.a_step, c.b_step);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
    long long h[148];
    cudaMemcpy(h, out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    const double cyc = double(mx) / iters;
    printf("%-26s %7.1f cycles/iter  -> %6.0f MAC/cycle/SM (peak ~4096)\n", c.name, cyc, double(c.M) * c.N * 16 * c.nissuers / cyc);
  }
  return 0;
}
