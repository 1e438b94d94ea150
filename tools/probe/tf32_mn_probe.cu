// tf32_mn_probe.cu -- layout semantics of tcgen05.mma.kind::tf32 with MN-major SWIZZLE_NONE operands (the weight-gradient
// kernel reads x and gz that way): element (mn, k) of an operand at  (mn / 4) * SBO + k * 16 + (mn % 4) * 4  bytes.
// One CTA, M = 128, N = 32..96, K = 8; small integers, so the result is exact.  Cases: group stride (SBO) a multiple of
// 128 B or not, start address not 128-B aligned, B with SBO = 0 (all N groups read one zero core matrix).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "../../predict_pv_yield_b200/csrc/tc_common.cuh"
using namespace pvb;

__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* __restrict__ a, int a_bytes, const uint8_t* __restrict__ b, int b_bytes,
                                                    uint32_t a_sbo, uint32_t b_sbo, uint32_t a_off, int N, int nmma, int kstep_bytes,
                                                    float* __restrict__ d, int a_mn, int b_mn, uint32_t a_lbo, uint32_t b_lbo,
                                                    int a_kstep, int b_kstep) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + ((a_bytes + 1023) & ~1023);
  for (int i = threadIdx.x; i < a_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(a_s)[i] = reinterpret_cast<const uint32_t*>(a)[i];
  for (int i = threadIdx.x; i < b_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(b_s)[i] = reinterpret_cast<const uint32_t*>(b)[i];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  tc::fence_proxy_async();
  if (warp == 0) tc::tmem_alloc(&tptr, 128);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = tc::umma_idesc(128, N, /*TF32*/ 2, a_mn, b_mn);
    for (int i = 0; i < nmma; ++i) {
      const uint64_t ad = tc::umma_desc(tc::smem_u32(a_s) + a_off + i * a_kstep, a_lbo, a_sbo);
      const uint64_t bd = tc::umma_desc(tc::smem_u32(b_s) + i * b_kstep, b_lbo, b_sbo);
      tc::umma_tf32(tmem, ad, bd, idesc, i > 0 ? 1u : 0u);
    }
    tc::umma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tc::tmem_ld_32x16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tc::tmem_ld_wait();
    for (int j = 0; j < 16; ++j) d[(warp * 32 + (threadIdx.x & 31)) * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

// K-major operand (rows x K, 8-row groups `sbo` apart, the two 4-element K groups of a K = 8 step `lbo` apart, next K step
// 2 * lbo further): element (r, k) at (r / 8) * sbo + (k / 4) * lbo + (r % 8) * 16 + (k % 4) * 4
static int run_mixed(const char* name, int a_mn, int b_mn, int N, int nk) {
  const int M = 128;
  const uint32_t a_sbo = a_mn ? 1024 : 128, b_sbo = b_mn ? 1024 : 128;
  const uint32_t a_lbo = a_mn ? 128 : 128 * 16, b_lbo = b_mn ? 128 : (N / 8) * 128;  // K-major: all row groups of one K group first
  const int a_bytes = 64 * 1024, b_bytes = 64 * 1024;
  std::vector<uint8_t> a(a_bytes, 0), b(b_bytes, 0);
  auto A = [&](int m, int k) { return ((m * 7 + k * 3) % 5) - 2; };
  auto B = [&](int n, int k) { return ((n * 5 + k) % 7) - 3; };
  auto off = [&](int mn, int r, int k, uint32_t sbo, uint32_t lbo) {
    return mn ? (r / 4) * sbo + k * 16 + (r % 4) * 4 : (r / 8) * sbo + (k / 4) * lbo + (r % 8) * 16 + (k % 4) * 4;
  };
  for (int m = 0; m < M; ++m) for (int k = 0; k < nk * 8; ++k) { float v = static_cast<float>(A(m, k)); memcpy(&a[off(a_mn, m, k, a_sbo, a_lbo)], &v, 4); }
  for (int n = 0; n < N; ++n) for (int k = 0; k < nk * 8; ++k) { float v = static_cast<float>(B(n, k)); memcpy(&b[off(b_mn, n, k, b_sbo, b_lbo)], &v, 4); }
  uint8_t *da, *db; float* dd;
  cudaMalloc(&da, a_bytes); cudaMalloc(&db, b_bytes); cudaMalloc(&dd, M * N * 4);
  cudaMemcpy(da, a.data(), a_bytes, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), b_bytes, cudaMemcpyHostToDevice);
  cudaMemset(dd, 0xff, M * N * 4);
  const int smem = a_bytes + b_bytes + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(da, a_bytes, db, b_bytes, a_sbo, b_sbo, 0, N, nk, 128, dd, a_mn, b_mn, a_lbo, b_lbo,
                                 a_mn ? 128 : 2 * (int)a_lbo, b_mn ? 128 : 2 * (int)b_lbo);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return 1; }
  std::vector<float> d(M * N);
  cudaMemcpy(d.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
  int bad = 0, first = -1;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    float ref = 0; for (int k = 0; k < nk * 8; ++k) ref += A(m, k) * B(n, k);
    if (d[m * N + n] != ref) { if (first < 0) first = m * N + n; ++bad; }
  }
  printf("%-58s: %5d of %d outputs differ", name, bad, M * N);
  if (bad) printf("  (first at m=%d n=%d: got %g)", first / N, first % N, d[first]);
  printf("\n");
  cudaFree(da); cudaFree(db); cudaFree(dd);
  return 0;
}

static int run_case(const char* name, uint32_t a_sbo, uint32_t b_sbo, uint32_t a_off, int N, int nk) {
  // operands cover nk K-steps of 8: K index k lives at k * 16 bytes inside a group segment
  const int M = 128;
  const int a_bytes = 32 * a_sbo + nk * 128 + a_off + 256, b_bytes = (N / 4) * (b_sbo ? b_sbo : 0) + nk * 128 + 256;
  std::vector<uint8_t> a(a_bytes, 0), b(b_bytes, 0);
  auto A = [&](int m, int k) { return ((m * 7 + k * 3) % 5) - 2; };
  auto B = [&](int n, int k) { return b_sbo ? ((n * 5 + k) % 7) - 3 : 0; };
  for (int m = 0; m < M; ++m) for (int k = 0; k < nk * 8; ++k) { float v = static_cast<float>(A(m, k)); memcpy(&a[(m / 4) * a_sbo + k * 16 + (m % 4) * 4 + a_off], &v, 4); }
  if (b_sbo) for (int n = 0; n < N; ++n) for (int k = 0; k < nk * 8; ++k) { float v = static_cast<float>(B(n, k)); memcpy(&b[(n / 4) * b_sbo + k * 16 + (n % 4) * 4], &v, 4); }
  uint8_t *da, *db; float* dd;
  cudaMalloc(&da, a_bytes); cudaMalloc(&db, b_bytes); cudaMalloc(&dd, M * N * 4);
  cudaMemcpy(da, a.data(), a_bytes, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), b_bytes, cudaMemcpyHostToDevice);
  cudaMemset(dd, 0xff, M * N * 4);
  const int smem = ((a_bytes + 1023) & ~1023) + b_bytes + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(da, a_bytes, db, b_bytes, a_sbo, b_sbo, a_off, N, nk, 128, dd, 1, 1, 128, 128, 128, 128);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); return 1; }
  std::vector<float> d(M * N);
  cudaMemcpy(d.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
  int bad = 0, first = -1;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    float ref = 0; for (int k = 0; k < nk * 8; ++k) ref += A(m, k) * B(n, k);
    if (d[m * N + n] != ref) { if (first < 0) first = m * N + n; ++bad; }
  }
  printf("%-58s: %5d of %d outputs differ", name, bad, M * N);
  if (bad) printf("  (first at m=%d n=%d: got %g)", first / N, first % N, d[first]);
  printf("\n");
  cudaFree(da); cudaFree(db); cudaFree(dd);
  return 0;
}

int main() {
  run_mixed("A K-major, B K-major (control), N=32, K=2x8", 0, 0, 32, 2);
  run_mixed("A MN-major, B K-major, N=32, K=2x8", 1, 0, 32, 2);
  run_mixed("A K-major, B MN-major, N=32, K=2x8", 0, 1, 32, 2);
  run_mixed("A MN-major, B MN-major, N=32, K=2x8", 1, 1, 32, 2);
  run_case("SBO 1024 / 1024, N=32, K=8", 1024, 1024, 0, 32, 1);
  run_case("SBO 1024 / 1024, N=96, K=3x8", 1024, 1024, 0, 96, 3);
  run_case("SBO 160 (not a multiple of 128) / 128, N=96, K=8", 160, 128, 0, 96, 1);
  run_case("SBO 992 / 1024, N=96, K=8x8, A start +16 B", 992, 1024, 16, 96, 8);
  run_case("SBO 992 / 1024, N=96, K=8x8, A start +32 B", 992, 1024, 32, 96, 8);
  run_case("B SBO = 0 over zeros (accumulator reset), N=96", 1024, 0, 0, 96, 1);
  return 0;
}
