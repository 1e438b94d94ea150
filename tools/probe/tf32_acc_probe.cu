// tf32_acc_probe.cu -- numeric semantics of tcgen05.mma.kind::tf32 that decide whether the fp32-mode convolutions can move
// to the tensor cores as 3xTF32 (tools/tf32x3_study.py):
//   (1) are 32-bit operands TRUNCATED or rounded to TF32 (10 mantissa bits)?
//   (2) how is the fp32 accumulator in TMEM rounded when an MMA adds to it (nearest / toward zero / other)?
//   (3) are the K = 8 products of one MMA summed exactly before they meet the accumulator?
//   (4) on random data (K = 864 = 108 accumulating MMAs): which host model reproduces the result bit for bit?
// Written at the end of round 1 WITHOUT a GPU at hand: it compiles (nvcc -gencode arch=compute_100a,code=sm_100a -O3
// -o tf32_acc_probe tf32_acc_probe.cu) but has not been run yet.  One CTA, M = 128, N = 16, K-major SWIZZLE_NONE operands.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "../../predict_pv_yield_b200/csrc/tc_common.cuh"
using namespace pvb;

constexpr int kM = 128, kN = 16, kK = 8;            // one MMA
constexpr int kABytes = kM * kK * 4, kBBytes = kN * kK * 4;
constexpr int kChunk = 32;                           // MMAs staged in shared memory at a time

// element (row, k) of a K-major SWIZZLE_NONE operand tile with 32-bit elements: core matrices of 8 rows x 16 bytes
// (4 elements); the two K core matrices of a row group are adjacent (LBO = 128), row groups 256 bytes apart (SBO = 256)
__host__ __device__ inline int tile_off(int row, int k) { return (row / 8) * 256 + (k / 4) * 128 + (row % 8) * 16 + (k % 4) * 4; }

// a: [nmma][kABytes], b: [nmma][kBBytes] operand tiles in the layout above; d: [128][16] fp32 result
__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int nmma,
                                                    float* __restrict__ d) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  // the operands go through shared memory in chunks of kChunk MMAs (108 MMAs of case (4) do not fit at once); the
  // accumulator stays in TMEM across the chunks
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + kChunk * kABytes;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc(&tptr, 32);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tptr;
  uint32_t phase = 0;
  for (int c0 = 0; c0 < nmma; c0 += kChunk) {
    const int nc = nmma - c0 < kChunk ? nmma - c0 : kChunk;
    for (int i = threadIdx.x; i < nc * kABytes / 4; i += blockDim.x)
      reinterpret_cast<uint32_t*>(a_s)[i] = reinterpret_cast<const uint32_t*>(a + static_cast<size_t>(c0) * kABytes)[i];
    for (int i = threadIdx.x; i < nc * kBBytes / 4; i += blockDim.x)
      reinterpret_cast<uint32_t*>(b_s)[i] = reinterpret_cast<const uint32_t*>(b + static_cast<size_t>(c0) * kBBytes)[i];
    tc::fence_proxy_async();  // generic-proxy writes of the operands -> visible to the tensor core's async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
      tc::tc_fence_after();
      const uint32_t idesc = tc::umma_idesc(kM, kN, /*fmt TF32=*/2, 0, 0);
      for (int i = 0; i < nc; ++i) {
        const uint64_t ad = tc::umma_desc(tc::smem_u32(a_s + i * kABytes), 128, 256);
        const uint64_t bd = tc::umma_desc(tc::smem_u32(b_s + i * kBBytes), 128, 256);
        tc::umma_tf32(tmem, ad, bd, idesc, (c0 + i) > 0 ? 1u : 0u);
      }
      tc::umma_commit(&bar);
    }
    tc::mbar_wait(&bar, phase);  // the chunk's MMAs have read their operands: shared memory may be overwritten
    phase ^= 1u;
    __syncthreads();
  }
  tc::tc_fence_after();
  uint32_t v[16];
  tc::tmem_ld_32x16(tmem + (static_cast<uint32_t>(warp * 32) << 16), v);  // warp w reads TMEM lanes 32w .. 32w+31
  tc::tmem_ld_wait();
  for (int j = 0; j < 16; ++j) d[(warp * 32 + (threadIdx.x & 31)) * kN + j] = __uint_as_float(v[j]);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 32);
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float f32_rz(double x) {
  float y = static_cast<float>(x);
  if (std::fabs(static_cast<double>(y)) > std::fabs(x)) y = std::nextafterf(y, 0.f);
  return y;
}

struct Run {
  int nmma;
  std::vector<uint8_t> a, b;
  std::vector<float> d;
  explicit Run(int n) : nmma(n), a(static_cast<size_t>(n) * kABytes, 0), b(static_cast<size_t>(n) * kBBytes, 0), d(kM * kN, 0.f) {}
  void setA(int i, int row, int k, float v) { memcpy(&a[static_cast<size_t>(i) * kABytes + tile_off(row, k)], &v, 4); }
  void setB(int i, int n, int k, float v) { memcpy(&b[static_cast<size_t>(i) * kBBytes + tile_off(n, k)], &v, 4); }
  float getA(int i, int row, int k) const { float v; memcpy(&v, &a[static_cast<size_t>(i) * kABytes + tile_off(row, k)], 4); return v; }
  float getB(int i, int n, int k) const { float v; memcpy(&v, &b[static_cast<size_t>(i) * kBBytes + tile_off(n, k)], 4); return v; }
  bool launch() {
    uint8_t *da, *db;
    float* dd;
    cudaMalloc(&da, a.size()); cudaMalloc(&db, b.size()); cudaMalloc(&dd, d.size() * 4);
    cudaMemcpy(da, a.data(), a.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size(), cudaMemcpyHostToDevice);
    const size_t smem = static_cast<size_t>(kChunk) * (kABytes + kBBytes);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    probe_kernel<<<1, 128, smem>>>(da, db, nmma, dd);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return false; }
    cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
    cudaFree(da); cudaFree(db); cudaFree(dd);
    return true;
  }
};

int main() {
  const float u23 = std::ldexp(1.f, -23);
  // ---- (0) sanity: D[r][n] = sum_k A[r][k] * B[n][k] with small integers --------------------------------------------
  {
    Run r(1);
    for (int row = 0; row < kM; ++row) for (int k = 0; k < kK; ++k) r.setA(0, row, k, static_cast<float>((row + k) % 5));
    for (int n = 0; n < kN; ++n) for (int k = 0; k < kK; ++k) r.setB(0, n, k, static_cast<float>((n * 3 + k) % 4));
    if (!r.launch()) return 1;
    int bad = 0;
    for (int row = 0; row < kM; ++row) for (int n = 0; n < kN; ++n) {
      float ref = 0;
      for (int k = 0; k < kK; ++k) ref += r.getA(0, row, k) * r.getB(0, n, k);
      bad += (r.d[row * kN + n] != ref);
    }
    printf("(0) layout sanity: %d of %d outputs differ from the integer reference%s\n", bad, kM * kN, bad ? "  <-- FIX THE PROBE FIRST" : "");
    if (bad) return 1;
  }
  // ---- (1)-(3): hand-made cases, one per output row, column 0 (B[0][k] = 1) --------------------------------------------
  {
    Run r(2);
    for (int i = 0; i < 2; ++i) for (int k = 0; k < kK; ++k) r.setB(i, 0, k, 1.f);
    r.setA(0, 0, 0, 1.f + std::ldexp(1.f, -11) + std::ldexp(1.f, -12));                     // row 0: operand rounding
    r.setA(0, 1, 0, 1.f);  r.setA(1, 1, 0, 1.5f * std::ldexp(1.f, -24));                     // row 1: + 0.75 ulp
    r.setA(0, 2, 0, -1.f); r.setA(1, 2, 0, -1.5f * std::ldexp(1.f, -24));                    // row 2: the same, negative
    r.setA(0, 3, 0, 1.f);  r.setA(1, 3, 0, 0.5f * std::ldexp(1.f, -24));                     // row 3: + 0.25 ulp
    r.setA(0, 4, 0, 1.f);  for (int k = 1; k < kK; ++k) r.setA(0, 4, k, std::ldexp(1.f, -25)); // row 4: 1 + 7 x 2^-25 in ONE MMA
    r.setA(0, 5, 0, 1.f);  for (int k = 0; k < kK; ++k) r.setA(1, 5, k, std::ldexp(1.f, -25)); // row 5: 1, then + 8 x 2^-25
    if (!r.launch()) return 1;
    auto ulps = [&](float v, float base) { return (v - base) / u23; };
    printf("(1) operand 1 + 2^-11 + 2^-12 read as %.10f  -> %s\n", r.d[0 * kN],
           r.d[0 * kN] == 1.f ? "TRUNCATED to TF32" : (r.d[0 * kN] == 1.f + std::ldexp(1.f, -10) ? "rounded to nearest" : "other"));
    printf("(2) 1 + 0.75 ulp across two MMAs = 1 %+g ulp ; -1 - 0.75 ulp = -1 %+g ulp ; 1 + 0.25 ulp = 1 %+g ulp\n",
           ulps(r.d[1 * kN], 1.f), ulps(r.d[2 * kN], -1.f), ulps(r.d[3 * kN], 1.f));
    printf("    -> nearest: +1 / -1 / +0 ; toward zero: +0 / -0 / +0 ; toward -inf: +0 / -1 / +0 ; toward +inf: +1 / -0 / +1\n");
    printf("(3) 1 + 7 x 2^-25 inside one MMA = 1 %+g ulp (exact sum 1.75 ulp: +2 nearest, +1 truncated, +0 = product-by-product truncation)\n",
           ulps(r.d[4 * kN], 1.f));
    printf("    1, then an MMA adding 8 x 2^-25 = 1 %+g ulp (+2 = the MMA's products are summed before they meet the accumulator)\n",
           ulps(r.d[5 * kN], 1.f));
  }
  // ---- (4) random TF32-exact data, K = 864: which accumulation model is bit-exact? ------------------------------------
  {
    const int nm = 108;
    Run r(nm);
    srand(518);
    auto rnd = [] { return static_cast<float>(rand()) / RAND_MAX * 2.f - 1.f; };
    for (int i = 0; i < nm; ++i) {
      for (int row = 0; row < kM; ++row) for (int k = 0; k < kK; ++k) r.setA(i, row, k, trunc_tf32(std::fmax(rnd(), 0.f)));
      for (int n = 0; n < kN; ++n) for (int k = 0; k < kK; ++k) r.setB(i, n, k, trunc_tf32(rnd() / 29.4f));
    }
    if (!r.launch()) return 1;
    int eq_rn = 0, eq_rz = 0;
    double err_max = 0, ref_max = 0;
    for (int row = 0; row < kM; ++row) for (int n = 0; n < kN; ++n) {
      float acc_rn = 0.f, acc_rz = 0.f;
      double exact = 0;
      for (int i = 0; i < nm; ++i) {
        double blk = 0;
        for (int k = 0; k < kK; ++k) blk += static_cast<double>(r.getA(i, row, k)) * r.getB(i, n, k);
        exact += blk;
        acc_rn = static_cast<float>(static_cast<double>(acc_rn) + blk);
        acc_rz = f32_rz(static_cast<double>(acc_rz) + blk);
      }
      const float got = r.d[row * kN + n];
      eq_rn += (got == acc_rn);
      eq_rz += (got == acc_rz);
      err_max = std::fmax(err_max, std::fabs(got - exact));
      ref_max = std::fmax(ref_max, std::fabs(exact));
    }
    printf("(4) K = 864 random: %d / %d outputs bit-equal to the round-to-nearest model, %d / %d to the truncating model; "
           "normalised error against float64 %.2e\n", eq_rn, kM * kN, eq_rz, kM * kN, err_max / ref_max);
  }
  return 0;
}
