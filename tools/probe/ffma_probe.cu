// ffma_probe.cu -- FP32 FMA pipe ceiling: scalar FFMA vs packed FFMA2 (fma.rn.f32x2, new on sm_100).
#include <cstdio>
#include <cuda_runtime.h>
template <bool PACKED>
__global__ void __launch_bounds__(256) probe(float* sink, int iters) {
  float2 c[8];
  for (int i = 0; i < 8; ++i) c[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
  const float2 a = make_float2(0.999f, 1.001f), b = make_float2(0.001f, -0.001f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (PACKED) c[i] = __ffma2_rn(c[i], a, b);
      else { c[i].x = fmaf(c[i].x, a.x, b.x); c[i].y = fmaf(c[i].y, a.y, b.y); }
    }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += c[i].x + c[i].y;
  if (s == 123.456f) sink[0] = s;
}
int main() {
  float* sink; cudaMalloc(&sink, 16);
  for (int packed = 0; packed < 2; ++packed) {
    for (int blocks_per_sm : {2, 4, 8}) {
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      const int iters = 8192, grid = 148 * blocks_per_sm;
      float best = 1e9f;
      for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        if (packed) probe<true><<<grid, 256>>>(sink, iters); else probe<false><<<grid, 256>>>(sink, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
      }
      const double flops = 2.0 * 16 * iters * 256.0 * grid;
      printf("%s  %d CTAs/SM: %.1f TFLOP/s\n", packed ? "FFMA2" : "FFMA ", blocks_per_sm, flops / best / 1e9);
    }
  }
  return 0;
}
