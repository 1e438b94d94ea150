// ffma2_probe.cu -- microbenchmark of the packed fp32 FMA (fma.rn.f32x2, SASS FFMA2) that the fp32 convolutions run on:
// how many independent accumulator pairs and how many warps per scheduler does the FMA pipe need to stay busy, and what
// do interleaved LDS.128 cost?  Motivation (profiles/ncu_fp32_r01f.txt): the warp-specialised weight gradient keeps its 8
// consumer warps (2 per scheduler) 98 % of the time inside a loop of 216 FFMA2 + 22 LDS, yet the FMA pipe is only 78 %
// busy; the direct kernel (4 warps per scheduler) reaches 79 %.  Written at the end of round 1 after the GPU budget was
// spent: it compiles, it has not been run.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/ffma2_probe tools/probe/ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

// NACC independent accumulator pairs per thread, each updated once per inner iteration (dependent distance = NACC FFMA2);
// BCAST: second operand is one scalar broadcast to both halves (the R.F32 operand form of the convolution loops);
// LDS_EVERY > 0: one LDS.128 per LDS_EVERY FFMA2 whose result feeds the multiplier (like the kernels' operand loads)
template <int NACC, bool BCAST, int LDS_EVERY>
__global__ void __launch_bounds__(1024) probe(float* sink, int iters, float a0, float s0) {
  __shared__ float4 tab[256];
  if (threadIdx.x < 256) tab[threadIdx.x] = make_float4(a0, a0 * 0.5f, a0 * 0.25f, a0 * 0.125f);
  __syncthreads();
  float2 acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
  float2 a = make_float2(a0, a0 * 0.5f);
  float s = s0;  // a run-time value: the multiplier must stay a REGISTER operand (R.F32 form), not an immediate
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (LDS_EVERY > 0 && i % LDS_EVERY == 0) {
        const float4 v = tab[(threadIdx.x + it + i) & 255];
        a = make_float2(v.x, v.y);
        s = v.z;
      }
      acc[i] = BCAST ? __ffma2_rn(a, make_float2(s, s), acc[i]) : __ffma2_rn(a, make_float2(s, 0.5f * s), acc[i]);
    }
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) r += acc[i].x + acc[i].y;
  if (r == 123.456f) sink[0] = r;
}

template <int NACC, bool BCAST, int LDS_EVERY>
static void run(const char* name, int sms, float* sink) {
  const int iters = 4096;
  for (int warps_per_sm : {4, 8, 12, 16, 32}) {
    const int threads = warps_per_sm * 32 > 1024 ? 1024 : warps_per_sm * 32;
    const int ctas_per_sm = warps_per_sm * 32 / threads;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      probe<NACC, BCAST, LDS_EVERY><<<sms * ctas_per_sm, threads>>>(sink, iters, 0.001f, 0.999f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    const double fma2 = static_cast<double>(iters) * NACC * warps_per_sm * 32 * sms;  // thread-level FFMA2
    printf("%-34s %2d warps/SM (%d per scheduler): %7.2f TFLOP/s\n", name, warps_per_sm, warps_per_sm / 4,
           fma2 * 4.0 / (best * 1e-3) / 1e12);
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* sink; cudaMalloc(&sink, 16);
  printf("FFMA2 probe on %d SMs (fp32 FMA pipe peak 128 FMA/clk/SM)\n", sms);
  run<2, true, 0>("2 chains, broadcast operand", sms, sink);
  run<4, true, 0>("4 chains, broadcast operand", sms, sink);
  run<8, true, 0>("8 chains, broadcast operand", sms, sink);
  run<16, true, 0>("16 chains, broadcast operand", sms, sink);
  run<32, true, 0>("32 chains, broadcast operand", sms, sink);
  run<16, false, 0>("16 chains, register-pair operand", sms, sink);
  run<32, false, 0>("32 chains, register-pair operand", sms, sink);
  run<32, true, 8>("32 chains, 1 LDS.128 per 8 FFMA2", sms, sink);
  run<32, true, 4>("32 chains, 1 LDS.128 per 4 FFMA2", sms, sink);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
  return 0;
}
