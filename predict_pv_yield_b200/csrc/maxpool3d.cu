// maxpool3d.cu -- MaxPool3d(kernel 3, stride (1, 2, 2), padding (1, 1, 1)) forward / backward in fp32.
//
// Reference: self.sat_maxpool = nn.MaxPool3d(3, stride=(1, 2, 2), padding=(1, 1, 1)) of Conv3dMaxPool,
// predict_pv_yield/models/perceiver/perceiver_conv3d_nwp_sat.py:42-57.
// HBM-bound elementwise work: one thread per output (forward) / per input (backward).  The forward records the
// arg-max as torch does (first maximum in (t, h, w) scan order; NaN propagates); the backward is a GATHER over the
// at most 3 x 2 x 2 windows that contain an input position, so it is deterministic and needs no atomics.
#include <math.h>

#include "common.cuh"

namespace pvb {

__global__ void __launch_bounds__(256) maxpool3d_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int* __restrict__ argmax,
                                                            long long total, int T, int H, int W, int Ho, int Wo) {
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int wo = static_cast<int>(idx % Wo);
    long long r = idx / Wo;
    const int ho = static_cast<int>(r % Ho); r /= Ho;
    const int t = static_cast<int>(r % T);
    const long long n = r / T;
    const float* xp = x + n * (static_cast<long long>(T) * H * W);
    float best = -INFINITY;
    int arg = -1;
    for (int kt = -1; kt <= 1; ++kt) {
      const int ti = t + kt;
      if (ti < 0 || ti >= T) continue;
      for (int kh = -1; kh <= 1; ++kh) {
        const int hi = 2 * ho + kh;
        if (hi < 0 || hi >= H) continue;
        for (int kw = -1; kw <= 1; ++kw) {
          const int wi = 2 * wo + kw;
          if (wi < 0 || wi >= W) continue;
          const int o = (ti * H + hi) * W + wi;
          const float v = xp[o];
          // aten's max_pool3d rule: the first element, then any strictly greater value or NaN replaces the running maximum
          if (arg < 0 || v > best || v != v) { best = v; arg = o; }
        }
      }
    }
    y[idx] = best;
    argmax[idx] = arg;
  }
}

__global__ void __launch_bounds__(256) maxpool3d_bwd_kernel(const float* __restrict__ gz, const int* __restrict__ argmax,
                                                            float* __restrict__ gx, long long total, int T, int H, int W, int Ho,
                                                            int Wo) {
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int wi = static_cast<int>(idx % W);
    long long r = idx / W;
    const int hi = static_cast<int>(r % H); r /= H;
    const int ti = static_cast<int>(r % T);
    const long long n = r / T;
    const int me = (ti * H + hi) * W + wi;
    const long long obase = n * (static_cast<long long>(T) * Ho * Wo);
    float s = 0.f;
    // outputs (t, ho, wo) whose window contains this input: |t - ti| <= 1, |2 ho - hi| <= 1, |2 wo - wi| <= 1
    for (int t = ti - 1; t <= ti + 1; ++t) {
      if (t < 0 || t >= T) continue;
      for (int ho = hi / 2; 2 * ho <= hi + 1 && ho < Ho; ++ho) {
        for (int wo = wi / 2; 2 * wo <= wi + 1 && wo < Wo; ++wo) {
          const long long o = obase + (static_cast<long long>(t) * Ho + ho) * Wo + wo;
          if (argmax[o] == me) s += gz[o];
        }
      }
    }
    gx[idx] = s;
  }
}

}  // namespace pvb

extern "C" {

int pvb200_maxpool3d_fwd_f32(const float* x, float* y, int* argmax, long long N, int T, int H, int W, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(x && y && argmax && N > 0 && T > 0 && H > 0 && W > 0, "maxpool3d_fwd: bad argument");
  PVB_REQUIRE(static_cast<long long>(T) * H * W < 0x7fffffffLL, "maxpool3d_fwd: plane too large for 32-bit arg-max indices");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = N * T * Ho * Wo;
  long long grid = ceil_div(total, 256LL);
  if (grid > 148LL * 32) grid = 148LL * 32;
  maxpool3d_fwd_kernel<<<static_cast<unsigned>(grid), 256, 0, as_stream(stream)>>>(x, y, argmax, total, T, H, W, Ho, Wo);
  PVB_LAUNCHED("maxpool3d_fwd");
  return PVB200_OK;
}

int pvb200_maxpool3d_bwd_f32(const float* gz, const int* argmax, float* gx, long long N, int T, int H, int W,
                             pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(gz && argmax && gx && N > 0 && T > 0 && H > 0 && W > 0, "maxpool3d_bwd: bad argument");
  const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  const long long total = N * T * H * W;
  long long grid = ceil_div(total, 256LL);
  if (grid > 148LL * 32) grid = 148LL * 32;
  maxpool3d_bwd_kernel<<<static_cast<unsigned>(grid), 256, 0, as_stream(stream)>>>(gz, argmax, gx, total, T, H, W, Ho, Wo);
  PVB_LAUNCHED("maxpool3d_bwd");
  return PVB200_OK;
}

}  // extern "C"
