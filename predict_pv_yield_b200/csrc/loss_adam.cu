// loss_adam.cu -- a10 (step losses) and a12 (Adam) of the Conv3d PV step.
//
// Loss reference: predict_pv_yield/models/base_model.py:95-103
//     y = yield[0:batch_size, -forecast_len:, 0]
//     mse_loss = F.mse_loss(y_hat, y); nmae_loss = (y_hat - y).abs().mean()        <- nmae is returned (:146)
//     mse_exp / mae_exp = nowcasting_utils WeightedLosses (mean(w * d^2), mean(w * |d|))
// Adam reference: base_model.py:255-257 -> torch.optim.Adam(lr=5e-4) single-tensor arithmetic.
#include "common.cuh"

namespace pvb {

// one CTA; B*FO is at most a few 10^5 elements
__global__ void __launch_bounds__(1024)
l1_loss_fwd_kernel(const float* __restrict__ y_hat, const float* __restrict__ y, long long y_sb, long long y_sf,
                   const float* __restrict__ wts, float* __restrict__ losses, int B, int FO) {
  __shared__ float red[4][32];
  const int n = B * FO;
  float s_abs = 0.f, s_sq = 0.f, s_wsq = 0.f, s_wabs = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int b = i / FO, f = i - b * FO;
    const float d = y_hat[i] - y[b * y_sb + f * y_sf];
    const float w = wts ? __ldg(wts + f) : 1.f;
    s_abs += fabsf(d);
    s_sq = fmaf(d, d, s_sq);
    s_wsq = fmaf(w * d, d, s_wsq);
    s_wabs = fmaf(w, fabsf(d), s_wabs);
  }
  s_abs = warp_sum(s_abs); s_sq = warp_sum(s_sq); s_wsq = warp_sum(s_wsq); s_wabs = warp_sum(s_wabs);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = s_abs; red[1][warp] = s_sq; red[2][warp] = s_wsq; red[3][warp] = s_wabs; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    float v0 = lane < nw ? red[0][lane] : 0.f, v1 = lane < nw ? red[1][lane] : 0.f;
    float v2 = lane < nw ? red[2][lane] : 0.f, v3 = lane < nw ? red[3][lane] : 0.f;
    v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3);
    if (lane == 0) {
      const float inv = 1.f / static_cast<float>(n);
      losses[0] = v0 * inv;  // nmae (L1) -- the loss that is back-propagated
      losses[1] = v1 * inv;  // mse
      losses[2] = v2 * inv;  // mse_exp
      losses[3] = v3 * inv;  // mae_exp
    }
  }
}

// g = gscale * sign(y_hat - y) / n     (autograd of abs().mean(): sign(0) = 0)
__global__ void l1_loss_bwd_kernel(const float* __restrict__ y_hat, const float* __restrict__ y, long long y_sb,
                                   long long y_sf, const float* __restrict__ gscale, float* __restrict__ g, int B, int FO) {
  const int n = B * FO;
  const float sc = (gscale ? __ldg(gscale) : 1.f) / static_cast<float>(n);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int b = i / FO, f = i - b * FO;
    const float d = y_hat[i] - y[b * y_sb + f * y_sf];
    g[i] = d > 0.f ? sc : (d < 0.f ? -sc : 0.f);
  }
}

// ---- multi-tensor Adam ------------------------------------------------------------------------------------
constexpr int kAdamMaxTensors = 48;
constexpr int kAdamChunk = 4096;  // elements per CTA work item (256 threads x 4 x float4)

struct AdamTensors {
  float* p[kAdamMaxTensors];
  const float* g[kAdamMaxTensors];
  float* m[kAdamMaxTensors];
  float* v[kAdamMaxTensors];
  long long chunk_start[kAdamMaxTensors + 1];  // prefix sum of per-tensor chunk counts
  long long numel[kAdamMaxTensors];
  int n;
};

struct AdamScalars {
  float beta1, beta2, one_minus_beta1, one_minus_beta2, eps, neg_step_size, bc2_sqrt, grad_scale;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamScalars& s) {
  g *= s.grad_scale;
  m = fmaf(g - m, s.one_minus_beta1, m);          // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(s.one_minus_beta2 * g, g, v * s.beta2);  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = __fdiv_rn(sqrtf(v), s.bc2_sqrt) + s.eps;  // (sqrt(v) / sqrt(bc2)).add_(eps)
  p = fmaf(s.neg_step_size, __fdiv_rn(m, denom), p);      // param.addcdiv_(exp_avg, denom, value=-step_size)
}

// MINB = CTAs per SM the register allocation aims at: 3 (80 registers) for the stand-alone launch, one CTA per chunk -- 0.64 ms
// for fc1.weight, 0.94 of HBM; with 2 per SM or a persistent grid it took 0.77 ms -- and 2 (86 registers) for the narrow
// side-stream launch, whose CTAs must leave room for the convolution CTAs of the next forward pass on their SMs
template <int MINB>
__global__ void __launch_bounds__(256, MINB) adam_multi_kernel(const AdamTensors t, const AdamScalars s, long long total_chunks) {
  for (long long ch = blockIdx.x; ch < total_chunks; ch += gridDim.x) {
    // locate tensor (n <= 48: linear scan over a kernel-parameter array)
    int ti = 0;
    while (ti + 1 < t.n && ch >= t.chunk_start[ti + 1]) ++ti;
    const long long off = (ch - t.chunk_start[ti]) * kAdamChunk;
    const long long n = t.numel[ti];
    float* __restrict__ p = t.p[ti];
    const float* __restrict__ g = t.g[ti];
    float* __restrict__ m = t.m[ti];
    float* __restrict__ v = t.v[ti];
    const bool vec = (off + kAdamChunk <= n) && (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                                                    reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0);
    if (vec) {
      float4 pv[4], gv[4], mv[4], vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long i = off + (u * 256 + threadIdx.x) * 4;
        // streaming hints both ways: nothing here is read again before 4 GB have passed through the 126 MB L2
        pv[u] = __ldcs(reinterpret_cast<const float4*>(p + i));
        gv[u] = __ldcs(reinterpret_cast<const float4*>(g + i));
        mv[u] = __ldcs(reinterpret_cast<const float4*>(m + i));
        vv[u] = __ldcs(reinterpret_cast<const float4*>(v + i));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long i = off + (u * 256 + threadIdx.x) * 4;
        adam_elem(pv[u].x, gv[u].x, mv[u].x, vv[u].x, s);
        adam_elem(pv[u].y, gv[u].y, mv[u].y, vv[u].y, s);
        adam_elem(pv[u].z, gv[u].z, mv[u].z, vv[u].z, s);
        adam_elem(pv[u].w, gv[u].w, mv[u].w, vv[u].w, s);
        __stcs(reinterpret_cast<float4*>(p + i), pv[u]);
        __stcs(reinterpret_cast<float4*>(m + i), mv[u]);
        __stcs(reinterpret_cast<float4*>(v + i), vv[u]);
      }
    } else {
      const long long end = min(n, off + kAdamChunk);
      for (long long i = off + threadIdx.x; i < end; i += 256) {
        float pp = p[i], mm = m[i], vv = v[i];
        adam_elem(pp, g[i], mm, vv, s);
        p[i] = pp; m[i] = mm; v[i] = vv;
      }
    }
  }
}


// ---- validation results (base_model.py:222-236): capacity-scaled MW values and per-horizon errors in one pass ----------
// out[0][b][f] = forecast MW = y_hat * capacity, out[1][b][f] = actual MW = y * capacity, out[2][b][f] = capacity;
// horizon[0][f] = mean_b (y_hat - y)^2, horizon[1][f] = mean_b |y_hat - y|  (the per-horizon metrics of base_model.py:121-136)
__global__ void validation_results_kernel(const float* __restrict__ y_hat, const float* __restrict__ y, long long y_sb, long long y_sf,
                                          const float* __restrict__ cap, long long c_sb, long long c_sf, float* __restrict__ out,
                                          float* __restrict__ horizon, int B, int FO) {
  const int f = blockIdx.x;
  float s_sq = 0.f, s_abs = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float p = y_hat[b * FO + f], t = y[b * y_sb + f * y_sf];
    const float c = cap ? cap[b * c_sb + f * c_sf] : 1.f;
    const long long o = static_cast<long long>(b) * FO + f, n = static_cast<long long>(B) * FO;
    out[o] = p * c;
    out[n + o] = t * c;
    out[2 * n + o] = c;
    const float d = p - t;
    s_sq = fmaf(d, d, s_sq);
    s_abs += fabsf(d);
  }
  __shared__ float red[2][32];
  s_sq = warp_sum(s_sq); s_abs = warp_sum(s_abs);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = s_sq; red[1][warp] = s_abs; }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    float a = lane < nw ? red[0][lane] : 0.f, c = lane < nw ? red[1][lane] : 0.f;
    a = warp_sum(a); c = warp_sum(c);
    if (lane == 0) { horizon[f] = a / B; horizon[FO + f] = c / B; }
  }
}

}  // namespace pvb

extern "C" {

int pvb200_l1_loss_fwd_f32(const float* y_hat, const float* y, long long y_sb, long long y_sf, const float* weights,
                           float* losses, int B, int FO, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(y_hat && y && losses, "l1_loss_fwd: null pointer");
  PVB_REQUIRE(B > 0 && FO > 0 && static_cast<long long>(B) * FO < (1LL << 30), "l1_loss_fwd: bad shape");
  l1_loss_fwd_kernel<<<1, 1024, 0, as_stream(stream)>>>(y_hat, y, y_sb, y_sf, weights, losses, B, FO);
  PVB_LAUNCHED("l1_loss_fwd");
  return PVB200_OK;
}

int pvb200_l1_loss_bwd_f32(const float* y_hat, const float* y, long long y_sb, long long y_sf, const float* gscale,
                           float* g, int B, int FO, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(y_hat && y && g, "l1_loss_bwd: null pointer");
  PVB_REQUIRE(B > 0 && FO > 0 && static_cast<long long>(B) * FO < (1LL << 30), "l1_loss_bwd: bad shape");
  const int n = B * FO;
  int grid = ceil_div(n, 256);
  if (grid > 1184) grid = 1184;
  l1_loss_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(y_hat, y, y_sb, y_sf, gscale, g, B, FO);
  PVB_LAUNCHED("l1_loss_bwd");
  return PVB200_OK;
}

int pvb200_adam_step_f32(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                         float* const* exp_avg_sq, const long long* numel, float lr, float beta1, float beta2,
                         float eps, int step, float grad_scale, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(n > 0 && params && grads && exp_avg && exp_avg_sq && numel, "adam_step: null argument");
  PVB_REQUIRE(step >= 1, "adam_step: step must be >= 1 (1-based count after increment)");
  const int sms = sm_count();
  PVB_REQUIRE(sms > 0, "adam_step: no CUDA device");
  // torch computes the bias corrections and step size in double on the host, then casts to the tensor dtype
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  AdamScalars s;
  s.beta1 = beta1; s.beta2 = beta2;
  s.one_minus_beta1 = static_cast<float>(1.0 - static_cast<double>(beta1));
  s.one_minus_beta2 = static_cast<float>(1.0 - static_cast<double>(beta2));
  s.eps = eps;
  s.neg_step_size = static_cast<float>(-(static_cast<double>(lr) / bc1));
  s.bc2_sqrt = static_cast<float>(sqrt(bc2));
  s.grad_scale = grad_scale;
  for (int base = 0; base < n; base += kAdamMaxTensors) {
    AdamTensors t;
    t.n = (n - base < kAdamMaxTensors) ? n - base : kAdamMaxTensors;
    long long chunks = 0;
    for (int i = 0; i < t.n; ++i) {
      PVB_REQUIRE(params[base + i] && grads[base + i] && exp_avg[base + i] && exp_avg_sq[base + i] && numel[base + i] > 0,
                  "adam_step: tensor %d is null or empty", base + i);
      t.p[i] = params[base + i]; t.g[i] = grads[base + i]; t.m[i] = exp_avg[base + i]; t.v[i] = exp_avg_sq[base + i];
      t.numel[i] = numel[base + i];
      t.chunk_start[i] = chunks;
      chunks += ceil_div(numel[base + i], (long long)kAdamChunk);
    }
    t.chunk_start[t.n] = chunks;
    // one CTA per chunk (measured 0.57 ms for fc1.weight against 0.63 ms with a persistent grid of 8 CTAs per SM) -- unless
    // SMs are reserved (pvb200_reserve_sms: the side-stream update under the next forward pass wants a NARROW grid)
    int dev = 0, dev_sms = sms;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms < dev_sms && chunks > static_cast<long long>(sms) * 8)
      adam_multi_kernel<2><<<static_cast<unsigned>(sms * 8), 256, 0, as_stream(stream)>>>(t, s, chunks);
    else
      adam_multi_kernel<3><<<static_cast<unsigned>(chunks), 256, 0, as_stream(stream)>>>(t, s, chunks);
    PVB_LAUNCHED("adam_multi");
  }
  return PVB200_OK;
}


int pvb200_validation_results_f32(const float* y_hat, const float* y, long long y_sb, long long y_sf, const float* capacity,
                                  long long c_sb, long long c_sf, float* out, float* horizon, int B, int FO,
                                  pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(y_hat && y && out && horizon && B > 0 && FO > 0, "validation_results: bad argument");
  validation_results_kernel<<<FO, 256, 0, as_stream(stream)>>>(y_hat, y, y_sb, y_sf, capacity, c_sb, c_sf, out, horizon, B, FO);
  PVB_LAUNCHED("validation_results");
  return PVB200_OK;
}

}  // extern "C"
