// conv3d_wgrad_f32.cu -- a11: Conv3d weight + bias gradient in fp32 (autograd of model.py:117-120).
//
//   dw[co,ci,kt,kh,kw] = sum_{b,t,h,w} gz[b,co,t,h,w] * x[b,ci,t+kt,h+kh,w+kw]      db[co] = sum gz
//
// A GEMM with M = Cout, N = Cin*27 and a reduction over B*To*Ho*Wo (1.1-2.1 M positions at B=32), so
// the whole [Cout x Cin*27] result lives in REGISTERS of one CTA (each thread: 4 output channels x
// 27 (or 9) taps for one input channel) while persistent CTAs stream disjoint position ranges through
// shared memory ("split-K"); a second tiny kernel sums the per-CTA partials in a fixed order
// (deterministic, no atomics).  FP32 FMA pipe bound: 432 FMAs per 22 LDS per thread iteration.
// Positions use the same flattened-pitch trick as the forward kernel (q = ho*Wps + wo): taps are plain
// offsets kh*Wps + kw into a contiguous staged window; gz is zero-filled at the wrap columns.
#include "common.cuh"

namespace pvb {

constexpr int kWgQC = 64;         // output positions staged per step
constexpr int kWgMaxCtas = 320;   // upper bound on persistent CTAs (workspace sizing)
constexpr int kWgMaxCi = 64;

struct WgradArgs {
  const void* x;      // [B,Ci,Ti,Hi,Wi] fp32 or int16
  const float* mean;
  const float* stdv;
  const float* gz;    // [B,Co,To,Ho,Wo]
  float* partial;     // [gridDim.x][Co*Ci*27 + Co]
  int B, Ci, Ti, Hi, Wi, Co, To, Ho, Wo;
  int Wps, NP, NPs;   // pitch, staged positions per plane, padded plane stride (NPs % 8 == 4)
  int tiles_per_plane;
  long long total_steps;  // B * tiles_per_plane * To
  int ncog;           // ceil(Co / 4)
  int items;          // ncog * Ci * KTS
  int pairs;          // Wi even and 8-byte aligned tensors: stage two positions per copy (never straddles a row)
  int pad_t;          // time padding of the layer (0 or 1): input plane to + kt - pad_t, zero outside [0, Ti)
  int pad_hw;         // spatial padding (0 or 1): staged position (hp, wp) is input (hp - pad_hw, wp - pad_hw), zero outside
};

// Work is ordered (b, tile, to) with `to` fastest: a CTA walks DOWN the time axis of one (sample, position tile)
// column, so consecutive steps share two of their three input planes.  Planes live in a 4-slot ring filled by
// cp.async (LDGSTS, zero-fill for out-of-range positions) one step ahead of the FMA loop.
template <bool kI16, int KTS>
__global__ void __launch_bounds__(KTS == 1 ? 256 : 384, KTS == 1 ? 1 : 2) conv3d_wgrad_f32_kernel(const WgradArgs a) {
  constexpr int NKT = 3 / KTS;  // kt values handled by one thread
  extern __shared__ __align__(16) float smem[];
  float* x_s = smem;                                        // [4 slots][Ci][NPs]
  float* gz_s = x_s + 4 * a.Ci * a.NPs;                     // [2][4*ncog][kWgQC]
  int* off_s = reinterpret_cast<int*>(gz_s + 2 * 4 * a.ncog * kWgQC);  // [NP] input-plane offsets
  int* goff_s = off_s + a.NP;                               // [kWgQC] gz-plane offsets

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  const int item = blockIdx.y * blockDim.x + tid;
  const bool active = item < a.items;
  int ci = 0, cog = 0, ktg = 0;
  if (active) {
    ci = item % a.Ci;
    const int r = item / a.Ci;
    cog = r % a.ncog;
    ktg = r / a.ncog;
  }

  float acc[NKT * 9][4];
#pragma unroll
  for (int t = 0; t < NKT * 9; ++t)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};

  const long long g_begin = a.total_steps * blockIdx.x / gridDim.x;
  const long long g_end = a.total_steps * (blockIdx.x + 1) / gridDim.x;
  const long long xplane = static_cast<long long>(a.Hi) * a.Wi;
  const long long gplane = static_cast<long long>(a.Ho) * a.Wo;

  // stage input time plane `ti` of sample b (all Ci channels) into ring slot `slot`
  auto stage_x = [&](int b, int ti, int slot) {
    ti -= a.pad_t;
    if (ti < 0 || ti >= a.Ti) {  // a plane of the time padding: zeros (block-uniform)
      for (int i = tid; i < a.Ci * a.NPs; i += blockDim.x) x_s[slot * a.Ci * a.NPs + i] = 0.f;
      return;
    }
    for (int c = warp; c < a.Ci; c += nwarp) {
      float* dst = x_s + (slot * a.Ci + c) * a.NPs;
      const long long base = ((static_cast<long long>(b) * a.Ci + c) * a.Ti + ti) * xplane;
      if (kI16) {
        const int16_t* src = static_cast<const int16_t*>(a.x) + base;
        const float m = __ldg(a.mean + c), s = __ldg(a.stdv + c);
        for (int i = lane; i < a.NP; i += 32) {
          const int o = off_s[i];
          dst[i] = (o >= 0) ? sat_norm(__ldg(src + o), m, s) : 0.f;
        }
      } else if (a.pairs) {
        const float* src = static_cast<const float*>(a.x) + base;
        const uint32_t d0 = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
        for (int i = 2 * lane; i < a.NP; i += 64) {
          const int o = off_s[i];
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d0 + 4u * i), "l"(src + (o >= 0 ? o : 0)),
                       "r"(o >= 0 ? 8 : 0)
                       : "memory");
        }
      } else {
        const float* src = static_cast<const float*>(a.x) + base;
        const uint32_t d0 = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
        for (int i = lane; i < a.NP; i += 32) {
          const int o = off_s[i];
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d0 + 4u * i), "l"(src + (o >= 0 ? o : 0)),
                       "r"(o >= 0 ? 4 : 0)
                       : "memory");
        }
      }
    }
  };
  // stage gz of output time `to` into buffer `buf`: [co][q], zero at the wrap columns / beyond the plane
  auto stage_gz = [&](int b, int to, int buf) {
    for (int p = warp; p < 4 * a.ncog; p += nwarp) {
      float* dst = gz_s + (buf * 4 * a.ncog + p) * kWgQC;
      const bool ok_p = p < a.Co;
      const float* src = a.gz + ((static_cast<long long>(b) * a.Co + (ok_p ? p : 0)) * a.To + to) * gplane;
      const uint32_t d0 = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
      if (a.pairs) {
        const int i = 2 * lane;  // kWgQC == 64
        const int o = goff_s[i];
        const bool ok = ok_p && (o >= 0);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d0 + 4u * i), "l"(src + (ok ? o : 0)),
                     "r"(ok ? 8 : 0)
                     : "memory");
      } else {
        for (int i = lane; i < kWgQC; i += 32) {
          const int o = goff_s[i];
          const bool ok = ok_p && (o >= 0);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d0 + 4u * i), "l"(src + (ok ? o : 0)),
                       "r"(ok ? 4 : 0)
                       : "memory");
        }
      }
    }
  };

  for (long long g = g_begin; g < g_end;) {
    // ---- a run: consecutive output times of one (b, tile) column ----
    const long long col = g / a.To;
    const int t0 = static_cast<int>(g - col * a.To);
    const int tile = static_cast<int>(col % a.tiles_per_plane);
    const int b = static_cast<int>(col / a.tiles_per_plane);
    const long long left = g_end - g;
    const int nstep = static_cast<int>(left < (a.To - t0) ? left : (a.To - t0));
    const int q0 = tile * kWgQC;

    __syncthreads();  // previous run fully consumed
    for (int i = tid; i < a.NP; i += blockDim.x) {
      const int pos = q0 + i;
      const int hp = pos / a.Wps, wp = pos - hp * a.Wps;
      const int hi = hp - a.pad_hw, wi = wp - a.pad_hw;
      off_s[i] = (hi >= 0 && hi < a.Hi && wi >= 0 && wi < a.Wi) ? hi * a.Wi + wi : -1;
    }
    for (int i = tid; i < kWgQC; i += blockDim.x) {
      const int pos = q0 + i;
      const int ho = pos / a.Wps, wo = pos - ho * a.Wps;
      goff_s[i] = (ho < a.Ho && wo < a.Wo) ? ho * a.Wo + wo : -1;
    }
    __syncthreads();
    // prologue: the three planes of the first step + its gz
    stage_x(b, t0, 0);
    stage_x(b, t0 + 1, 1);
    stage_x(b, t0 + 2, 2);
    stage_gz(b, t0, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");

    for (int s = 0; s < nstep; ++s) {
      if (s + 1 < nstep) {  // prefetch the one new plane of the next step, and its gz
        stage_x(b, t0 + s + 3, (s + 3) & 3);
        stage_gz(b, t0 + s + 1, (s + 1) & 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();

      if (active) {
        const float* gp = gz_s + ((s & 1) * 4 * a.ncog + 4 * cog) * kWgQC;
        const float* xk[NKT];
#pragma unroll
        for (int k = 0; k < NKT; ++k) {
          const int kt = (KTS == 1) ? k : ktg;
          xk[k] = x_s + (((s + kt) & 3) * a.Ci + ci) * a.NPs;
        }
#pragma unroll 1
        for (int q = 0; q < kWgQC; q += 4) {
          float gv[4][4];  // [co j][pos i]
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 v = *reinterpret_cast<const float4*>(gp + j * kWgQC + q);
            gv[j][0] = v.x; gv[j][1] = v.y; gv[j][2] = v.z; gv[j][3] = v.w;
            bacc[j] += (v.x + v.y) + (v.z + v.w);
          }
#pragma unroll
          for (int k = 0; k < NKT; ++k) {
            const float* xp = xk[k] + q;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
              const float4 v0 = *reinterpret_cast<const float4*>(xp + kh * a.Wps);
              const float2 v1 = *reinterpret_cast<const float2*>(xp + kh * a.Wps + 4);
              const float xv[6] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y};
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    acc[k * 9 + kh * 3 + kw][j] = fmaf(gv[j][i], xv[i + kw], acc[k * 9 + kh * 3 + kw][j]);
                }
              }
            }
          }
        }
      }
      __syncthreads();  // slot (s & 3) and gz buffer (s & 1) may be overwritten by the next prefetch
    }
    g += nstep;
  }

  // ---- write this CTA's partial ----
  if (active) {
    float* part = a.partial + static_cast<long long>(blockIdx.x) * (static_cast<long long>(a.Co) * a.Ci * 27 + a.Co);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = 4 * cog + j;
      if (co >= a.Co) continue;
#pragma unroll
      for (int k = 0; k < NKT; ++k) {
        const int kt = (KTS == 1) ? k : ktg;
#pragma unroll
        for (int t = 0; t < 9; ++t)
          part[(static_cast<long long>(co) * a.Ci + ci) * 27 + kt * 9 + t] = acc[k * 9 + t][j];
      }
      if (ci == 0 && ktg == 0) part[static_cast<long long>(a.Co) * a.Ci * 27 + co] = bacc[j];
    }
  }
}

// dw[i] = sum over CTAs (fixed order) ; the last Co entries are db
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, float* __restrict__ db,
                                    int n_w, int n_b, int n_part) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = n_w + n_b;
  if (i >= n) return;
  float s = 0.f;
  for (int p = 0; p < n_part; ++p) s += partial[static_cast<long long>(p) * n + i];
  if (i < n_w) dw[i] = s;
  else if (db) db[i - n_w] = s;
}

template <bool kI16, int KTS>
static int launch_wgrad(WgradArgs a, int grid_x, int grid_y, int threads, size_t smem, cudaStream_t stream) {
  PVB_CUDA(cudaFuncSetAttribute(conv3d_wgrad_f32_kernel<kI16, KTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  conv3d_wgrad_f32_kernel<kI16, KTS><<<dim3(grid_x, grid_y), threads, smem, stream>>>(a);
  PVB_LAUNCHED("conv3d_wgrad_f32");
  return PVB200_OK;
}

}  // namespace pvb

extern "C" {

size_t pvb200_conv3d_wgrad_workspace_bytes(int Cin, int Cout) {
  return static_cast<size_t>(pvb::kWgMaxCtas) * (static_cast<size_t>(Cout) * Cin * 27 + Cout) * sizeof(float);
}

int pvb200_conv3d_wgrad_f32(const void* x, int x_is_i16, const float* mean, const float* std, const float* gz,
                            float* dw, float* db, void* workspace, size_t workspace_bytes, int B, int Cin, int Ti,
                            int Hi, int Wi, int Cout, pvb200_stream_t stream) {
  return pvb200_conv3d_wgrad_f32_tpad(x, x_is_i16, mean, std, gz, dw, db, workspace, workspace_bytes, B, Cin, Ti, Hi, Wi, Cout, 0,
                                      stream);
}

int pvb200_conv3d_wgrad_f32_tpad(const void* x, int x_is_i16, const float* mean, const float* std, const float* gz,
                                 float* dw, float* db, void* workspace, size_t workspace_bytes, int B, int Cin, int Ti,
                                 int Hi, int Wi, int Cout, int pad_t, pvb200_stream_t stream) {
  return pvb200_conv3d_wgrad_f32_pad(x, x_is_i16, mean, std, gz, dw, db, workspace, workspace_bytes, B, Cin, Ti, Hi, Wi, Cout, pad_t,
                                     0, stream);
}

int pvb200_conv3d_wgrad_f32_pad(const void* x, int x_is_i16, const float* mean, const float* std, const float* gz,
                                float* dw, float* db, void* workspace, size_t workspace_bytes, int B, int Cin, int Ti,
                                int Hi, int Wi, int Cout, int pad_t, int pad_hw, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(x && gz && dw, "conv3d_wgrad: null pointer");
  PVB_REQUIRE((pad_t == 0 || pad_t == 1) && (pad_hw == 0 || pad_hw == 1), "conv3d_wgrad: padding (%d, %d) not in {0, 1}", pad_t, pad_hw);
  PVB_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && Ti + 2 * pad_t > 2 && Hi + 2 * pad_hw > 2 && Wi + 2 * pad_hw > 2, "conv3d_wgrad: bad shape");
  PVB_REQUIRE(Cin <= kWgMaxCi, "conv3d_wgrad: Cin=%d > %d not supported", Cin, kWgMaxCi);
  PVB_REQUIRE(!x_is_i16 || (mean && std), "conv3d_wgrad: int16 input needs mean/std");
  const size_t need = pvb200_conv3d_wgrad_workspace_bytes(Cin, Cout);
  if (!workspace || workspace_bytes < need) {
    set_error("conv3d_wgrad: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return PVB200_ERR_WORKSPACE;
  }
  WgradArgs a;
  a.x = x; a.mean = mean; a.stdv = std; a.gz = gz; a.partial = static_cast<float*>(workspace);
  a.B = B; a.Ci = Cin; a.Ti = Ti; a.Hi = Hi; a.Wi = Wi; a.Co = Cout;
  a.To = Ti + 2 * pad_t - 2; a.Ho = Hi + 2 * pad_hw - 2; a.Wo = Wi + 2 * pad_hw - 2;
  a.pad_t = pad_t; a.pad_hw = pad_hw;
  a.Wps = round_up(Wi + 2 * pad_hw, 4);
  a.NP = kWgQC + 2 * a.Wps + 8;
  a.NPs = round_up(a.NP, 8) + 4;
  const int Qtot = (a.Ho - 1) * a.Wps + a.Wo;
  a.tiles_per_plane = ceil_div(Qtot, kWgQC);
  a.total_steps = static_cast<long long>(B) * a.tiles_per_plane * a.To;
  a.ncog = ceil_div(Cout, 4);
  // narrow layers (conv0: Cin = 12) split the 27 taps over 3 threads to keep the CTA full
  const int kts = (a.ncog * Cin <= 128) ? 3 : 1;
  a.items = a.ncog * Cin * kts;
  a.pairs = (!x_is_i16 && pad_hw == 0 && Wi % 2 == 0 && reinterpret_cast<uintptr_t>(x) % 8 == 0 &&
             reinterpret_cast<uintptr_t>(gz) % 8 == 0) ? 1 : 0;
  const int cap = (kts == 1) ? 256 : 384;
  const int grid_y = ceil_div(a.items, cap);
  const int threads = round_up(ceil_div(a.items, grid_y), 32);
  const int sms = sm_count();
  PVB_REQUIRE(sms > 0, "conv3d_wgrad: no CUDA device");
  const size_t smem = (static_cast<size_t>(4) * Cin * a.NPs + 2 * 4 * a.ncog * kWgQC) * sizeof(float) +
                      (a.NP + kWgQC) * sizeof(int);
  // persistent CTAs: one per SM for the wide layers (254 registers, ~120 KB of shared memory), two per SM for the
  // narrow ones (kts == 3: 36 accumulators per thread), which doubles the warps that hide the shared-memory latency
  long long gx = (kts == 3 && 2 * smem <= 220 * 1024) ? 2LL * sms : sms;
  if (gx > kWgMaxCtas) gx = kWgMaxCtas;
  if (gx > a.total_steps) gx = a.total_steps;
  PVB_REQUIRE(smem <= 227 * 1024, "conv3d_wgrad: Cin=%d, width %d needs %zu B of shared memory (> 227 KB)", Cin, Wi, smem);
  cudaStream_t st = as_stream(stream);
  int rc;
  if (x_is_i16)
    rc = (kts == 1) ? launch_wgrad<true, 1>(a, (int)gx, grid_y, threads, smem, st)
                    : launch_wgrad<true, 3>(a, (int)gx, grid_y, threads, smem, st);
  else
    rc = (kts == 1) ? launch_wgrad<false, 1>(a, (int)gx, grid_y, threads, smem, st)
                    : launch_wgrad<false, 3>(a, (int)gx, grid_y, threads, smem, st);
  if (rc != PVB200_OK) return rc;
  // when the item space is split over grid_y, every y-slice wrote disjoint entries of the same partial rows
  const int n_w = Cout * Cin * 27;
  wgrad_reduce_kernel<<<ceil_div(n_w + Cout, 256), 256, 0, st>>>(a.partial, dw, db, n_w, Cout, (int)gx);
  PVB_LAUNCHED("wgrad_reduce");
  return PVB200_OK;
}

}  // extern "C"
