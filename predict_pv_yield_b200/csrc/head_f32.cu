// head_f32.cu -- a5-a9/a11: the fully connected head of the Conv3d PV model, forward and backward, fp32.
//
// Reference: predict_pv_yield/models/conv3d/model.py:122-154
//   out = reshape(B, cnn_output_size); relu(fc1); relu(fc2); cat(PV history); cat(relu(fc_nwp(nwp)));
//   relu(fc3); fc4; reshape(B, forecast_len)            (+ the autograd of those lines)
//
// fc1 is [F1=128] x [K1=1 103 872]: 565 MB of fp32 weights against a 32..256-row batch, i.e. a
// weight-streaming GEMM that is HBM-bound on W1 for per-GPU batches <~ 200.  Kernels:
//   fc1_fwd_splitk   : split-K over ~4 CTAs/SM, each streams a K-range of W1 and x through smem and
//                      writes a [B x F1] partial (deterministic: partials summed in order by the tail)
//   head_tail_fwd    : per-sample CTA: sum partials + bias + ReLU, fc2, concat, fc_nwp, fc3, fc4
//   head_tail_bwd    : per-sample CTA: back through fc4, fc3, concat split, fc2 (ReLU masks fused)
//   linear_wgrad_small: dW = Gz^T In, db = sum Gz for the small layers (fc2, fc3, fc4, fc_nwp, fc1 bias)
//   fc1_wgrad        : dW1 tile [128 j x 128 k] = G1^T X, written once (565 MB, HBM-write-bound)
//   fc1_dgrad        : gx = (G1 W1) * (x > 0): streams W1 once more, ReLU mask of the last conv fused
#include <float.h>

#include "common.cuh"

namespace pvb {

constexpr int kHeadThreads = 256;
constexpr int kFc1KT = 64;          // K columns per smem stage (forward)
constexpr int kFc1BT = 32;          // batch rows per tile
constexpr int kFc1JT = 128;         // fc1 output features per tile
constexpr int kFc1TargetCtas = 296; // K splits per feature tile: 2 CTAs per SM on 148 SMs at one batch tile

// ---- split-K planning shared by the workspace query and the launcher --------------------------------
struct Fc1Plan {
  int nbt;              // batch tiles
  int njt;              // feature tiles
  int S;                // K splits
  long long k_per_split;
};

static Fc1Plan fc1_plan(int B, int F1, long long K1) {
  Fc1Plan p;
  p.nbt = ceil_div(B, kFc1BT);
  p.njt = ceil_div(F1, kFc1JT);
  long long stages = ceil_div(K1, (long long)kFc1KT);
  // the K split does NOT depend on the batch size: a sample's forecast is bit-identical whatever batch it rides in
  long long S = kFc1TargetCtas / p.njt;
  if (S < 1) S = 1;
  if (S > stages) S = stages;
  p.k_per_split = ceil_div(stages, S) * kFc1KT;
  p.S = static_cast<int>(ceil_div(K1, p.k_per_split));
  return p;
}

// ---- fc1 forward, split-K -----------------------------------------------------------------------------
// partial[s][b][j] = sum_{k in split s} x[b][k] * w[j][k]
__global__ void __launch_bounds__(kHeadThreads)
fc1_fwd_splitk_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ partial, int B,
                      int F1, long long K1, int nbt, int njt, long long k_per_split, int vec_ok) {
  constexpr int LD = kFc1KT + 4;  // row stride 68 words: == 4 (mod 32) -> conflict-free LDS.128 across rows
  __shared__ __align__(16) float ws[kFc1JT * LD];
  __shared__ __align__(16) float xs[kFc1BT * LD];

  int id = blockIdx.x;
  const int bt = id % nbt; id /= nbt;
  const int jt = id % njt;
  const int s = id / njt;
  const int b0 = bt * kFc1BT, j0 = jt * kFc1JT;
  const long long kb = s * k_per_split;
  const long long ke = min(K1, kb + k_per_split);

  const int tid = threadIdx.x;
  const int jg = tid & 31;  // rows jg, jg+32, jg+64, jg+96 of the feature tile
  const int bg = tid >> 5;  // batch rows 4bg..4bg+3 (warp-uniform -> x reads broadcast)

  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

  for (long long k0 = kb; k0 < ke; k0 += kFc1KT) {
    __syncthreads();
    const bool full = vec_ok && (k0 + kFc1KT <= ke);
    if (full) {
      // W tile: 128 rows x 16 float4
      for (int idx = tid; idx < kFc1JT * (kFc1KT / 4); idx += kHeadThreads) {
        const int r = idx >> 4, c4 = idx & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j0 + r < F1) v = __ldcs(reinterpret_cast<const float4*>(w + static_cast<long long>(j0 + r) * K1 + k0) + c4);
        *reinterpret_cast<float4*>(ws + r * LD + 4 * c4) = v;
      }
      for (int idx = tid; idx < kFc1BT * (kFc1KT / 4); idx += kHeadThreads) {
        const int r = idx >> 4, c4 = idx & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b0 + r < B) v = __ldg(reinterpret_cast<const float4*>(x + static_cast<long long>(b0 + r) * K1 + k0) + c4);
        *reinterpret_cast<float4*>(xs + r * LD + 4 * c4) = v;
      }
    } else {
      for (int idx = tid; idx < kFc1JT * kFc1KT; idx += kHeadThreads) {
        const int r = idx / kFc1KT, c = idx % kFc1KT;
        float v = 0.f;
        if (j0 + r < F1 && k0 + c < ke) v = w[static_cast<long long>(j0 + r) * K1 + k0 + c];
        ws[r * LD + c] = v;
      }
      for (int idx = tid; idx < kFc1BT * kFc1KT; idx += kHeadThreads) {
        const int r = idx / kFc1KT, c = idx % kFc1KT;
        float v = 0.f;
        if (b0 + r < B && k0 + c < ke) v = x[static_cast<long long>(b0 + r) * K1 + k0 + c];
        xs[r * LD + c] = v;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int k4 = 0; k4 < kFc1KT; k4 += 4) {
      float4 wv[4], xv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) wv[r] = *reinterpret_cast<const float4*>(ws + (jg + 32 * r) * LD + k4);
#pragma unroll
      for (int c = 0; c < 4; ++c) xv[c] = *reinterpret_cast<const float4*>(xs + (4 * bg + c) * LD + k4);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          acc[r][c] = fmaf(wv[r].x, xv[c].x, acc[r][c]);
          acc[r][c] = fmaf(wv[r].y, xv[c].y, acc[r][c]);
          acc[r][c] = fmaf(wv[r].z, xv[c].z, acc[r][c]);
          acc[r][c] = fmaf(wv[r].w, xv[c].w, acc[r][c]);
        }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int j = j0 + jg + 32 * r;
    if (j >= F1) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int b = b0 + 4 * bg + c;
      if (b < B) partial[(static_cast<long long>(s) * B + b) * F1 + j] = acc[r][c];
    }
  }
}

// ---- fc1 forward, split-K, fast path (K1 % 4 == 0, 16-byte aligned) ---------------------------------------
// Same contract as fc1_fwd_splitk_kernel.  256 threads = 4 k-groups x 64 threads; a 64-thread group owns the
// whole 32 (batch) x 128 (feature) tile with 8 x 8 register tiles (W rows interleaved by 16 so that LDS.128 across
// lanes is conflict-free, x rows warp-half-uniform -> broadcast) and takes every 4th k4-step of a stage, so a W
// element is read from shared memory by 4 threads instead of 8 and smem bandwidth stops being the limit; stages of
// 64 K columns stream in through a 2-stage cp.async pipeline; the four groups are summed through smem at the end.
constexpr int kFc1V2Threads = 256;
__global__ void __launch_bounds__(kFc1V2Threads, 2)
fc1_fwd_splitk_v2_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ partial, int B, int F1,
                         long long K1, int nbt, int njt, long long k_per_split) {
  constexpr int LD = kFc1KT + 4;                       // 68 words
  constexpr int STAGE = (kFc1JT + kFc1BT) * LD;        // floats per stage
  extern __shared__ __align__(16) float sm2[];         // [2][STAGE]; reused as [4][32][128] for the final reduce

  int id = blockIdx.x;
  const int bt = id % nbt; id /= nbt;
  const int jt = id % njt;
  const int s = id / njt;
  const int b0 = bt * kFc1BT, j0 = jt * kFc1JT;
  const long long kb = s * k_per_split;
  const long long ke = min(K1, kb + k_per_split);
  const int nstage = static_cast<int>((ke - kb + kFc1KT - 1) / kFc1KT);

  const int tid = threadIdx.x;
  const int grp = tid >> 6;          // k-group 0..3
  const int t64 = tid & 63;
  const int jl = t64 & 15;           // rows jl + 16 r, r = 0..7
  const int bg = t64 >> 4;           // batch rows 8 bg .. 8 bg + 7

  float acc[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

  auto issue = [&](int st, int buf) {
    const long long k0 = kb + static_cast<long long>(st) * kFc1KT;
    float* ws = sm2 + buf * STAGE;
    float* xs = ws + kFc1JT * LD;
    const uint32_t ws0 = static_cast<uint32_t>(__cvta_generic_to_shared(ws));
    const uint32_t xs0 = static_cast<uint32_t>(__cvta_generic_to_shared(xs));
    for (int idx = tid; idx < kFc1JT * (kFc1KT / 4); idx += kFc1V2Threads) {
      const int r = idx >> 4, c4 = idx & 15;
      const bool ok = (j0 + r < F1) && (k0 + 4 * c4 < ke);
      const float* src = w + static_cast<long long>(ok ? j0 + r : 0) * K1 + (ok ? k0 + 4 * c4 : 0);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ws0 + 4u * (r * LD + 4 * c4)), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    for (int idx = tid; idx < kFc1BT * (kFc1KT / 4); idx += kFc1V2Threads) {
      const int r = idx >> 4, c4 = idx & 15;
      const bool ok = (b0 + r < B) && (k0 + 4 * c4 < ke);
      const float* src = x + static_cast<long long>(ok ? b0 + r : 0) * K1 + (ok ? k0 + 4 * c4 : 0);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(xs0 + 4u * (r * LD + 4 * c4)), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  issue(0, 0);
  for (int st = 0; st < nstage; ++st) {
    if (st + 1 < nstage) {
      issue(st + 1, (st + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* ws = sm2 + (st & 1) * STAGE;
    const float* xs = ws + kFc1JT * LD;
#pragma unroll
    for (int kk = 0; kk < kFc1KT / 16; ++kk) {
      const int k4 = 16 * kk + 4 * grp;  // this group's k4-step
      float4 wv[8], xv[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) wv[r] = *reinterpret_cast<const float4*>(ws + (jl + 16 * r) * LD + k4);
#pragma unroll
      for (int c = 0; c < 8; ++c) xv[c] = *reinterpret_cast<const float4*>(xs + (8 * bg + c) * LD + k4);
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          acc[r][c] = fmaf(wv[r].x, xv[c].x, acc[r][c]);
          acc[r][c] = fmaf(wv[r].y, xv[c].y, acc[r][c]);
          acc[r][c] = fmaf(wv[r].z, xv[c].z, acc[r][c]);
          acc[r][c] = fmaf(wv[r].w, xv[c].w, acc[r][c]);
        }
    }
    __syncthreads();
  }
  // ---- sum the four k-groups (fixed order) and write the partial ----
  float* red = sm2;  // [4][32 b][128 j]
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) red[(grp * kFc1BT + 8 * bg + c) * kFc1JT + jl + 16 * r] = acc[r][c];
  __syncthreads();
  for (int idx = tid; idx < kFc1BT * kFc1JT; idx += kFc1V2Threads) {
    const int b = idx >> 7, j = idx & 127;
    const float v = (red[idx] + red[kFc1BT * kFc1JT + idx]) + (red[2 * kFc1BT * kFc1JT + idx] + red[3 * kFc1BT * kFc1JT + idx]);
    if (b0 + b < B && j0 + j < F1) partial[(static_cast<long long>(s) * B + b0 + b) * F1 + j0 + j] = v;
  }
}

__device__ __forceinline__ float nan_to_num_f(float v) {  // torch.nan_to_num defaults (model.py:131)
  if (isnan(v)) return 0.f;
  if (isinf(v)) return v > 0.f ? FLT_MAX : -FLT_MAX;
  return v;
}

// warp-cooperative dot product of a weight row with a shared-memory vector
__device__ __forceinline__ float warp_dot(const float* __restrict__ wrow, const float* vec, int n, int lane) {
  float s = 0.f;
  for (int i = lane; i < n; i += 32) s = fmaf(__ldg(wrow + i), vec[i], s);
  return warp_sum(s);
}

// ---- tail forward: one CTA per sample -------------------------------------------------------------------
__global__ void __launch_bounds__(kHeadThreads) head_tail_fwd_kernel(const pvb200_head_t h, const float* __restrict__ partial, int S) {
  extern __shared__ float sm[];
  const int NCAT = h.F2 + h.NPV + (h.NNWP > 0 ? h.FNWP : 0);
  float* h1 = sm;                 // [F1]
  float* cat = h1 + h.F1;         // [NCAT]
  float* h3 = cat + NCAT;         // [F3]
  float* nw = h3 + h.F3;          // [NNWP]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kHeadThreads / 32;

  // sum the split-K partials: warp w takes partials w, w+8, ... (coalesced rows of F1 floats), then the eight warp
  // sums are added in warp order -- a fixed order, so the result is deterministic
  {
    float* wsum = nw + h.NNWP;  // [8][F1] scratch
    for (int j0 = 0; j0 < h.F1; j0 += 32) {
      const int j = j0 + lane;
      float s = 0.f;
      if (j < h.F1)
        for (int p = warp; p < S; p += nwarp) s += partial[(static_cast<long long>(p) * h.B + b) * h.F1 + j];
      if (j < h.F1) wsum[warp * h.F1 + j] = s;
    }
    __syncthreads();
    for (int j = tid; j < h.F1; j += kHeadThreads) {
      float s = 0.f;
      for (int w = 0; w < nwarp; ++w) s += wsum[w * h.F1 + j];
      s += __ldg(h.b1 + j);
      s = (s < 0.f) ? 0.f : s;
      h1[j] = s;
      h.h1[static_cast<long long>(b) * h.F1 + j] = s;
    }
  }
  for (int i = tid; i < h.NNWP; i += kHeadThreads) nw[i] = h.nwp[static_cast<long long>(b) * h.NNWP + i];
  for (int i = tid; i < h.NPV; i += kHeadThreads) {
    const int t = i / h.pv_ns, s = i - t * h.pv_ns;
    cat[h.F2 + i] = nan_to_num_f(h.pv[b * h.pv_sb + t * h.pv_st + s]);
  }
  __syncthreads();
  for (int i = warp; i < h.F2; i += nwarp) {  // fc2 + ReLU
    float s = warp_dot(h.w2 + static_cast<long long>(i) * h.F1, h1, h.F1, lane) + __ldg(h.b2 + i);
    if (lane == 0) cat[i] = (s < 0.f) ? 0.f : s;
  }
  if (h.NNWP > 0) {
    for (int i = warp; i < h.FNWP; i += nwarp) {  // fc_nwp + ReLU
      float s = warp_dot(h.wn + static_cast<long long>(i) * h.NNWP, nw, h.NNWP, lane) + __ldg(h.bn + i);
      if (lane == 0) cat[h.F2 + h.NPV + i] = (s < 0.f) ? 0.f : s;
    }
  }
  __syncthreads();
  for (int i = tid; i < NCAT; i += kHeadThreads) h.cat[static_cast<long long>(b) * NCAT + i] = cat[i];
  for (int i = warp; i < h.F3; i += nwarp) {  // fc3 + ReLU
    float s = warp_dot(h.w3 + static_cast<long long>(i) * NCAT, cat, NCAT, lane) + __ldg(h.b3 + i);
    s = (s < 0.f) ? 0.f : s;
    if (lane == 0) { h3[i] = s; h.h3[static_cast<long long>(b) * h.F3 + i] = s; }
  }
  __syncthreads();
  for (int i = warp; i < h.FO; i += nwarp) {  // fc4
    float s = warp_dot(h.w4 + static_cast<long long>(i) * h.F3, h3, h.F3, lane) + __ldg(h.b4 + i);
    if (lane == 0) h.out[static_cast<long long>(b) * h.FO + i] = s;
  }
}

// ---- tail backward: one CTA per sample ------------------------------------------------------------------
__global__ void __launch_bounds__(kHeadThreads) head_tail_bwd_kernel(const pvb200_head_t h) {
  extern __shared__ float sm[];
  const int NCAT = h.F2 + h.NPV + (h.NNWP > 0 ? h.FNWP : 0);
  float* go = sm;            // [FO]
  float* g3 = go + h.FO;     // [F3]
  float* gc = g3 + h.F3;     // [NCAT]
  const int b = blockIdx.x, tid = threadIdx.x;

  for (int i = tid; i < h.FO; i += kHeadThreads) go[i] = h.g_out[static_cast<long long>(b) * h.FO + i];
  __syncthreads();
  for (int i = tid; i < h.F3; i += kHeadThreads) {  // through fc4, ReLU mask of fc3
    float s = 0.f;
    for (int o = 0; o < h.FO; ++o) s = fmaf(__ldg(h.w4 + static_cast<long long>(o) * h.F3 + i), go[o], s);
    s = (h.h3[static_cast<long long>(b) * h.F3 + i] > 0.f) ? s : 0.f;
    g3[i] = s;
    h.g_h3[static_cast<long long>(b) * h.F3 + i] = s;
  }
  __syncthreads();
  for (int c = tid; c < NCAT; c += kHeadThreads) {  // through fc3, split the concat, ReLU masks of fc2 / fc_nwp
    float s = 0.f;
    for (int i = 0; i < h.F3; ++i) s = fmaf(__ldg(h.w3 + static_cast<long long>(i) * NCAT + c), g3[i], s);
    const bool is_pv = (c >= h.F2) && (c < h.F2 + h.NPV);
    if (is_pv) s = 0.f;  // inputs need no gradient
    else s = (h.cat[static_cast<long long>(b) * NCAT + c] > 0.f) ? s : 0.f;
    gc[c] = s;
    h.g_cat[static_cast<long long>(b) * NCAT + c] = s;
  }
  __syncthreads();
  for (int j = tid; j < h.F1; j += kHeadThreads) {  // through fc2, ReLU mask of fc1
    float s = 0.f;
    for (int i = 0; i < h.F2; ++i) s = fmaf(__ldg(h.w2 + static_cast<long long>(i) * h.F1 + j), gc[i], s);
    s = (h.h1[static_cast<long long>(b) * h.F1 + j] > 0.f) ? s : 0.f;
    h.g_h1[static_cast<long long>(b) * h.F1 + j] = s;
  }
}

// ---- small weight gradients: dW[o][i] = sum_b gz[b*ldg + o] * in[b*ldi + i];  db[o] = sum_b gz -----------
__global__ void linear_wgrad_small_kernel(const float* __restrict__ gz, int ldg, const float* __restrict__ in, int ldi,
                                          float* __restrict__ dW, float* __restrict__ db, int B, int O, int I) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(O) * I;
  if (idx < n) {
    const int o = static_cast<int>(idx / I), i = static_cast<int>(idx - static_cast<long long>(o) * I);
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(gz[static_cast<long long>(b) * ldg + o], in[static_cast<long long>(b) * ldi + i], s);
    dW[idx] = s;
  } else if (idx < n + O && db) {
    const int o = static_cast<int>(idx - n);
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += gz[static_cast<long long>(b) * ldg + o];
    db[o] = s;
  }
}

// ---- fc1 weight gradient: dW1[j][k] = sum_b g1[b][j] * x[b][k] ----------------------------------------------
__global__ void __launch_bounds__(kHeadThreads)
fc1_wgrad_kernel(const float* __restrict__ g1, const float* __restrict__ x, float* __restrict__ dW, int B, int F1,
                 long long K1, int njt, int vec_ok) {
  constexpr int KT = 128;
  __shared__ __align__(16) float gs[kFc1BT * kFc1JT];  // [b][j]
  __shared__ __align__(16) float xs[kFc1BT * KT];      // [b][k]
  int id = blockIdx.x;
  const int jt = id % njt;
  const long long k0 = static_cast<long long>(id / njt) * KT;
  const int j0 = jt * kFc1JT;
  const int tid = threadIdx.x;
  const int tk = tid & 15;  // k columns 4tk..4tk+3 and 64+4tk..64+4tk+3
  const int tj = tid >> 4;  // j rows 8tj..8tj+7

  float acc[8][8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

  for (int b0 = 0; b0 < B; b0 += kFc1BT) {
    __syncthreads();
    for (int idx = tid; idx < kFc1BT * kFc1JT; idx += kHeadThreads) {
      const int r = idx >> 7, c = idx & 127;
      gs[idx] = (b0 + r < B && j0 + c < F1) ? g1[static_cast<long long>(b0 + r) * F1 + j0 + c] : 0.f;
    }
    if (vec_ok && k0 + KT <= K1) {
      for (int idx = tid; idx < kFc1BT * (KT / 4); idx += kHeadThreads) {
        const int r = idx >> 5, c4 = idx & 31;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b0 + r < B) v = __ldg(reinterpret_cast<const float4*>(x + static_cast<long long>(b0 + r) * K1 + k0) + c4);
        *reinterpret_cast<float4*>(xs + r * KT + 4 * c4) = v;
      }
    } else {
      for (int idx = tid; idx < kFc1BT * KT; idx += kHeadThreads) {
        const int r = idx >> 7, c = idx & 127;
        xs[idx] = (b0 + r < B && k0 + c < K1) ? x[static_cast<long long>(b0 + r) * K1 + k0 + c] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int b = 0; b < kFc1BT; ++b) {
      const float4 ga = *reinterpret_cast<const float4*>(gs + b * kFc1JT + 8 * tj);
      const float4 gb = *reinterpret_cast<const float4*>(gs + b * kFc1JT + 8 * tj + 4);
      const float4 xa = *reinterpret_cast<const float4*>(xs + b * KT + 4 * tk);
      const float4 xb = *reinterpret_cast<const float4*>(xs + b * KT + 64 + 4 * tk);
      const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
      const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(gv[r], xv[c], acc[r][c]);
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int j = j0 + 8 * tj + r;
    if (j >= F1) continue;
    float* row = dW + static_cast<long long>(j) * K1 + k0;
    if (vec_ok && k0 + KT <= K1) {
      __stcs(reinterpret_cast<float4*>(row + 4 * tk), make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]));
      __stcs(reinterpret_cast<float4*>(row + 64 + 4 * tk), make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]));
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const long long k = k0 + (c < 4 ? 4 * tk + c : 64 + 4 * tk + (c - 4));
        if (k < K1) dW[static_cast<long long>(j) * K1 + k] = acc[r][c];
      }
    }
  }
}

// ---- fc1 data gradient: gx[b][k] = (sum_j g1[b][j] * w[j][k]) * (x[b][k] > 0) ----------------------------------
__global__ void __launch_bounds__(kHeadThreads)
fc1_dgrad_kernel(const float* __restrict__ g1, const float* __restrict__ w, const float* __restrict__ x,
                 float* __restrict__ gx, int B, int F1, long long K1, int nbt, int vec_ok) {
  constexpr int KT = 128, JT = 32;
  __shared__ __align__(16) float ws[JT * KT];       // [j][k]
  __shared__ __align__(16) float gsT[JT * kFc1BT];  // [j][b]
  int id = blockIdx.x;
  const int bt = id % nbt;
  const long long k0 = static_cast<long long>(id / nbt) * KT;
  const int b0 = bt * kFc1BT;
  const int tid = threadIdx.x;
  const int tk = tid & 31;  // k columns 4tk..4tk+3
  const int tb = tid >> 5;  // batch rows 4tb..4tb+3 (warp-uniform)
  const bool full = vec_ok && (k0 + KT <= K1);

  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

  for (int j0 = 0; j0 < F1; j0 += JT) {
    __syncthreads();
    if (full) {
      for (int idx = tid; idx < JT * (KT / 4); idx += kHeadThreads) {
        const int r = idx >> 5, c4 = idx & 31;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j0 + r < F1) v = __ldcs(reinterpret_cast<const float4*>(w + static_cast<long long>(j0 + r) * K1 + k0) + c4);
        *reinterpret_cast<float4*>(ws + r * KT + 4 * c4) = v;
      }
    } else {
      for (int idx = tid; idx < JT * KT; idx += kHeadThreads) {
        const int r = idx >> 7, c = idx & 127;
        ws[idx] = (j0 + r < F1 && k0 + c < K1) ? w[static_cast<long long>(j0 + r) * K1 + k0 + c] : 0.f;
      }
    }
    for (int idx = tid; idx < JT * kFc1BT; idx += kHeadThreads) {
      const int r = idx >> 5, c = idx & 31;  // r: j, c: b
      gsT[idx] = (j0 + r < F1 && b0 + c < B) ? g1[static_cast<long long>(b0 + c) * F1 + j0 + r] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < JT; ++j) {
      const float4 wv = *reinterpret_cast<const float4*>(ws + j * KT + 4 * tk);
      const float4 gv = *reinterpret_cast<const float4*>(gsT + j * kFc1BT + 4 * tb);
      const float g[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        acc[r][0] = fmaf(g[r], wv.x, acc[r][0]);
        acc[r][1] = fmaf(g[r], wv.y, acc[r][1]);
        acc[r][2] = fmaf(g[r], wv.z, acc[r][2]);
        acc[r][3] = fmaf(g[r], wv.w, acc[r][3]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int b = b0 + 4 * tb + r;
    if (b >= B) continue;
    const long long base = static_cast<long long>(b) * K1 + k0 + 4 * tk;
    if (full) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + base));
      float4 o;
      o.x = xv.x > 0.f ? acc[r][0] : 0.f;
      o.y = xv.y > 0.f ? acc[r][1] : 0.f;
      o.z = xv.z > 0.f ? acc[r][2] : 0.f;
      o.w = xv.w > 0.f ? acc[r][3] : 0.f;
      *reinterpret_cast<float4*>(gx + base) = o;
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (k0 + 4 * tk + c < K1) gx[base + c] = (x[base + c] > 0.f) ? acc[r][c] : 0.f;
    }
  }
}

// ---- fc1 data gradient, fast path: gx[b][k] = (sum_j g1[b][j] * w[j][k]) * (x[b][k] > 0) ---------------------------------
// Persistent CTAs (one per SM and batch tile) stream [128 j][128 k] slabs of W1 (64 KB) through a 2-stage cp.async
// pipeline: the next slab lands while the current one is in the FMA loop and while the result is reduced, masked and
// stored.  256 threads = 4 j-groups (j = grp mod 4) x 64 threads with 8 (b) x 8 (k) register tiles (k columns split
// 4 + 4 so LDS.128 across lanes is conflict-free); the j-groups are summed through smem in fixed order.
constexpr int kFc1DgThreads = 256;
__global__ void __launch_bounds__(kFc1DgThreads, 1)
fc1_dgrad_v2_kernel(const float* __restrict__ g1, const float* __restrict__ w, const float* __restrict__ x,
                    float* __restrict__ gx, int B, int F1, long long K1, long long ktiles) {
  constexpr int KT = 128;
  extern __shared__ __align__(16) float smd[];
  float* gsT = smd;                       // [128 j][32 b]
  float* wbuf = smd + 128 * kFc1BT;       // [2][128 j][128 k]; a consumed buffer is reused as [4][32 b][128 k]
  const int b0 = blockIdx.y * kFc1BT;
  const int tid = threadIdx.x;
  const int grp = tid >> 6, t64 = tid & 63;
  const int tk = t64 & 15;  // k columns 4tk..4tk+3 and 64+4tk..64+4tk+3
  const int tb = t64 >> 4;  // batch rows 8tb..8tb+7

  auto issue = [&](long long tile, int buf) {
    const long long k0 = tile * KT;
    const uint32_t ws0 = static_cast<uint32_t>(__cvta_generic_to_shared(wbuf + buf * 128 * KT));
    for (int idx = tid; idx < 128 * (KT / 4); idx += kFc1DgThreads) {
      const int r = idx >> 5, c4 = idx & 31;
      const bool ok = (r < F1) && (k0 + 4 * c4 < K1);
      const float* src = w + static_cast<long long>(ok ? r : 0) * K1 + (ok ? k0 + 4 * c4 : 0);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ws0 + 16u * idx), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  for (int idx = tid; idx < 128 * kFc1BT; idx += kFc1DgThreads) {
    const int j = idx >> 5, b = idx & 31;
    gsT[idx] = (j < F1 && b0 + b < B) ? g1[static_cast<long long>(b0 + b) * F1 + j] : 0.f;
  }
  long long tile = blockIdx.x;
  if (tile < ktiles) issue(tile, 0);
  for (int it = 0; tile < ktiles; ++it, tile += gridDim.x) {
    const long long next = tile + gridDim.x;
    if (next < ktiles) {
      issue(next, (it + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    float* ws = wbuf + (it & 1) * 128 * KT;
    float acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
#pragma unroll 4
    for (int jj = 0; jj < 32; ++jj) {
      const int j = 4 * jj + grp;
      const float4 wa = *reinterpret_cast<const float4*>(ws + j * KT + 4 * tk);
      const float4 wb = *reinterpret_cast<const float4*>(ws + j * KT + 64 + 4 * tk);
      const float4 ga = *reinterpret_cast<const float4*>(gsT + j * kFc1BT + 8 * tb);
      const float4 gb = *reinterpret_cast<const float4*>(gsT + j * kFc1BT + 8 * tb + 4);
      const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
      const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(gv[r], wv[c], acc[r][c]);
    }
    __syncthreads();  // everyone is done reading this W slab
    float* red = ws;  // [4][32 b][128 k]
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float* row = red + (grp * kFc1BT + 8 * tb + r) * KT;
      *reinterpret_cast<float4*>(row + 4 * tk) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
      *reinterpret_cast<float4*>(row + 64 + 4 * tk) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
    }
    __syncthreads();
    const long long k0 = tile * KT;
    for (int idx = tid; idx < kFc1BT * (KT / 4); idx += kFc1DgThreads) {
      const int b = idx >> 5, c4 = idx & 31;
      if (b0 + b >= B || k0 + 4 * c4 >= K1) continue;
      const float4 p0 = *reinterpret_cast<const float4*>(red + (0 * kFc1BT + b) * KT + 4 * c4);
      const float4 p1 = *reinterpret_cast<const float4*>(red + (1 * kFc1BT + b) * KT + 4 * c4);
      const float4 p2 = *reinterpret_cast<const float4*>(red + (2 * kFc1BT + b) * KT + 4 * c4);
      const float4 p3 = *reinterpret_cast<const float4*>(red + (3 * kFc1BT + b) * KT + 4 * c4);
      const long long off = static_cast<long long>(b0 + b) * K1 + k0 + 4 * c4;
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + off));
      float4 o;
      o.x = xv.x > 0.f ? (p0.x + p1.x) + (p2.x + p3.x) : 0.f;
      o.y = xv.y > 0.f ? (p0.y + p1.y) + (p2.y + p3.y) : 0.f;
      o.z = xv.z > 0.f ? (p0.z + p1.z) + (p2.z + p3.z) : 0.f;
      o.w = xv.w > 0.f ? (p0.w + p1.w) + (p2.w + p3.w) : 0.f;
      __stcs(reinterpret_cast<float4*>(gx + off), o);
    }
    __syncthreads();  // the buffer may be refilled by the prefetch issued at the top of the next iteration
  }
}

// ---- fc1 weight gradient, fast path: dW1[j][k] = sum_b g1[b][j] * x[b][k] ---------------------------------------------
// Persistent CTAs; pipeline items are (k tile, batch tile) pairs: x [32 b][128 k] + g1 [32 b][128 j] arrive by cp.async
// one item ahead; 8 (j) x 8 (k) register tiles accumulate over the batch tiles and are streamed out (565 MB, the
// HBM-write-bound part of the step) while the next item is already in flight.
__global__ void __launch_bounds__(kHeadThreads, 2)
fc1_wgrad_v2_kernel(const float* __restrict__ g1, const float* __restrict__ x, float* __restrict__ dW, int B, int F1,
                    long long K1, long long ktiles, int nbt) {
  constexpr int KT = 128;
  constexpr int STAGE = kFc1BT * (kFc1JT + KT);
  extern __shared__ __align__(16) float smw[];  // [2][ gs 32x128 | xs 32x128 ]
  const int tid = threadIdx.x;
  const int tk = tid & 15;  // k columns 4tk..4tk+3 and 64+4tk..64+4tk+3
  const int tj = tid >> 4;  // j rows 8tj..8tj+7
  const long long items = ktiles * nbt;

  auto issue = [&](long long item, int buf) {
    const long long k0 = (item / nbt) * KT;
    const int b0 = static_cast<int>(item % nbt) * kFc1BT;
    const uint32_t gs0 = static_cast<uint32_t>(__cvta_generic_to_shared(smw + buf * STAGE));
    const uint32_t xs0 = gs0 + 4u * kFc1BT * kFc1JT;
    for (int idx = tid; idx < kFc1BT * (kFc1JT / 4); idx += kHeadThreads) {
      const int r = idx >> 5, c4 = idx & 31;
      const bool ok = (b0 + r < B) && (4 * c4 < F1);  // F1 % 4 == 0 on this path
      const float* src = g1 + static_cast<long long>(ok ? b0 + r : 0) * F1 + (ok ? 4 * c4 : 0);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(gs0 + 16u * idx), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    for (int idx = tid; idx < kFc1BT * (KT / 4); idx += kHeadThreads) {
      const int r = idx >> 5, c4 = idx & 31;
      const bool ok = (b0 + r < B) && (k0 + 4 * c4 < K1);
      const float* src = x + static_cast<long long>(ok ? b0 + r : 0) * K1 + (ok ? k0 + 4 * c4 : 0);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(xs0 + 16u * idx), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  float acc[8][8];
  long long item = static_cast<long long>(blockIdx.x) * nbt;  // a CTA owns whole k tiles (all their batch tiles)
  const long long stride = static_cast<long long>(gridDim.x) * nbt;
  if (item < items) issue(item, 0);
  int it = 0;
  for (; item < items; item += stride) {
    for (int bt = 0; bt < nbt; ++bt, ++it) {
      const long long cur = item + bt;
      long long next = cur + 1;
      if (bt == nbt - 1) next = item + stride;
      if (next < items) {
        issue(next, (it + 1) & 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();
      const float* gs = smw + (it & 1) * STAGE;
      const float* xs = gs + kFc1BT * kFc1JT;
      if (bt == 0) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
      }
#pragma unroll 4
      for (int b = 0; b < kFc1BT; ++b) {
        const float4 ga = *reinterpret_cast<const float4*>(gs + b * kFc1JT + 8 * tj);
        const float4 gb = *reinterpret_cast<const float4*>(gs + b * kFc1JT + 8 * tj + 4);
        const float4 xa = *reinterpret_cast<const float4*>(xs + b * KT + 4 * tk);
        const float4 xb = *reinterpret_cast<const float4*>(xs + b * KT + 64 + 4 * tk);
        const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
        const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(gv[r], xv[c], acc[r][c]);
      }
      __syncthreads();  // stage consumed
    }
    const long long k0 = (item / nbt) * KT;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int j = 8 * tj + r;
      if (j >= F1) continue;
      float* row = dW + static_cast<long long>(j) * K1 + k0;
      if (k0 + 4 * tk < K1) __stcs(reinterpret_cast<float4*>(row + 4 * tk), make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]));
      if (k0 + 64 + 4 * tk < K1)
        __stcs(reinterpret_cast<float4*>(row + 64 + 4 * tk), make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]));
    }
  }
}

static int check_head(const pvb200_head_t* h) {
  PVB_REQUIRE(h != nullptr, "head: null descriptor");
  PVB_REQUIRE(h->struct_size == sizeof(pvb200_head_t), "head: struct_size %zu != %zu (ABI mismatch)", h->struct_size,
              sizeof(pvb200_head_t));
  PVB_REQUIRE(h->B > 0 && h->F1 > 0 && h->F2 > 0 && h->F3 > 0 && h->FO > 0 && h->K1 > 0, "head: bad sizes");
  PVB_REQUIRE(h->NPV >= 0 && h->NNWP >= 0, "head: bad branch sizes");
  PVB_REQUIRE(h->NPV == 0 || (h->pv && h->pv_ns > 0 && h->NPV % h->pv_ns == 0), "head: bad PV-history description");
  PVB_REQUIRE(h->NNWP == 0 || (h->nwp && h->wn && h->bn && h->FNWP > 0), "head: NWP branch needs nwp, wn, bn");
  PVB_REQUIRE(h->w1 && h->b1 && h->w2 && h->b2 && h->w3 && h->b3 && h->w4 && h->b4, "head: null parameter");
  return PVB200_OK;
}

// fc1 on the tensor cores (fc1_bf16x3.cu); the FMA-pipe kernels above serve the shapes it does not take
bool fc1x3_ok(int B, int max_b, int F1, long long K1, const void* x, const void* w);
int fc1x3_fwd_ctas(long long K1);
int fc1x3_fwd(const float* x, const float* w, float* partial, int S, int B, int F1, long long K1, cudaStream_t st);
int fc1x3_dgrad(const float* g, const float* w, const float* x, float* gx, int B, int F1, long long K1, cudaStream_t st);
int fc1x3_wgrad(const float* g, const float* x, float* dw, int B, int F1, long long K1, cudaStream_t st);
constexpr int kFc1x3FwdMaxB = 1 << 30, kFc1x3BwdMaxB = 1 << 30;  // all three kernels chunk their batch (48 samples per launch)

static bool vec4_ok(const void* p, long long ld) { return (ld % 4 == 0) && (reinterpret_cast<uintptr_t>(p) % 16 == 0); }

}  // namespace pvb

extern "C" {

size_t pvb200_head_fwd_workspace_bytes(int B, int F1, long long K1) {
  if (B <= 0 || F1 <= 0 || K1 <= 0) return 0;
  const pvb::Fc1Plan p = pvb::fc1_plan(B, F1, K1);
  const int s3 = pvb::fc1x3_fwd_ctas(K1);
  return static_cast<size_t>(p.S > s3 ? p.S : s3) * B * F1 * sizeof(float);
}

/* tail only: h->workspace holds S split-K partials [S][B][F1] of fc1 (e.g. from pvb200_fc1_fwd_bf16); fills h1, cat, h3, out */
int pvb200_head_tail_fwd_f32(const pvb200_head_t* h, int S, pvb200_stream_t stream) {
  using namespace pvb;
  int rc = check_head(h);
  if (rc) return rc;
  PVB_REQUIRE(h->h1 && h->cat && h->h3 && h->out && S > 0, "head_tail_fwd: null output");
  PVB_REQUIRE(h->workspace && h->workspace_bytes >= static_cast<size_t>(S) * h->B * h->F1 * sizeof(float),
              "head_tail_fwd: partials buffer too small");
  const int NCAT = h->F2 + h->NPV + (h->NNWP > 0 ? h->FNWP : 0);
  const size_t smem = static_cast<size_t>(h->F1 + NCAT + h->F3 + h->NNWP + 8 * h->F1) * sizeof(float);
  PVB_REQUIRE(smem <= 48 * 1024, "head_tail_fwd: feature sizes too large for the tail kernel (%zu B smem)", smem);
  head_tail_fwd_kernel<<<h->B, kHeadThreads, smem, as_stream(stream)>>>(*h, static_cast<const float*>(h->workspace), S);
  PVB_LAUNCHED("head_tail_fwd");
  return PVB200_OK;
}

int pvb200_head_fwd_f32(const pvb200_head_t* h, pvb200_stream_t stream) {
  using namespace pvb;
  int rc = check_head(h);
  if (rc) return rc;
  PVB_REQUIRE(h->x, "head_fwd: null features");
  PVB_REQUIRE(h->h1 && h->cat && h->h3 && h->out, "head_fwd: null output");
  Fc1Plan p = fc1_plan(h->B, h->F1, h->K1);
  const bool tc3 = fc1x3_ok(h->B, kFc1x3FwdMaxB, h->F1, h->K1, h->x, h->w1);
  if (tc3) {
    p.S = fc1x3_fwd_ctas(h->K1);
    PVB_REQUIRE(p.S > 0, "head_fwd: no CUDA device");
  }
  const size_t need = static_cast<size_t>(p.S) * h->B * h->F1 * sizeof(float);
  if (!h->workspace || h->workspace_bytes < need) {
    set_error("head_fwd: workspace too small (%zu < %zu bytes)", h->workspace_bytes, need);
    return PVB200_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  float* partial = static_cast<float*>(h->workspace);
  const int NCAT = h->F2 + h->NPV + (h->NNWP > 0 ? h->FNWP : 0);
  const size_t smem = static_cast<size_t>(h->F1 + NCAT + h->F3 + h->NNWP + 8 * h->F1) * sizeof(float);
  PVB_REQUIRE(smem <= 48 * 1024, "head_fwd: feature sizes too large for the tail kernel (%zu B smem)", smem);
  if (tc3) {
    if ((rc = fc1x3_fwd(h->x, h->w1, partial, p.S, h->B, h->F1, h->K1, st))) return rc;
    head_tail_fwd_kernel<<<h->B, kHeadThreads, smem, st>>>(*h, partial, p.S);
    PVB_LAUNCHED("head_tail_fwd");
    return PVB200_OK;
  }
  const int vec = vec4_ok(h->x, h->K1) && vec4_ok(h->w1, h->K1);
  const long long ctas = static_cast<long long>(p.S) * p.nbt * p.njt;
  if (vec) {
    const size_t smem2 = 2 * static_cast<size_t>(kFc1JT + kFc1BT) * (kFc1KT + 4) * sizeof(float);  // 87 KB >= the 64 KB reduce buffer
    PVB_CUDA(cudaFuncSetAttribute(fc1_fwd_splitk_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    fc1_fwd_splitk_v2_kernel<<<static_cast<unsigned>(ctas), kFc1V2Threads, smem2, st>>>(h->x, h->w1, partial, h->B, h->F1, h->K1,
                                                                                        p.nbt, p.njt, p.k_per_split);
  } else {
    fc1_fwd_splitk_kernel<<<static_cast<unsigned>(ctas), kHeadThreads, 0, st>>>(h->x, h->w1, partial, h->B, h->F1, h->K1,
                                                                                 p.nbt, p.njt, p.k_per_split, vec);
  }
  PVB_LAUNCHED("fc1_fwd_splitk");
  head_tail_fwd_kernel<<<h->B, kHeadThreads, smem, st>>>(*h, partial, p.S);
  PVB_LAUNCHED("head_tail_fwd");
  return PVB200_OK;
}

static int head_bwd_impl(const pvb200_head_t* h, bool with_fc1, pvb200_stream_t stream);

int pvb200_head_bwd_f32(const pvb200_head_t* h, pvb200_stream_t stream) { return head_bwd_impl(h, true, stream); }

/* tail only: everything except fc1's weight / data gradient (g_h1 and db1 ARE produced); dw1 and x may be NULL */
int pvb200_head_tail_bwd_f32(const pvb200_head_t* h, pvb200_stream_t stream) { return head_bwd_impl(h, false, stream); }

}  // extern "C"

static int head_bwd_impl(const pvb200_head_t* h, bool with_fc1, pvb200_stream_t stream) {
  using namespace pvb;
  int rc = check_head(h);
  if (rc) return rc;
  PVB_REQUIRE(h->h1 && h->cat && h->h3 && h->g_out && h->g_h3 && h->g_cat && h->g_h1, "head_bwd: null buffer");
  PVB_REQUIRE(h->db1 && h->dw2 && h->db2 && h->dw3 && h->db3 && h->dw4 && h->db4, "head_bwd: null grad");
  PVB_REQUIRE(!with_fc1 || (h->dw1 && h->x), "head_bwd: null fc1 gradient / features");
  PVB_REQUIRE(h->NNWP == 0 || (h->dwn && h->dbn), "head_bwd: NWP branch needs dwn, dbn");
  cudaStream_t st = as_stream(stream);
  const int NCAT = h->F2 + h->NPV + (h->NNWP > 0 ? h->FNWP : 0);
  const size_t smem = static_cast<size_t>(h->FO + h->F3 + NCAT) * sizeof(float);
  PVB_REQUIRE(smem <= 48 * 1024, "head_bwd: feature sizes too large for the tail kernel (%zu B smem)", smem);
  head_tail_bwd_kernel<<<h->B, kHeadThreads, smem, st>>>(*h);
  PVB_LAUNCHED("head_tail_bwd");

  auto small = [&](const float* gz, int ldg, const float* in, int ldi, float* dW, float* db, int O, int I) -> int {
    const long long n = static_cast<long long>(O) * I + O;
    linear_wgrad_small_kernel<<<static_cast<unsigned>(ceil_div(n, 256LL)), 256, 0, st>>>(gz, ldg, in, ldi, dW, db, h->B, O, I);
    PVB_LAUNCHED("linear_wgrad_small");
    return PVB200_OK;
  };
  if ((rc = small(h->g_out, h->FO, h->h3, h->F3, h->dw4, h->db4, h->FO, h->F3))) return rc;
  if ((rc = small(h->g_h3, h->F3, h->cat, NCAT, h->dw3, h->db3, h->F3, NCAT))) return rc;
  if ((rc = small(h->g_cat, NCAT, h->h1, h->F1, h->dw2, h->db2, h->F2, h->F1))) return rc;
  if (h->NNWP > 0)
    if ((rc = small(h->g_cat + h->F2 + h->NPV, NCAT, h->nwp, h->NNWP, h->dwn, h->dbn, h->FNWP, h->NNWP))) return rc;
  // fc1 bias gradient (I = 0 columns: only the db part of the kernel runs)
  {
    linear_wgrad_small_kernel<<<ceil_div(h->F1, 256), 256, 0, st>>>(h->g_h1, h->F1, h->g_h1, h->F1, h->db1, h->db1, h->B,
                                                                    h->F1, 0);
    PVB_LAUNCHED("fc1_bias_grad");
  }
  if (!with_fc1) return PVB200_OK;
  if (fc1x3_ok(h->B, kFc1x3BwdMaxB, h->F1, h->K1, h->x, h->w1) && vec4_ok(h->dw1, h->K1) && (!h->g_x || vec4_ok(h->g_x, h->K1))) {
    if ((rc = fc1x3_wgrad(h->g_h1, h->x, h->dw1, h->B, h->F1, h->K1, st))) return rc;
    if (h->g_x && (rc = fc1x3_dgrad(h->g_h1, h->w1, h->x, h->g_x, h->B, h->F1, h->K1, st))) return rc;
    return PVB200_OK;
  }
  const int vec = vec4_ok(h->x, h->K1) && vec4_ok(h->w1, h->K1) && vec4_ok(h->dw1, h->K1) && (!h->g_x || vec4_ok(h->g_x, h->K1));
  const int njt = ceil_div(h->F1, kFc1JT);
  const long long kt = ceil_div(h->K1, 128LL);
  PVB_REQUIRE(kt * njt <= 0x7fffffffLL, "head_bwd: K1 too large");
  if (vec && h->F1 <= 128 && h->F1 % 4 == 0 && vec4_ok(h->g_h1, h->F1)) {
    const size_t smemw = 2 * static_cast<size_t>(kFc1BT) * (kFc1JT + 128) * sizeof(float);
    PVB_CUDA(cudaFuncSetAttribute(fc1_wgrad_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemw));
    const int sms = sm_count();
    PVB_REQUIRE(sms > 0, "head_bwd: no CUDA device");
    const long long gw = kt < 2LL * sms ? kt : 2LL * sms;
    fc1_wgrad_v2_kernel<<<static_cast<unsigned>(gw), kHeadThreads, smemw, st>>>(h->g_h1, h->x, h->dw1, h->B, h->F1, h->K1, kt,
                                                                               ceil_div(h->B, kFc1BT));
  } else {
    fc1_wgrad_kernel<<<static_cast<unsigned>(kt * njt), kHeadThreads, 0, st>>>(h->g_h1, h->x, h->dw1, h->B, h->F1, h->K1, njt, vec);
  }
  PVB_LAUNCHED("fc1_wgrad");
  if (h->g_x) {
    const int nbt = ceil_div(h->B, kFc1BT);
    PVB_REQUIRE(kt * nbt <= 0x7fffffffLL, "head_bwd: K1*B too large");
    if (vec && h->F1 <= 128) {
      const size_t smemd = (128 * kFc1BT + 2 * 128 * 128) * sizeof(float);
      PVB_CUDA(cudaFuncSetAttribute(fc1_dgrad_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemd));
      const int sms = sm_count();
      PVB_REQUIRE(sms > 0, "head_bwd: no CUDA device");
      const long long gx_ = kt < sms ? kt : sms;
      fc1_dgrad_v2_kernel<<<dim3(static_cast<unsigned>(gx_), nbt), kFc1DgThreads, smemd, st>>>(h->g_h1, h->w1, h->x, h->g_x, h->B,
                                                                                                h->F1, h->K1, kt);
    } else {
      fc1_dgrad_kernel<<<static_cast<unsigned>(kt * nbt), kHeadThreads, 0, st>>>(h->g_h1, h->w1, h->x, h->g_x, h->B, h->F1, h->K1, nbt, vec);
    }
    PVB_LAUNCHED("fc1_dgrad");
  }
  return PVB200_OK;
}



// =====================================================================================================================
// Generic Linear (+ReLU) forward / backward, torch layout w[N][K]  (the heads of conv3d_sat_nwp,
// predict_pv_yield/models/conv3d/model_sat_nwp.py:102-172,196-266 -- fc1/fc2 of the satellite tower, nwp_fc1/nwp_fc2,
// pv_fc1, fc3, fc4 -- are plain nn.Linear layers joined by torch.cat).  Rows may be column slices of wider buffers
// (ld* strides), which is how the concatenations are done without copies.  K >= 8192 ("tower" layers, 1-2 M input
// features) goes through the weight-streaming fc1 kernels above, everything else through small per-sample kernels.
// =====================================================================================================================
namespace pvb {

constexpr long long kLinBigK = 8192;

// y[b][n] = act(bias[n] + sum_s partial[s][b][n])
__global__ void linear_finish_kernel(const float* __restrict__ partial, int S, const float* __restrict__ bias, float* __restrict__ y,
                                     long long ldy, int B, int N, int relu) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N) return;
  const int b = idx / N, n = idx - b * N;
  float s = 0.f;
  for (int i = 0; i < S; ++i) s += partial[(static_cast<long long>(i) * B + b) * N + n];  // fixed order: deterministic
  s += bias ? bias[n] : 0.f;
  y[b * ldy + n] = (relu && s < 0.f) ? 0.f : s;
}

// one CTA per sample: x row in shared memory, one warp per output feature (strided), lanes over K
__global__ void __launch_bounds__(256) linear_fwd_small_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                                                               const float* __restrict__ bias, float* __restrict__ y, long long ldy,
                                                               int K, int N, int relu) {
  extern __shared__ float xs[];
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += blockDim.x) xs[k] = x[b * ldx + k];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int n = warp; n < N; n += blockDim.x >> 5) {
    const float* wr = w + static_cast<long long>(n) * K;
    float s = 0.f;
    for (int k = lane; k < K; k += 32) s = fmaf(xs[k], __ldg(wr + k), s);
    s = warp_sum(s);
    if (lane == 0) {
      s += bias ? bias[n] : 0.f;
      y[b * ldy + n] = (relu && s < 0.f) ? 0.f : s;
    }
  }
}

// g_pre[b][n] = gy[b][n] * (y[b][n] > 0 if relu)
__global__ void linear_gpre_kernel(const float* __restrict__ gy, long long ldgy, const float* __restrict__ y, long long ldy,
                                   float* __restrict__ g_pre, int B, int N) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * N) return;
  const int b = idx / N, n = idx - b * N;
  const float g = gy[b * ldgy + n];
  g_pre[idx] = (y == nullptr || y[b * ldy + n] > 0.f) ? g : 0.f;
}

// gx[b][k] = sum_n g_pre[b][n] * w[n][k]   (optionally masked by x[b][k] > 0); one CTA per (sample, 256 columns)
__global__ void __launch_bounds__(256) linear_dgrad_small_kernel(const float* __restrict__ g_pre, const float* __restrict__ w,
                                                                 const float* __restrict__ x, long long ldx, float* __restrict__ gx,
                                                                 long long ldgx, int K, int N) {
  extern __shared__ float gs[];
  const int b = blockIdx.y;
  for (int n = threadIdx.x; n < N; n += blockDim.x) gs[n] = g_pre[static_cast<long long>(b) * N + n];
  __syncthreads();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float s = 0.f;
  for (int n = 0; n < N; ++n) s = fmaf(gs[n], __ldg(w + static_cast<long long>(n) * K + k), s);
  if (x != nullptr && !(x[b * ldx + k] > 0.f)) s = 0.f;
  gx[b * ldgx + k] = s;
}

__global__ void embedding_fwd_kernel(const float* __restrict__ table, const int* __restrict__ ids, float* __restrict__ y,
                                     long long ldy, int B, int V, int D) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * D) return;
  const int b = idx / D, d = idx - b * D;
  const int v = ids[b];
  y[b * ldy + d] = (v >= 0 && v < V) ? table[static_cast<long long>(v) * D + d] : 0.f;
}

__global__ void embedding_bwd_kernel(const float* __restrict__ gy, long long ldgy, const int* __restrict__ ids,
                                     float* __restrict__ dtable, int B, int V, int D) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= V * D) return;
  const int v = idx / D, d = idx - v * D;
  float s = 0.f;
  for (int b = 0; b < B; ++b)
    if (ids[b] == v) s += gy[b * ldgy + d];
  dtable[idx] = s;
}

__global__ void history_flatten_kernel(const float* __restrict__ src, long long sb, long long st, float* __restrict__ dst,
                                       long long lddst, int B, int nt, int ns) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * nt * ns) return;
  const int s = idx % ns;
  const int t = (idx / ns) % nt;
  const int b = idx / (ns * nt);
  const float v = src[b * sb + t * st + s];
  dst[b * lddst + t * ns + s] = nan_to_num_f(v);  // torch.nan_to_num defaults: NaN -> 0, +-inf -> +-FLT_MAX
}

}  // namespace pvb

extern "C" {

size_t pvb200_linear_workspace_bytes(int B, int N, long long K) {
  if (B <= 0 || N <= 0 || K <= 0) return 0;
  size_t ws = static_cast<size_t>(B) * N * sizeof(float);  // g_pre of the backward pass
  if (K >= pvb::kLinBigK) {
    const pvb::Fc1Plan p = pvb::fc1_plan(B, N, K);
    const size_t f = static_cast<size_t>(p.S) * B * N * sizeof(float);
    if (f > ws) ws = f;
  }
  return ws;
}

int pvb200_linear_fwd_f32(const float* x, long long ldx, const float* w, const float* bias, float* y, long long ldy, int B,
                          long long K, int N, int relu, void* workspace, size_t workspace_bytes, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(x && w && y && B > 0 && K > 0 && N > 0 && ldx >= K && ldy >= N, "linear_fwd: bad argument");
  cudaStream_t st = as_stream(stream);
  if (K >= kLinBigK) {
    PVB_REQUIRE(ldx == K, "linear_fwd: K=%lld >= %lld needs contiguous rows", K, kLinBigK);
    const Fc1Plan p = fc1_plan(B, N, K);
    const size_t need = static_cast<size_t>(p.S) * B * N * sizeof(float);
    if (!workspace || workspace_bytes < need) {
      set_error("linear_fwd: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
      return PVB200_ERR_WORKSPACE;
    }
    float* partial = static_cast<float*>(workspace);
    const long long ctas = static_cast<long long>(p.S) * p.nbt * p.njt;
    if (vec4_ok(x, K) && vec4_ok(w, K)) {
      const size_t smem2 = 2 * static_cast<size_t>(kFc1JT + kFc1BT) * (kFc1KT + 4) * sizeof(float);
      PVB_CUDA(cudaFuncSetAttribute(fc1_fwd_splitk_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      fc1_fwd_splitk_v2_kernel<<<static_cast<unsigned>(ctas), kFc1V2Threads, smem2, st>>>(x, w, partial, B, N, K, p.nbt, p.njt,
                                                                                          p.k_per_split);
    } else {
      fc1_fwd_splitk_kernel<<<static_cast<unsigned>(ctas), kHeadThreads, 0, st>>>(x, w, partial, B, N, K, p.nbt, p.njt, p.k_per_split, 0);
    }
    PVB_LAUNCHED("linear_fwd_splitk");
    linear_finish_kernel<<<ceil_div(B * N, 256), 256, 0, st>>>(partial, p.S, bias, y, ldy, B, N, relu);
    PVB_LAUNCHED("linear_finish");
    return PVB200_OK;
  }
  PVB_REQUIRE(K * sizeof(float) <= 48 * 1024, "linear_fwd: K=%lld too large for the small kernel", K);
  linear_fwd_small_kernel<<<B, 256, static_cast<size_t>(K) * sizeof(float), st>>>(x, ldx, w, bias, y, ldy, static_cast<int>(K), N, relu);
  PVB_LAUNCHED("linear_fwd_small");
  return PVB200_OK;
}

/* gy = gradient w.r.t. the layer's OUTPUT (post-activation); y = the saved output when the layer has a ReLU, else NULL.
 * Writes dw [N][K], db [N] and, if gx != NULL, gx = g_pre . W (times (x > 0) when mask_gx_with_x: the input is itself a
 * post-ReLU activation whose producer wants the gradient of its PRE-activation).  workspace >= B*N floats. */
int pvb200_linear_bwd_f32(const float* x, long long ldx, const float* w, const float* y, long long ldy, const float* gy,
                          long long ldgy, float* gx, long long ldgx, int mask_gx_with_x, float* dw, float* db, int B, long long K,
                          int N, void* workspace, size_t workspace_bytes, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(x && w && gy && dw && db && B > 0 && K > 0 && N > 0 && ldx >= K && ldgy >= N, "linear_bwd: bad argument");
  PVB_REQUIRE(workspace && workspace_bytes >= static_cast<size_t>(B) * N * sizeof(float), "linear_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  float* g_pre = static_cast<float*>(workspace);
  linear_gpre_kernel<<<ceil_div(B * N, 256), 256, 0, st>>>(gy, ldgy, y, ldy, g_pre, B, N);
  PVB_LAUNCHED("linear_gpre");
  if (K >= kLinBigK) {
    PVB_REQUIRE(ldx == K && (!gx || ldgx == K), "linear_bwd: K=%lld >= %lld needs contiguous rows", K, kLinBigK);
    PVB_REQUIRE(!gx || mask_gx_with_x, "linear_bwd: the streaming data gradient always applies the input's ReLU mask");
    // bias gradient (I = 0 columns: only the db part of the kernel runs)
    linear_wgrad_small_kernel<<<ceil_div(N, 256), 256, 0, st>>>(g_pre, N, g_pre, N, db, db, B, N, 0);
    PVB_LAUNCHED("linear_bias_grad");
    const int vec = vec4_ok(x, K) && vec4_ok(w, K) && vec4_ok(dw, K) && (!gx || vec4_ok(gx, K));
    const int njt = ceil_div(N, kFc1JT);
    const long long kt = ceil_div(K, 128LL);
    PVB_REQUIRE(kt * njt <= 0x7fffffffLL, "linear_bwd: K too large");
    const int sms = sm_count();
    PVB_REQUIRE(sms > 0, "linear_bwd: no CUDA device");
    if (vec && N <= 128 && N % 4 == 0 && vec4_ok(g_pre, N)) {
      const size_t smemw = 2 * static_cast<size_t>(kFc1BT) * (kFc1JT + 128) * sizeof(float);
      PVB_CUDA(cudaFuncSetAttribute(fc1_wgrad_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemw));
      const long long gw = kt < 2LL * sms ? kt : 2LL * sms;
      fc1_wgrad_v2_kernel<<<static_cast<unsigned>(gw), kHeadThreads, smemw, st>>>(g_pre, x, dw, B, N, K, kt, ceil_div(B, kFc1BT));
    } else {
      fc1_wgrad_kernel<<<static_cast<unsigned>(kt * njt), kHeadThreads, 0, st>>>(g_pre, x, dw, B, N, K, njt, vec);
    }
    PVB_LAUNCHED("linear_wgrad_stream");
    if (gx) {
      const int nbt = ceil_div(B, kFc1BT);
      if (vec && N <= 128) {
        const size_t smemd = (128 * kFc1BT + 2 * 128 * 128) * sizeof(float);
        PVB_CUDA(cudaFuncSetAttribute(fc1_dgrad_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemd));
        const long long gx_ = kt < sms ? kt : sms;
        fc1_dgrad_v2_kernel<<<dim3(static_cast<unsigned>(gx_), nbt), kFc1DgThreads, smemd, st>>>(g_pre, w, x, gx, B, N, K, kt);
      } else {
        fc1_dgrad_kernel<<<static_cast<unsigned>(kt * nbt), kHeadThreads, 0, st>>>(g_pre, w, x, gx, B, N, K, nbt, vec);
      }
      PVB_LAUNCHED("linear_dgrad_stream");
    }
    return PVB200_OK;
  }
  PVB_REQUIRE(N * sizeof(float) <= 48 * 1024 && K <= 0x7fffffffLL && ldx <= 0x7fffffffLL, "linear_bwd: sizes too large for the small kernels");
  {
    const long long n = static_cast<long long>(N) * K + N;
    linear_wgrad_small_kernel<<<static_cast<unsigned>(ceil_div(n, 256LL)), 256, 0, st>>>(g_pre, N, x, static_cast<int>(ldx), dw, db, B, N,
                                                                                         static_cast<int>(K));
    PVB_LAUNCHED("linear_wgrad_small");
  }
  if (gx) {
    linear_dgrad_small_kernel<<<dim3(static_cast<unsigned>(ceil_div(K, 256LL)), B), 256, static_cast<size_t>(N) * sizeof(float), st>>>(
        g_pre, w, mask_gx_with_x ? x : nullptr, ldx, gx, ldgx, static_cast<int>(K), N);
    PVB_LAUNCHED("linear_dgrad_small");
  }
  return PVB200_OK;
}


/* y[b*ldy + d] = table[ids[b]][d]   (nn.Embedding(940, 16), model_sat_nwp.py:146-149,252-260) */
int pvb200_embedding_fwd_f32(const float* table, const int* ids, float* y, long long ldy, int B, int V, int D, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(table && ids && y && B > 0 && V > 0 && D > 0 && ldy >= D, "embedding_fwd: bad argument");
  embedding_fwd_kernel<<<ceil_div(B * D, 256), 256, 0, as_stream(stream)>>>(table, ids, y, ldy, B, V, D);
  PVB_LAUNCHED("embedding_fwd");
  return PVB200_OK;
}

/* dtable[v][d] = sum_{b: ids[b] == v} gy[b*ldgy + d]   (dense, deterministic: every entry is written) */
int pvb200_embedding_bwd_f32(const float* gy, long long ldgy, const int* ids, float* dtable, int B, int V, int D,
                             pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(gy && ids && dtable && B > 0 && V > 0 && D > 0 && ldgy >= D, "embedding_bwd: bad argument");
  embedding_bwd_kernel<<<ceil_div(V * D, 256), 256, 0, as_stream(stream)>>>(gy, ldgy, ids, dtable, B, V, D);
  PVB_LAUNCHED("embedding_bwd");
  return PVB200_OK;
}

/* dst[b*lddst + t*ns + s] = nan_to_num(src[b*sb + t*st + s], 0)  for t < nt, s < ns
 * (x.pv.pv_yield[:, :history_len+1, :ns].nan_to_num(0).reshape(B, -1), model_sat_nwp.py:207-232) */
int pvb200_history_flatten_f32(const float* src, long long sb, long long st, float* dst, long long lddst, int B, int nt, int ns,
                               pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(src && dst && B > 0 && nt > 0 && ns > 0 && lddst >= static_cast<long long>(nt) * ns, "history_flatten: bad argument");
  history_flatten_kernel<<<ceil_div(B * nt * ns, 256), 256, 0, as_stream(stream)>>>(src, sb, st, dst, lddst, B, nt, ns);
  PVB_LAUNCHED("history_flatten");
  return PVB200_OK;
}


/* y[b*ldy + n] = act(bias[n] + sum_s partial[s][b][n]): finishes a split-K forward (pvb200_fc1_fwd_bf16) as a Linear + ReLU */
int pvb200_linear_finish_f32(const float* partial, int S, const float* bias, float* y, long long ldy, int B, int N, int relu,
                             pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(partial && y && S > 0 && B > 0 && N > 0 && ldy >= N, "linear_finish: bad argument");
  linear_finish_kernel<<<ceil_div(B * N, 256), 256, 0, as_stream(stream)>>>(partial, S, bias, y, ldy, B, N, relu);
  PVB_LAUNCHED("linear_finish");
  return PVB200_OK;
}

/* g_pre[b][n] = gy[b][n] * (y[b][n] > 0) (y = NULL: no activation) and db[n] = sum_b g_pre[b][n]: the start of a Linear's
 * backward when the weight / data gradients are computed elsewhere (pvb200_fc1_{wgrad,dgrad}_bf16) */
int pvb200_linear_gpre_f32(const float* gy, long long ldgy, const float* y, long long ldy, float* g_pre, float* db, int B, int N,
                           pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(gy && g_pre && db && B > 0 && N > 0 && ldgy >= N, "linear_gpre: bad argument");
  cudaStream_t st = as_stream(stream);
  linear_gpre_kernel<<<ceil_div(B * N, 256), 256, 0, st>>>(gy, ldgy, y, ldy, g_pre, B, N);
  PVB_LAUNCHED("linear_gpre");
  linear_wgrad_small_kernel<<<ceil_div(N, 256), 256, 0, st>>>(g_pre, N, g_pre, N, db, db, B, N, 0);
  PVB_LAUNCHED("linear_bias_grad");
  return PVB200_OK;
}

}  // extern "C"
