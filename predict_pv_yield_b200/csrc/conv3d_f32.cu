// conv3d_f32.cu -- a3/a4/a11: Conv3d 3x3x3 (stride 1) forward and data-gradient in fp32 on the FMA pipe.
//
// Reference call sites: predict_pv_yield/models/conv3d/model.py:80-90 (layers), :117-120 (forward);
// the data gradient is the autograd of the same lines.
//
// fp32 mode must match torch to <= 1e-5 (normalised), which rules out TF32/bf16 tensor-core math,
// so this is a register-blocked direct convolution on CUDA cores, bounded by the FP32 FMA pipe
// (128 FMA/clk/SM).  Design (one CTA = 512 output positions x 32 output channels):
//   * "flattened pitch" tiling: an output plane (Ho x Wo) is addressed as q = ho*Wps + wo with
//     Wps = round_up(Wi + 2P, 4).  A tile is 512 consecutive q, so the input window a tile needs
//     for tap (kh,kw) is simply the contiguous run [q0 + kh*Wps + kw, ... + 512): no im2col, no
//     per-row halo logic.  Columns wo >= Wo are computed and discarded (<= 6 % waste).
//   * input channels are consumed 4 at a time through a 2-stage cp.async pipeline: 4 x 3 (kt) planes of
//     512 + 2*Wps + 8 floats are copied global->shared without registers while the previous chunk is in
//     the FMA loop (zero fill implements the padding of the data gradient; the int16 normalisation of
//     layer 0 is fused into a synchronous staging path), with the matching [4][27][32] weight slab
//     (pre-transposed once per call so the copy is contiguous).
//   * 256 threads = 64 position-threads x 4 channel groups; each thread owns 2 x 4 positions
//     x 8 output channels = 64 accumulators and performs 192 FMAs per (ci,kt,kh) from
//     4 x LDS.128 of input (conflict-free: 16 B lane stride) + 6 x LDS.128 of weights (warp
//     broadcast).
//   * epilogue fused: bias + ReLU (forward) or the ReLU mask of the layer below (data gradient).
//   * data gradient = the same kernel with P = 2, taps flipped and the weight roles swapped.
#include <type_traits>

#include "common.cuh"

namespace pvb {

constexpr int kQT = 512;       // output positions per CTA tile
constexpr int kCoT = 32;       // output channels per CTA tile
constexpr int kCC = 4;         // input channels per shared-memory chunk (double-buffered)
constexpr int kConvThreads = 256;

struct ConvArgs {
  const void* x;      // [B,Ci,Ti,Hi,Wi] fp32 or int16
  const float* mean;  // per input channel (int16 input only)
  const float* stdv;
  const float* wt;    // pre-transposed weights [Ci][27][CoPad]  (CoPad = round_up(Co,32))
  const float* bias;  // [Co] or null
  const float* mask;  // [B,Co,To,Ho,Wo] or null: y = mask > 0 ? y : 0
  float* y;           // [B,Co,To,Ho,Wo]
  int B, Ci, Ti, Hi, Wi;
  int Co, To, Ho, Wo;
  int P;      // implicit zero padding on every side of H,W (0: forward, 2: data gradient)
  int Pt;     // implicit zero padding on both sides of T (forward: the layer's time padding pt; data gradient: 2 - pt)
  int Wps;    // pitch of the flattened position space, multiple of 4, >= Wi + 2P
  int NP;     // staged positions per (ci,kt) plane = kQT + 2*Wps + 8
  int tiles_per_plane;
  int co_tiles;
  int CoPad;
  int relu;
  int pairs;  // Wi even (and the tensor 8-byte aligned): stage two positions per copy (a pair never straddles a row)
  int contig; // P == 0 and Wi a multiple of 4: the staged run IS a contiguous run of the input plane (16-byte copies)
};

// dw layout transform: wt[ci_role][tap][co_role] = w[co_role*s_co + ci_role*s_ci + (flip ? 26-tap : tap)]
__global__ void conv_weight_prep_kernel(const float* __restrict__ w, float* __restrict__ wt, int Ci, int Co, int CoPad,
                                        long long s_co, long long s_ci, int flip) {
  const int total = Ci * 27 * CoPad;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int co = idx % CoPad;
    const int tap = (idx / CoPad) % 27;
    const int ci = idx / (CoPad * 27);
    float v = 0.f;
    if (co < Co) v = w[co * s_co + ci * s_ci + (flip ? 26 - tap : tap)];
    wt[idx] = v;
  }
}

template <bool kI16>
__global__ void __launch_bounds__(kConvThreads, 2) conv3d_direct_f32_kernel(const ConvArgs a) {
  extern __shared__ __align__(16) float smem[];
  float* in_s = smem;                                  // [2][kCC][3][NP]
  float* w_s = smem + 2 * kCC * 3 * a.NP;              // [2][kCC][27][32]
  int* off_s = reinterpret_cast<int*>(w_s + 2 * kCC * 27 * kCoT);  // [NP] BYTE offset inside an input plane, or -1
  constexpr int kElemBytes = kI16 ? 2 : 4;

  const int tid = threadIdx.x;
  const int tp = tid & 63;   // position thread: positions {4tp..4tp+3} and {256+4tp..256+4tp+3}
  const int cg = tid >> 6;   // channel group: output channels 8cg..8cg+7 of the tile (warp-uniform)

  // tile decode
  int t = blockIdx.x;
  const int cot = t % a.co_tiles; t /= a.co_tiles;
  const int tile = t % a.tiles_per_plane; t /= a.tiles_per_plane;
  const int to = t % a.To;
  const int b = t / a.To;
  const int q0 = tile * kQT;
  const int co0 = cot * kCoT;

  // per-tile offset table: staged position i <-> padded position q0+i <-> input (hi,wi)
  for (int i = tid; i < a.NP; i += kConvThreads) {
    const int pos = q0 + i;
    const int hp = pos / a.Wps;
    const int wp = pos - hp * a.Wps;
    const int hi = hp - a.P, wi = wp - a.P;
    off_s[i] = (hi >= 0 && hi < a.Hi && wi >= 0 && wi < a.Wi) ? (hi * a.Wi + wi) * kElemBytes : -1;
  }

  // accumulators as channel PAIRS: acc2[j][i] = {channel 2j, channel 2j+1} of position i.  The FMA loop issues the packed
  // fma.rn.f32x2 of sm_100 (SASS FFMA2: two IEEE fp32 FMAs per issue slot, bit-identical to two fmaf) -- the scalar
  // loop was bound by the issue rate, not by the FMA pipe (ncu: FFMA 83 % of the instructions at 77 % issue utilisation).
  float2 acc2[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc2[j][i] = make_float2(0.f, 0.f);

  const long long plane_sz = static_cast<long long>(a.Hi) * a.Wi;
  const int warp = tid >> 5, lane = tid & 31;

  // ---- staging of one chunk of kCC input channels (+ its weight slab) into buffer `buf` ----
  // fp32 inputs go through cp.async (LDGSTS: global -> shared without registers, zero-fill for the padding), so the
  // NEXT chunk streams in while the FMA loop runs on the current one; int16 inputs are converted on the fly
  // (normalisation fused) and therefore staged synchronously.
  auto stage = [&](int c0, int buf) {
    float* in_b = in_s + buf * (kCC * 3 * a.NP);
    float* w_b = w_s + buf * (kCC * 27 * kCoT);
    for (int pl = warp; pl < kCC * 3; pl += kConvThreads / 32) {
      const int c = pl / 3, kt = pl - c * 3;
      const int ci = c0 + c;
      const int ti = to + kt - a.Pt;
      float* dst = in_b + pl * a.NP;
      const bool plane_ok = (ci < a.Ci) && (ti >= 0) && (ti < a.Ti);
      const long long base = plane_ok ? ((static_cast<long long>(b) * a.Ci + ci) * a.Ti + ti) * plane_sz : 0;
      if (kI16) {
        const char* src = reinterpret_cast<const char*>(static_cast<const int16_t*>(a.x) + base);
        const float m = plane_ok ? __ldg(a.mean + ci) : 0.f, s = plane_ok ? __ldg(a.stdv + ci) : 1.f;
        for (int i = lane; i < a.NP; i += 32) {
          const int o = off_s[i];
          dst[i] = (plane_ok && o >= 0) ? sat_norm(__ldg(reinterpret_cast<const int16_t*>(src + o)), m, s) : 0.f;
        }
      } else if (a.contig) {
        // Wps == Wi and no padding: staged position i is element q0 + i of the plane; 16-byte copies, zero fill past the end
        const float* src = static_cast<const float*>(a.x) + base + q0;
        const uint32_t d0 = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
        const int left = static_cast<int>(plane_sz) - q0;  // elements of the plane from q0 on (multiple of 4)
        for (int i = 4 * lane; i < a.NP; i += 128) {
          const bool ok = plane_ok && (i < left);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d0 + 4u * i), "l"(src + (ok ? i : 0)), "r"(ok ? 16 : 0)
                       : "memory");
        }
      } else if (a.pairs) {
        // positions 2i, 2i+1 of the staged run are neighbours in the same input row (Wps, q0, P and Wi are even):
        // one 8-byte copy, both valid or both padding
        const char* src = reinterpret_cast<const char*>(static_cast<const float*>(a.x) + base);
        const uint32_t d0 = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
        for (int i = 2 * lane; i < a.NP; i += 64) {
          const int o = off_s[i];
          const bool ok = plane_ok && (o >= 0);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d0 + 4u * i),
                       "l"(src + (ok ? static_cast<uint32_t>(o) : 0u)), "r"(ok ? 8 : 0)
                       : "memory");
        }
      } else {
        const char* src = reinterpret_cast<const char*>(static_cast<const float*>(a.x) + base);
        const uint32_t d0 = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
        for (int i = lane; i < a.NP; i += 32) {
          const int o = off_s[i];
          const bool ok = plane_ok && (o >= 0);
          // src-size 0 => 4 bytes of zeros are written (the implicit padding); the source pointer stays in range
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d0 + 4u * i),
                       "l"(src + (ok ? static_cast<uint32_t>(o) : 0u)), "r"(ok ? 4 : 0)
                       : "memory");
        }
      }
    }
    // weights: contiguous [kCC][27][32] slab out of wt[Ci][27][CoPad], 16 bytes per copy
    const uint32_t w0 = static_cast<uint32_t>(__cvta_generic_to_shared(w_b));
    const int rows_ok = (a.Ci - c0) * 27;  // rows (c, tap) of the slab that exist (the last chunk of Ci = 12k + r is short)
    const float* wsrc = a.wt + static_cast<long long>(c0) * 27 * a.CoPad + co0;
    for (int idx = tid; idx < kCC * 27 * kCoT / 4; idx += kConvThreads) {
      const int co4 = idx & (kCoT / 4 - 1);
      const int r = idx >> 3;  // c*27 + tap
      const bool ok = r < rows_ok;
      const float* src = wsrc + (ok ? r * a.CoPad + 4 * co4 : 0);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(w0 + 16u * idx), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // NG = position groups of the thread that hold real outputs: 2, or 1 when the second half of the (last) tile of a plane
  // lies beyond the plane -- its FMAs are skipped (CTA-uniform)
  auto compute = [&](int buf, int cmax, auto ng_tag) {
    constexpr int NG = decltype(ng_tag)::value;
    const float* in_b = in_s + buf * (kCC * 3 * a.NP);
    const float* w_b = w_s + buf * (kCC * 27 * kCoT);
    for (int c = 0; c < cmax; ++c) {
#pragma unroll
      for (int kt = 0; kt < 3; ++kt) {
        // a time plane that lies entirely in the zero padding of the data gradient contributes nothing (block-uniform)
        const int ti = to + kt - a.Pt;
        if (ti < 0 || ti >= a.Ti) continue;
        const float* ip = in_b + (c * 3 + kt) * a.NP + 4 * tp;
        const float* wp = w_b + (c * 27 + kt * 9) * kCoT + 8 * cg;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          // the 6 + 6 input values a kh row needs, each duplicated into both halves of a register pair
          float2 in2[2][6];
          {
            const float4 v0 = *reinterpret_cast<const float4*>(ip + kh * a.Wps);
            const float2 v1 = *reinterpret_cast<const float2*>(ip + kh * a.Wps + 4);
            in2[0][0] = make_float2(v0.x, v0.x); in2[0][1] = make_float2(v0.y, v0.y);
            in2[0][2] = make_float2(v0.z, v0.z); in2[0][3] = make_float2(v0.w, v0.w);
            in2[0][4] = make_float2(v1.x, v1.x); in2[0][5] = make_float2(v1.y, v1.y);
            if (NG == 2) {
              const float4 v2 = *reinterpret_cast<const float4*>(ip + kh * a.Wps + 256);
              const float2 v3 = *reinterpret_cast<const float2*>(ip + kh * a.Wps + 260);
              in2[1][0] = make_float2(v2.x, v2.x); in2[1][1] = make_float2(v2.y, v2.y);
              in2[1][2] = make_float2(v2.z, v2.z); in2[1][3] = make_float2(v2.w, v2.w);
              in2[1][4] = make_float2(v3.x, v3.x); in2[1][5] = make_float2(v3.y, v3.y);
            }
          }
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float4 w0 = *reinterpret_cast<const float4*>(wp + (kh * 3 + kw) * kCoT);
            const float4 w1 = *reinterpret_cast<const float4*>(wp + (kh * 3 + kw) * kCoT + 4);
            const float2 wv[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y),
                                  make_float2(w1.z, w1.w)};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                acc2[j][i] = __ffma2_rn(wv[j], in2[0][i + kw], acc2[j][i]);
                if (NG == 2) acc2[j][4 + i] = __ffma2_rn(wv[j], in2[1][i + kw], acc2[j][4 + i]);
              }
            }
          }
        }
      }
    }
  };

  // ---- main loop: 2-stage software pipeline over chunks of kCC input channels ----
  const int nchunk = (a.Ci + kCC - 1) / kCC;
  const bool half_tile = q0 + kQT / 2 >= (a.Ho - 1) * a.Wps + a.Wo;  // positions q0 + 256.. are all outside the plane
  __syncthreads();  // off_s visible
  stage(0, 0);
  for (int k = 0; k < nchunk; ++k) {
    if (k + 1 < nchunk) {
      stage((k + 1) * kCC, (k + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();  // chunk k landed for every thread
    if (half_tile)
      compute(k & 1, min(kCC, a.Ci - k * kCC), std::integral_constant<int, 1>());
    else
      compute(k & 1, min(kCC, a.Ci - k * kCC), std::integral_constant<int, 2>());
    __syncthreads();  // buffer k&1 may be overwritten by the stage issued in the next iteration
  }

  // ---- epilogue ----
  const long long oplane = static_cast<long long>(a.Ho) * a.Wo;
  const long long ochan = static_cast<long long>(a.To) * oplane;
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const int q = q0 + g * 256 + 4 * tp;
    const int ho = q / a.Wps;
    const int wo0 = q - ho * a.Wps;
    if (ho >= a.Ho) continue;
    const long long obase0 =
        ((static_cast<long long>(b) * a.Co + co0 + 8 * cg) * a.To + to) * oplane + static_cast<long long>(ho) * a.Wo + wo0;
    // data gradient: all 32 ReLU-mask loads of the group are issued before the first store (predicated, no branches between
    // them), so their latency is paid once per group instead of once per output row
    float mv[8][4];
    if (a.mask) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          mv[j][i] = (co0 + 8 * cg + j < a.Co && wo0 + i < a.Wo) ? __ldg(a.mask + obase0 + j * ochan + i) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = co0 + 8 * cg + j;
      if (co >= a.Co) continue;
      const float bv = a.bias ? __ldg(a.bias + co) : 0.f;
      const long long obase = obase0 + j * ochan;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (wo0 + i < a.Wo) {
          float v = ((j & 1) ? acc2[j >> 1][g * 4 + i].y : acc2[j >> 1][g * 4 + i].x) + bv;
          if (a.relu) v = (v < 0.f) ? 0.f : v;
          if (a.mask) v = (mv[j][i] > 0.f) ? v : 0.f;
          a.y[obase + i] = v;
        }
      }
    }
  }
}

static size_t conv_ws_bytes(int Ci_role, int Co_role) {
  return static_cast<size_t>(Ci_role) * 27 * round_up(Co_role, kCoT) * sizeof(float);
}

// role-level launcher shared by forward and data gradient
static int launch_conv(const void* x, bool i16, const float* mean, const float* stdv, const float* w, long long s_co,
                       long long s_ci, int flip, const float* bias, const float* mask, float* y, int B, int Ci, int Ti,
                       int Hi, int Wi, int Co, int P, int Pt, int relu, void* ws, size_t ws_bytes, cudaStream_t stream) {
  ConvArgs a;
  a.x = x; a.mean = mean; a.stdv = stdv; a.bias = bias; a.mask = mask; a.y = y;
  a.B = B; a.Ci = Ci; a.Ti = Ti; a.Hi = Hi; a.Wi = Wi; a.Co = Co;
  a.P = P; a.Pt = Pt;
  a.To = Ti + 2 * Pt - 2; a.Ho = Hi + 2 * P - 2; a.Wo = Wi + 2 * P - 2;
  PVB_REQUIRE(a.To > 0 && a.Ho > 0 && a.Wo > 0, "conv3d: input %dx%dx%d too small for a 3x3x3 kernel", Ti, Hi, Wi);
  a.Wps = round_up(Wi + 2 * P, 4);
  a.NP = kQT + 2 * a.Wps + 8;
  const int Qtot = (a.Ho - 1) * a.Wps + a.Wo;
  a.tiles_per_plane = ceil_div(Qtot, kQT);
  a.co_tiles = ceil_div(Co, kCoT);
  a.CoPad = a.co_tiles * kCoT;
  a.relu = relu;
  a.pairs = (!i16 && (Wi % 2 == 0) && (P % 2 == 0) && (reinterpret_cast<uintptr_t>(x) % 8 == 0)) ? 1 : 0;
  a.contig = (!i16 && P == 0 && (Wi % 4 == 0) && (static_cast<long long>(Hi) * Wi % 4 == 0) &&
              (reinterpret_cast<uintptr_t>(x) % 16 == 0)) ? 1 : 0;
  const size_t need = conv_ws_bytes(Ci, Co);
  if (ws == nullptr || ws_bytes < need) {
    set_error("conv3d: workspace too small (%zu < %zu bytes)", ws_bytes, need);
    return PVB200_ERR_WORKSPACE;
  }
  float* wt = static_cast<float*>(ws);
  a.wt = wt;
  {
    const int total = Ci * 27 * a.CoPad;
    conv_weight_prep_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(w, wt, Ci, Co, a.CoPad, s_co, s_ci, flip);
    PVB_LAUNCHED("conv_weight_prep");
  }
  const size_t smem = 2 * (static_cast<size_t>(kCC) * 3 * a.NP + kCC * 27 * kCoT) * sizeof(float) + a.NP * sizeof(int);
  PVB_REQUIRE(smem <= 227 * 1024, "conv3d: image width %d needs %zu B of shared memory (> 227 KB)", Wi, smem);
  const long long tiles = static_cast<long long>(B) * a.To * a.tiles_per_plane * a.co_tiles;
  PVB_REQUIRE(tiles <= 0x7fffffffLL, "conv3d: too many tiles");
  if (i16) {
    PVB_CUDA(cudaFuncSetAttribute(conv3d_direct_f32_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3d_direct_f32_kernel<true><<<static_cast<unsigned>(tiles), kConvThreads, smem, stream>>>(a);
  } else {
    PVB_CUDA(cudaFuncSetAttribute(conv3d_direct_f32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3d_direct_f32_kernel<false><<<static_cast<unsigned>(tiles), kConvThreads, smem, stream>>>(a);
  }
  PVB_LAUNCHED("conv3d_direct_f32");
  return PVB200_OK;
}

}  // namespace pvb

extern "C" {

size_t pvb200_conv3d_workspace_bytes(int Cin, int Cout) {
  // forward needs [Cin][27][pad32(Cout)], the data gradient [Cout][27][pad32(Cin)]
  const size_t f = pvb::conv_ws_bytes(Cin, Cout), d = pvb::conv_ws_bytes(Cout, Cin);
  return f > d ? f : d;
}

int pvb200_conv3d_fwd_f32_pad(const void* x, int x_is_i16, const float* mean, const float* std, const float* w,
                              const float* bias, float* y, void* workspace, size_t workspace_bytes, int B, int Cin, int Ti,
                              int Hi, int Wi, int Cout, int relu, int pad_t, int pad_hw, pvb200_stream_t stream) {
  PVB_REQUIRE(x && w && y, "conv3d_fwd: null pointer");
  PVB_REQUIRE(B > 0 && Cin > 0 && Cout > 0, "conv3d_fwd: bad shape");
  PVB_REQUIRE((pad_t == 0 || pad_t == 1) && (pad_hw == 0 || pad_hw == 1), "conv3d_fwd: padding (%d, %d) not in {0, 1}", pad_t, pad_hw);
  PVB_REQUIRE(!x_is_i16 || (mean && std), "conv3d_fwd: int16 input needs mean/std");
  return pvb::launch_conv(x, x_is_i16 != 0, mean, std, w, /*s_co=*/static_cast<long long>(Cin) * 27, /*s_ci=*/27,
                          /*flip=*/0, bias, nullptr, y, B, Cin, Ti, Hi, Wi, Cout, /*P=*/pad_hw, /*Pt=*/pad_t, relu, workspace,
                          workspace_bytes, pvb::as_stream(stream));
}

int pvb200_conv3d_fwd_f32_tpad(const void* x, int x_is_i16, const float* mean, const float* std, const float* w,
                               const float* bias, float* y, void* workspace, size_t workspace_bytes, int B, int Cin, int Ti,
                               int Hi, int Wi, int Cout, int relu, int pad_t, pvb200_stream_t stream) {
  return pvb200_conv3d_fwd_f32_pad(x, x_is_i16, mean, std, w, bias, y, workspace, workspace_bytes, B, Cin, Ti, Hi, Wi, Cout, relu,
                                   pad_t, 0, stream);
}

int pvb200_conv3d_fwd_f32(const void* x, int x_is_i16, const float* mean, const float* std, const float* w,
                          const float* bias, float* y, void* workspace, size_t workspace_bytes, int B, int Cin, int Ti,
                          int Hi, int Wi, int Cout, int relu, pvb200_stream_t stream) {
  return pvb200_conv3d_fwd_f32_tpad(x, x_is_i16, mean, std, w, bias, y, workspace, workspace_bytes, B, Cin, Ti, Hi, Wi, Cout, relu,
                                    0, stream);
}

int pvb200_conv3d_dgrad_f32(const float* gz, const float* w, const float* mask_src, float* gx, void* workspace,
                            size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout,
                            pvb200_stream_t stream) {
  return pvb200_conv3d_dgrad_f32_tpad(gz, w, mask_src, gx, workspace, workspace_bytes, B, Cin, Ti, Hi, Wi, Cout, 0, stream);
}

int pvb200_conv3d_dgrad_f32_tpad(const float* gz, const float* w, const float* mask_src, float* gx, void* workspace,
                                 size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout, int pad_t,
                                 pvb200_stream_t stream) {
  return pvb200_conv3d_dgrad_f32_pad(gz, w, mask_src, gx, workspace, workspace_bytes, B, Cin, Ti, Hi, Wi, Cout, pad_t, 0, stream);
}

int pvb200_conv3d_dgrad_f32_pad(const float* gz, const float* w, const float* mask_src, float* gx, void* workspace,
                                size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout, int pad_t, int pad_hw,
                                pvb200_stream_t stream) {
  PVB_REQUIRE(gz && w && gx, "conv3d_dgrad: null pointer");
  PVB_REQUIRE((pad_t == 0 || pad_t == 1) && (pad_hw == 0 || pad_hw == 1), "conv3d_dgrad: padding (%d, %d) not in {0, 1}", pad_t, pad_hw);
  PVB_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && Ti + 2 * pad_t > 2 && Hi + 2 * pad_hw > 2 && Wi + 2 * pad_hw > 2, "conv3d_dgrad: bad shape");
  // roles: the kernel's "input" is gz [B,Cout,Ti-2,Hi-2,Wi-2] padded by 2, its "output" is gx [B,Cin,Ti,Hi,Wi];
  // real weight w[co][ci][tap] is read as w[in-role=co][out-role=ci][26-tap]
  return pvb::launch_conv(gz, false, nullptr, nullptr, w, /*s_co (out-role=ci)=*/27,
                          /*s_ci (in-role=co)=*/static_cast<long long>(Cin) * 27, /*flip=*/1, nullptr, mask_src, gx, B,
                          /*Ci role=*/Cout, Ti + 2 * pad_t - 2, Hi + 2 * pad_hw - 2, Wi + 2 * pad_hw - 2, /*Co role=*/Cin,
                          /*P=*/2 - pad_hw, /*Pt=*/2 - pad_t,
                          /*relu=*/0, workspace, workspace_bytes, pvb::as_stream(stream));
}

}  // extern "C"
