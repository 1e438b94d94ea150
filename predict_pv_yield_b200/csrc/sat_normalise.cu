// sat_normalise.cu -- a1/a2: int16 satellite cube -> normalised fp32 / bf16, NCDHW in, NCDHW out.
//
// Reference arithmetic: predict_pv_yield/netcdf_dataset.py:96-101
//     sat = sat.astype(np.float32); sat = sat - SAT_MEAN; sat /= SAT_STD
// i.e. two separately rounded IEEE fp32 operations with a true division; reproduced with
// __fsub_rn / __fdiv_rn so the compiler can neither contract to FMA nor use a reciprocal.
//
// HBM-bound streaming kernel: one (b,c) plane per blockIdx.x so the per-channel constants are
// block-uniform; each thread moves 8 int16 per 128-bit load (4 loads in flight) and writes
// 2 x 128-bit (fp32) or 1 x 128-bit (bf16) stores.  Algorithmic bytes per element: 2 in + 4 (2) out.
#include "common.cuh"

namespace pvb {

constexpr int kNormThreads = 256;
constexpr int kNormUnroll = 4;

__device__ __forceinline__ uint4 ld_stream_u4(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ void norm8(const uint4& q, float mean, float stdv, float (&o)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[2 * i + 0] = sat_norm(static_cast<int16_t>(w[i] & 0xffffu), mean, stdv);
    o[2 * i + 1] = sat_norm(static_cast<int16_t>(w[i] >> 16), mean, stdv);
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  // RNE conversion of each half (cvt.rn.bf16.f32)
  __nv_bfloat16 a = __float2bfloat16_rn(lo), b = __float2bfloat16_rn(hi);
  return static_cast<uint32_t>(__bfloat16_as_ushort(a)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
}

template <bool kBf16>
__global__ void __launch_bounds__(kNormThreads)
sat_normalise_vec_kernel(const int16_t* __restrict__ x, void* __restrict__ y, const float* __restrict__ mean,
                         const float* __restrict__ stdv, int C, long long thw) {
  const long long plane = blockIdx.x;
  const int c = static_cast<int>(plane % C);
  const float m = __ldg(mean + c), s = __ldg(stdv + c);
  const long long nvec = thw >> 3;
  const uint4* xin = reinterpret_cast<const uint4*>(x + plane * thw);
  long long v0 = (static_cast<long long>(blockIdx.y) * kNormUnroll) * kNormThreads + threadIdx.x;
  const long long stride = static_cast<long long>(gridDim.y) * kNormUnroll * kNormThreads;
  for (; v0 < nvec; v0 += stride) {
    uint4 q[kNormUnroll];
#pragma unroll
    for (int u = 0; u < kNormUnroll; ++u) {
      const long long v = v0 + static_cast<long long>(u) * kNormThreads;
      if (v < nvec) q[u] = ld_stream_u4(xin + v);
    }
#pragma unroll
    for (int u = 0; u < kNormUnroll; ++u) {
      const long long v = v0 + static_cast<long long>(u) * kNormThreads;
      if (v < nvec) {
        float o[8];
        norm8(q[u], m, s, o);
        if (kBf16) {
          uint4 r = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                               pack_bf16x2(o[6], o[7]));
          __stcs(reinterpret_cast<uint4*>(static_cast<uint16_t*>(y) + plane * thw) + v, r);
        } else {
          float4* yo = reinterpret_cast<float4*>(static_cast<float*>(y) + plane * thw) + 2 * v;
          __stcs(yo, make_float4(o[0], o[1], o[2], o[3]));
          __stcs(yo + 1, make_float4(o[4], o[5], o[6], o[7]));
        }
      }
    }
  }
}

// scalar fallback for planes whose size is not a multiple of 8 (or misaligned pointers)
template <bool kBf16>
__global__ void __launch_bounds__(kNormThreads)
sat_normalise_scalar_kernel(const int16_t* __restrict__ x, void* __restrict__ y, const float* __restrict__ mean,
                            const float* __restrict__ stdv, int C, long long thw) {
  const long long plane = blockIdx.x;
  const int c = static_cast<int>(plane % C);
  const float m = __ldg(mean + c), s = __ldg(stdv + c);
  for (long long i = static_cast<long long>(blockIdx.y) * kNormThreads + threadIdx.x; i < thw;
       i += static_cast<long long>(gridDim.y) * kNormThreads) {
    const float o = sat_norm(x[plane * thw + i], m, s);
    if (kBf16)
      static_cast<uint16_t*>(y)[plane * thw + i] = __bfloat16_as_ushort(__float2bfloat16_rn(o));
    else
      static_cast<float*>(y)[plane * thw + i] = o;
  }
}

template <bool kBf16>
static int launch_normalise(const int16_t* x, void* y, const float* mean, const float* stdv, int B, int C,
                            long long thw, cudaStream_t stream) {
  PVB_REQUIRE(x && y && mean && stdv, "sat_normalise: null pointer");
  PVB_REQUIRE(B > 0 && C > 0 && thw > 0, "sat_normalise: bad shape B=%d C=%d thw=%lld", B, C, thw);
  const long long planes = static_cast<long long>(B) * C;
  PVB_REQUIRE(planes <= 0x7fffffffLL, "sat_normalise: too many planes");
  const bool vec = (thw % 8 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(y) % 16 == 0);
  if (vec) {
    const long long per_block = static_cast<long long>(kNormThreads) * kNormUnroll;
    long long gy = ceil_div(thw >> 3, per_block);
    if (gy > 64) gy = 64;
    dim3 grid(static_cast<unsigned>(planes), static_cast<unsigned>(gy));
    sat_normalise_vec_kernel<kBf16><<<grid, kNormThreads, 0, stream>>>(x, y, mean, stdv, C, thw);
  } else {
    long long gy = ceil_div(thw, static_cast<long long>(kNormThreads) * 8);
    if (gy > 64) gy = 64;
    dim3 grid(static_cast<unsigned>(planes), static_cast<unsigned>(gy));
    sat_normalise_scalar_kernel<kBf16><<<grid, kNormThreads, 0, stream>>>(x, y, mean, stdv, C, thw);
  }
  PVB_LAUNCHED("sat_normalise");
  return PVB200_OK;
}

}  // namespace pvb

extern "C" {

int pvb200_sat_normalise_f32(const int16_t* x, float* y, const float* mean, const float* std, int B, int C,
                             long long thw, pvb200_stream_t stream) {
  return pvb::launch_normalise<false>(x, y, mean, std, B, C, thw, pvb::as_stream(stream));
}

int pvb200_sat_normalise_bf16(const int16_t* x, uint16_t* y, const float* mean, const float* std, int B, int C,
                              long long thw, pvb200_stream_t stream) {
  return pvb::launch_normalise<true>(x, y, mean, std, B, C, thw, pvb::as_stream(stream));
}

}  // extern "C"

// ---- int16 [B][C][T][H][W] -> normalised blocked bf16 [B][Cg][T][H][W][8] (bf16 tensor-core path input) ------------
// Same two-rounding fp32 arithmetic, then RNE to bf16; channels >= C are zero.  One thread per (channel group,
// position): 8 coalesced 2-byte loads (one per channel plane), one 16-byte store.
namespace pvb {
// vector path: one thread = 8 consecutive positions x the 8 channels of one group: eight 128-bit streaming loads
// (one per channel plane).  A warp produces 256 consecutive positions = 4 KB of the blocked tensor; the 16-byte results
// go through a per-warp shared-memory tile so that every store instruction writes 512 contiguous bytes.
// Round-1 measurements (B = 32 cube, 140 MB of traffic): 0.077 ms = 1.8 TB/s with direct stores AND with the coalesced
// stores; a shared-memory table of the 10-bit values instead of the IEEE division: 0.094 ms.  The kernel is bound by
// its instruction stream (64 exact divisions + conversions per thread), not by memory; 2 % of the bf16 step.
__global__ void __launch_bounds__(256)
sat_normalise_blocked_vec_kernel(const int16_t* __restrict__ x, uint4* __restrict__ y, const float* __restrict__ mean,
                                 const float* __restrict__ stdv, int C, int Cg, long long thw, long long total8) {
  __shared__ uint4 tile[8][32 * 9];  // per warp: 256 positions, lane stride 9 x 16 B: conflict-free both ways
  const long long thw8 = thw >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // whole warps iterate together (the tile exchange needs all lanes): loop over warp-sized groups of 8-position units
  const long long nwarp_units = (total8 + 31) >> 5;
  for (long long wu = static_cast<long long>(blockIdx.x) * 8 + warp; wu < nwarp_units; wu += static_cast<long long>(gridDim.x) * 8) {
    const long long idx = (wu << 5) + lane;
    const bool ok = idx < total8;
    const long long p8 = ok ? idx % thw8 : 0;
    const long long r = ok ? idx / thw8 : 0;
    const int cg = static_cast<int>(r % Cg);
    const long long b = r / Cg;
    float f[8][8];  // [channel][position]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cg * 8 + j;
      if (ok && c < C) {
        const uint4 q = ld_stream_u4(reinterpret_cast<const uint4*>(x + (b * C + c) * thw) + p8);
        norm8(q, __ldg(mean + c), __ldg(stdv + c), f[j]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[j][i] = 0.f;
      }
    }
    // the 32 lanes of a warp cover one (b, cg) row only if thw8 is a multiple of 32; otherwise store directly
    const long long first = wu << 5;
    const bool same_row = (first / thw8) == ((first + 31 < total8 ? first + 31 : total8 - 1) / thw8) && (first + 31 < total8);
    if (same_row) {
      uint4* t = tile[warp];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        t[lane * 9 + i] = make_uint4(pack_bf16x2(f[0][i], f[1][i]), pack_bf16x2(f[2][i], f[3][i]),
                                                   pack_bf16x2(f[4][i], f[5][i]), pack_bf16x2(f[6][i], f[7][i]));
      __syncwarp();
      const long long fp8 = first % thw8, fr = first / thw8;
      uint4* dst = y + fr * thw + 8 * fp8;  // (b * Cg + cg) == fr
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int p = k * 32 + lane;  // position within the warp's 256
        dst[p] = t[p + (p >> 3)];
      }
      __syncwarp();
    } else if (ok) {
      uint4* dst = y + (b * Cg + cg) * thw + 8 * p8;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        dst[i] = make_uint4(pack_bf16x2(f[0][i], f[1][i]), pack_bf16x2(f[2][i], f[3][i]), pack_bf16x2(f[4][i], f[5][i]),
                            pack_bf16x2(f[6][i], f[7][i]));
    }
  }
}

__global__ void __launch_bounds__(256)
sat_normalise_blocked_kernel(const int16_t* __restrict__ x, uint4* __restrict__ y, const float* __restrict__ mean,
                             const float* __restrict__ stdv, int C, int Cg, long long thw, long long total) {
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pos = idx % thw;
    const long long r = idx / thw;
    const int cg = static_cast<int>(r % Cg);
    const long long b = r / Cg;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cg * 8 + j;
      f[j] = (c < C) ? sat_norm(__ldg(x + (b * C + c) * thw + pos), __ldg(mean + c), __ldg(stdv + c)) : 0.f;
    }
    uint4 o;
    o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]); o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
    y[idx] = o;
  }
}
}  // namespace pvb

extern "C" int pvb200_sat_normalise_blocked_bf16(const int16_t* x, uint16_t* yb, const float* mean, const float* std,
                                                 int B, int C, int T, int H, int W, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(x && yb && mean && std && B > 0 && C > 0 && T > 0 && H > 0 && W > 0, "sat_normalise_blocked_bf16: bad argument");
  const int Cg = 2 * ceil_div(C, 16);
  const long long thw = static_cast<long long>(T) * H * W;
  const long long total = static_cast<long long>(B) * Cg * thw;
  if (thw % 8 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0) {
    const long long total8 = total / 8;
    long long grid = ceil_div(total8, 256LL);
    if (grid > 148 * 16) grid = 148 * 16;
    sat_normalise_blocked_vec_kernel<<<static_cast<unsigned>(grid), 256, 0, as_stream(stream)>>>(
        x, reinterpret_cast<uint4*>(yb), mean, std, C, Cg, thw, total8);
  } else {
    long long grid = ceil_div(total, 256LL);
    if (grid > 148 * 32) grid = 148 * 32;
    sat_normalise_blocked_kernel<<<static_cast<unsigned>(grid), 256, 0, as_stream(stream)>>>(x, reinterpret_cast<uint4*>(yb), mean,
                                                                                              std, C, Cg, thw, total);
  }
  PVB_LAUNCHED("sat_normalise_blocked_bf16");
  return PVB200_OK;
}
