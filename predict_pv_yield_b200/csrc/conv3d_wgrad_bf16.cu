// conv3d_wgrad_bf16.cu -- a11 in bf16: Conv3d weight gradient on the tensor cores (tcgen05, MN-major operands).
//
//   dw[co,ci,kt,kh,kw] = sum_{b,t,h,w} gz[b,co,t,h,w] * x[b,ci,t+kt,h+kh,w+kw]        (autograd of model.py:117-120)
//
// GEMM view per tap (kh,kw):  D[(p,ci), co] += sum_pos  X_p[ci, pos + kh*Wi + kw] * GZ[co, pos]
//   * the reduction (K) axis is the flattened position q = h*Wi + w of ONE (b,t) output plane.  Both operands are
//     read straight from the blocked layout [cg][pos][8ch] as MN-major SWIZZLE_NONE matrices (8 channels contiguous
//     per position, K stride 16 B, channel groups SBO apart): no transposition anywhere.
//   * gz must use the INPUT pitch Wi (tensor "gzw": [B][Cg][To][Ho][Wi][8], zero in the wrap columns wo >= Wo and in
//     the tail that rounds the plane up to a multiple of 128 positions), so x and gz share q and garbage never enters.
//   * M = 128 rows = 16 channel-group planes adjacent in shared memory (time plane p, channel group cg -> plane
//     p*Cgx + cg, all SBO apart): rows of p < 3 are the taps kt = p; the remaining planes are never loaded and their
//     rows are discarded (M = 96 is not a legal UMMA shape).  N = Cout.  The nine (kh,kw) taps are nine accumulators in TMEM
//     (9 x 32 columns) that live for the whole kernel; each persistent CTA streams its share of (b,t,q-chunk) tiles
//     and writes ONE partial at the end; a second kernel reduces the partials in fixed order (deterministic).
// Warp roles: warp 0 = bulk-copy producer (3-stage ring; lane i issues copy i of a tile -- a single thread issuing the
// 16 copies was the bottleneck), warps 1-3 = MMA issuers for kh = 0,1,2 (one thread cannot issue more than one
// tcgen05.mma per ~54 clk whatever its size; the three issuers own disjoint accumulators, so no ordering between
// them is needed), warps 4-7 = final TMEM read-out.
#include "common.cuh"
#include "tc_common.cuh"

namespace pvb {

constexpr int kWbThreads = 256;
constexpr int kWbQ = 128;  // positions (GEMM K) per tile
constexpr int kWbMaxStages = 4;

struct WgradBf16Args {
  const uint4* x;    // [B][Cgx][Ti][Hi][Wi]
  const uint4* gzw;  // [B][Cgo][To][QP]   QP = plane of Ho*Wi positions rounded up to a multiple of 128
  float* partial;    // [grid][9][M][N]
  int B, Cgx, Cgo, Ti, Hi, Wi, To;
  int M, N;      // M = 128 (16 channel-group planes: 16/Cgx time planes, 3 used), N = 8*Cgo
  int QP;        // padded gz plane size (positions)
  int chunks;    // QP / 128
  int NPOS;      // staged x positions per (plane, channel group)
  int nstage;
  long long tiles;  // B * To * chunks
  int pad_t;         // time padding of the layer: x plane of (t, kt) is t + kt - pad_t, zero outside [0, Ti)
  const uint4* zeros;  // NPOS 16-byte zeros (source of the copies for the planes of the time padding)
};

__global__ void __launch_bounds__(kWbThreads, 1) conv3d_wgrad_bf16_kernel(const WgradBf16Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [4]
  uint64_t* empty = full + kWbMaxStages;               // [4]
  uint64_t* done = empty + kWbMaxStages;               // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done + 1);
  uint8_t* stage_s = smem + 128;
  const uint32_t x_bytes = 16u * a.NPOS * 16u;  // 16 channel-group planes = M 128 rows (only 3 time planes are loaded)
  const uint32_t g_bytes = static_cast<uint32_t>(a.Cgo) * kWbQ * 16u;
  const uint32_t stage_bytes = x_bytes + g_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (9u * a.N <= 256u) ? 256u : 512u;

  // initialise the stages once: clamped copies at plane ends leave tails untouched, and stale bits must stay finite.
  // The never-loaded time plane p = 3 gets a constant 1.0 in its channel 0, so row 3*CiP of the (kh,kw) = (0,0)
  // accumulator becomes sum(gz) = the BIAS gradient, for free.
  {
    const uint32_t per_stage16 = stage_bytes / 16u;
    const uint32_t ones_lo = 3u * a.Cgx * a.NPOS, ones_hi = ones_lo + a.NPOS;  // 16-byte elements of plane (p=3, cg=0)
    for (uint32_t i = threadIdx.x; i < per_stage16 * a.nstage; i += kWbThreads) {
      const uint32_t e = i % per_stage16;
      reinterpret_cast<uint4*>(stage_s)[i] = (e >= ones_lo && e < ones_hi) ? make_uint4(0x00003f80u, 0, 0, 0) : make_uint4(0, 0, 0, 0);
    }
  }
  tc::fence_proxy_async();
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWbMaxStages; ++i) { tc::mbar_init(full + i, 1); tc::mbar_init(empty + i, 3); }
    tc::mbar_init(done, 3);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const long long g_begin = a.tiles * blockIdx.x / gridDim.x;
  const long long g_end = a.tiles * (blockIdx.x + 1) / gridDim.x;
  const long long x_plane = static_cast<long long>(a.Hi) * a.Wi;
  const uint32_t nstage = static_cast<uint32_t>(a.nstage);

  if (warp == 0) {
    // bulk-copy producer: lane i issues copy i of a tile (3 time planes x Cgx channel groups of x, then Cgo of gz)
    uint32_t seq = 0;
    const int ncopy_x = 3 * a.Cgx;
    for (long long g = g_begin; g < g_end; ++g, ++seq) {
      const int ch = static_cast<int>(g % a.chunks);
      const long long bt = g / a.chunks;
      const int t = static_cast<int>(bt % a.To);
      const int b = static_cast<int>(bt / a.To);
      const int q0 = ch * kWbQ;
      const uint32_t st = seq % nstage;
      const long long avail = x_plane - q0;
      const uint32_t npos = static_cast<uint32_t>(avail < a.NPOS ? (avail > 0 ? avail : 0) : a.NPOS);
      if (lane == 0) {
        tc::mbar_wait(empty + st, ((seq / nstage) & 1u) ^ 1u);
        tc::mbar_arrive_expect_tx(full + st, 3u * a.Cgx * npos * 16u + g_bytes);
      }
      __syncwarp();
      uint8_t* dst = stage_s + st * stage_bytes;
      // rows of the remaining (never loaded, zero-initialised) planes are computed and discarded
      if (lane < ncopy_x) {
        const int p = lane / a.Cgx, cg = lane - p * a.Cgx;
        const int tp = t + p - a.pad_t;
        const uint4* src = (tp >= 0 && tp < a.Ti) ? a.x + ((static_cast<long long>(b) * a.Cgx + cg) * a.Ti + tp) * x_plane + q0 : a.zeros;
        if (npos) tc::bulk_g2s(dst + static_cast<uint32_t>(lane) * a.NPOS * 16u, src, npos * 16u, full + st);
      } else if (lane < ncopy_x + a.Cgo) {
        const int cg = lane - ncopy_x;
        const uint4* src = a.gzw + ((static_cast<long long>(b) * a.Cgo + cg) * a.To + t) * a.QP + q0;
        tc::bulk_g2s(dst + x_bytes + static_cast<uint32_t>(cg) * kWbQ * 16u, src, kWbQ * 16u, full + st);
      }
    }
  } else if (warp <= 3) {
    const int kh = warp - 1;
    const bool leader = tc::elect_one();
    const uint32_t idesc = tc::umma_idesc(a.M, a.N, /*bf16*/ 1, /*A MN-major*/ 1, /*B MN-major*/ 1);
    // MN-major SWIZZLE_NONE: LBO = stride between 8-position K groups (128 B), SBO = stride between channel groups
    const uint32_t a_hi = ((static_cast<uint32_t>(a.NPOS) * 16u) >> 4) | (1u << 14);
    const uint32_t b_hi = ((kWbQ * 16u) >> 4) | (1u << 14);
    const uint32_t lo_lbo = (128u >> 4) << 16;
    const uint32_t stage16 = tc::smem_u32(stage_s) >> 4;
    const uint32_t wi = static_cast<uint32_t>(a.Wi);
    uint32_t seq = 0;
    for (long long g = g_begin; g < g_end; ++g, ++seq) {
      const uint32_t st = seq % nstage;
      tc::mbar_wait(full + st, (seq / nstage) & 1u);
      tc::tc_fence_after();
      const uint32_t xs16 = stage16 + st * (stage_bytes >> 4);
      const uint32_t gs16 = xs16 + (x_bytes >> 4);
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(kh * 3 + kw) * a.N;
#pragma unroll
        for (int ks = 0; ks < kWbQ / 16; ++ks) {
          const uint32_t a16 = xs16 + 16u * ks + kh * wi + kw;
          const uint32_t b16 = gs16 + 16u * ks;
          if (leader)
            tc::umma_bf16_lohi(d_tmem, lo_lbo | (a16 & 0x3fffu), a_hi, lo_lbo | (b16 & 0x3fffu), b_hi, idesc,
                         (seq | static_cast<uint32_t>(ks)) ? 1u : 0u);
        }
      }
      __syncwarp();
      if (leader) tc::umma_commit(empty + st);
      __syncwarp();
    }
    if (leader) tc::umma_commit(done);
    __syncwarp();
  } else {
    // read-out: D[tap][row][n] -> partial (rows of plane p = 3 are written too and ignored by the reduction)
    const int qd = warp & 3;
    tc::mbar_wait(done, 0);
    tc::tc_fence_after();
    const int row = qd * 32 + lane;
    float* part = a.partial + static_cast<long long>(blockIdx.x) * 9 * a.M * a.N;
    const bool has_tiles = g_end > g_begin;
    for (int tap = 0; tap < 9; ++tap) {
      for (int c0 = 0; c0 < a.N; c0 += 16) {
        uint32_t v[16];
        tc::tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + tap * a.N + c0, v);
        tc::tmem_ld_wait();
        if (row < a.M) {
          float4* dst = reinterpret_cast<float4*>(part + (static_cast<long long>(tap) * a.M + row) * a.N + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dst[j] = has_tiles ? make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                             __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]))
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// dw[co][ci][kt][kh][kw] = sum_cta partial[cta][kh*3+kw][kt*CiP + ci][co];  db[co] = sum_cta partial[cta][0][3*CiP][co]
__global__ void wgrad_bf16_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, float* __restrict__ db,
                                         int Co, int Ci, int CiP, int M, int N, int n_part) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Co * Ci * 27) {
    const int co = idx - Co * Ci * 27;
    if (db && co < Co) {
      const long long off = static_cast<long long>(3 * CiP) * N + co, stride = 9LL * M * N;
      float s = 0.f;
      for (int p = 0; p < n_part; ++p) s += partial[p * stride + off];
      db[co] = s;
    }
    return;
  }
  const int tap = idx % 27;
  const int ci = (idx / 27) % Ci;
  const int co = idx / (27 * Ci);
  const int kt = tap / 9, khw = tap % 9;
  const long long off = (static_cast<long long>(khw) * M + kt * CiP + ci) * N + co;
  const long long stride = 9LL * M * N;
  float s = 0.f;
  for (int p = 0; p < n_part; ++p) s += partial[p * stride + off];
  dw[idx] = s;
}

}  // namespace pvb

extern "C" {

/* gzw geometry: plane of Ho*Wi positions (INPUT pitch) rounded up to a multiple of 128 */
long long pvb200_conv3d_wgrad_bf16_gz_plane(int Hi, int Wi) {
  return pvb::round_up(static_cast<long long>(Hi - 2) * Wi, 128LL);
}

static const size_t kWbZeroBytes = 16 * 1024;  // zero page behind the partials (time-padding planes are copied from it)

size_t pvb200_conv3d_wgrad_bf16_workspace_bytes(int Cin, int Cout) {
  const int Cgx = 2 * pvb::ceil_div(Cin, 16), Cgo = 2 * pvb::ceil_div(Cout, 16);
  (void)Cgx;
  return static_cast<size_t>(160) * 9 * 128 * (8 * Cgo) * sizeof(float) + kWbZeroBytes;
}

int pvb200_conv3d_wgrad_bf16(const uint16_t* xb, const uint16_t* gzw, float* dw, float* db, void* workspace,
                             size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout,
                             pvb200_stream_t stream) {
  return pvb200_conv3d_wgrad_bf16_tpad(xb, gzw, dw, db, workspace, workspace_bytes, B, Cin, Ti, Hi, Wi, Cout, 0, stream);
}

int pvb200_conv3d_wgrad_bf16_tpad(const uint16_t* xb, const uint16_t* gzw, float* dw, float* db, void* workspace,
                                  size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout, int pad_t,
                                  pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(xb && gzw && dw, "conv3d_wgrad_bf16: null pointer");
  PVB_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && (pad_t == 0 || pad_t == 1) && Ti + 2 * pad_t > 2 && Hi > 2 && Wi > 2,
              "conv3d_wgrad_bf16: bad shape");
  PVB_REQUIRE(Cin <= 32 && Cout <= 32, "conv3d_wgrad_bf16: channels > 32 are not supported by the tensor-core path");
  WgradBf16Args a;
  a.x = reinterpret_cast<const uint4*>(xb);
  a.gzw = reinterpret_cast<const uint4*>(gzw);
  a.partial = static_cast<float*>(workspace);
  a.B = B; a.Cgx = 2 * ceil_div(Cin, 16); a.Cgo = 2 * ceil_div(Cout, 16);
  a.Ti = Ti; a.Hi = Hi; a.Wi = Wi; a.To = Ti + 2 * pad_t - 2;
  a.pad_t = pad_t;
  a.M = 128; a.N = 8 * a.Cgo;
  a.QP = static_cast<int>(pvb200_conv3d_wgrad_bf16_gz_plane(Hi, Wi));
  a.chunks = a.QP / kWbQ;
  a.NPOS = round_up(kWbQ + 2 * Wi + 2, 8);
  a.tiles = static_cast<long long>(B) * a.To * a.chunks;
  const size_t stage_bytes = static_cast<size_t>(16) * a.NPOS * 16 + static_cast<size_t>(a.Cgo) * kWbQ * 16;
  long long nstage = (227 * 1024 - 128) / static_cast<long long>(stage_bytes);
  if (nstage > kWbMaxStages) nstage = kWbMaxStages;
  PVB_REQUIRE(nstage >= 2, "conv3d_wgrad_bf16: width %d does not fit in shared memory", Wi);
  a.nstage = static_cast<int>(nstage);
  const int sms = sm_count();
  PVB_REQUIRE(sms > 0, "conv3d_wgrad_bf16: no CUDA device");
  long long grid = a.tiles < sms ? a.tiles : sms;
  if (grid > 160) grid = 160;
  const size_t need = static_cast<size_t>(grid) * 9 * a.M * a.N * sizeof(float) + kWbZeroBytes;
  PVB_REQUIRE(static_cast<size_t>(a.NPOS) * 16 <= kWbZeroBytes, "conv3d_wgrad_bf16: width %d too large for the zero page", Wi);
  if (!workspace || workspace_bytes < need) {
    set_error("conv3d_wgrad_bf16: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return PVB200_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  a.zeros = reinterpret_cast<const uint4*>(static_cast<uint8_t*>(workspace) + need - kWbZeroBytes);
  if (pad_t) PVB_CUDA(cudaMemsetAsync(const_cast<uint4*>(a.zeros), 0, kWbZeroBytes, st));
  const size_t smem = 128 + nstage * stage_bytes;
  PVB_CUDA(cudaFuncSetAttribute(conv3d_wgrad_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  conv3d_wgrad_bf16_kernel<<<static_cast<unsigned>(grid), kWbThreads, smem, st>>>(a);
  PVB_LAUNCHED("conv3d_wgrad_bf16");
  wgrad_bf16_reduce_kernel<<<ceil_div(Cout * Cin * 27 + Cout, 256), 256, 0, st>>>(a.partial, dw, db, Cout, Cin, 8 * a.Cgx, a.M,
                                                                                 a.N, (int)grid);
  PVB_LAUNCHED("wgrad_bf16_reduce");
  return PVB200_OK;
}

}  // extern "C"

// [B][Co][To][Ho][Wo] fp32 -> gzw blocked bf16 [B][Cg][To][QP][8] with the input pitch Wi = Wo + 2 (zeros elsewhere)
namespace pvb {
__global__ void nc_to_gzw_bf16_kernel(const float* __restrict__ gz, uint4* __restrict__ gzw, int Co, int Cg, int To, int Ho,
                                      int Wo, int QP, long long total) {
  const int Wi = Wo + 2;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int q = static_cast<int>(idx % QP);
    long long r = idx / QP;
    const int t = static_cast<int>(r % To); r /= To;
    const int cg = static_cast<int>(r % Cg);
    const long long b = r / Cg;
    const int ho = q / Wi, wo = q - ho * Wi;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (ho < Ho && wo < Wo) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = cg * 8 + j;
        if (c < Co) f[j] = gz[(((b * Co + c) * To + t) * Ho + ho) * Wo + wo];
      }
    }
    gzw[idx] = make_uint4(tc::pack_bf16(f[0], f[1]), tc::pack_bf16(f[2], f[3]), tc::pack_bf16(f[4], f[5]), tc::pack_bf16(f[6], f[7]));
  }
}
}  // namespace pvb

extern "C" int pvb200_nc_to_gzw_bf16(const float* gz, uint16_t* gzw, int B, int Cout, int To, int Ho, int Wo,
                                     pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(gz && gzw && B > 0 && Cout > 0 && To > 0 && Ho > 0 && Wo > 0, "nc_to_gzw_bf16: bad argument");
  const int Cg = 2 * ceil_div(Cout, 16);
  const int QP = static_cast<int>(pvb200_conv3d_wgrad_bf16_gz_plane(Ho + 2, Wo + 2));
  const long long total = static_cast<long long>(B) * Cg * To * QP;
  long long grid = ceil_div(total, 256LL);
  if (grid > 148 * 32) grid = 148 * 32;
  nc_to_gzw_bf16_kernel<<<static_cast<unsigned>(grid), 256, 0, as_stream(stream)>>>(gz, reinterpret_cast<uint4*>(gzw), Cout, Cg,
                                                                                     To, Ho, Wo, QP, total);
  PVB_LAUNCHED("nc_to_gzw_bf16");
  return PVB200_OK;
}
