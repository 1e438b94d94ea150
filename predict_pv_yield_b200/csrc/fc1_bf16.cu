// fc1_bf16.cu -- a6/a11 in bf16: the 128 x 1.1 M fc1 layer of the head as weight-streaming tensor-core GEMMs.
//
// Reference: self.fc1 = nn.Linear(cnn_output_size, 128) and F.relu(self.fc1(out)), predict_pv_yield/models/conv3d/
// model.py:92,125 (and their autograd).  fc1.weight is 99.9 % of the parameters; at per-GPU batches <~ 200 every pass
// over it is HBM-bound, so the bf16 path keeps a bf16 SHADOW of the fp32 master weight (half the bytes) in the layout
// the tensor cores consume directly:
//
//     W1s[kg][j][8]    kg = 8-channel group of the blocked feature index (kg = cg*THW + pos), j = output feature
//
// The conv stack's last activation stays blocked bf16, [b][kg][8], so its feature order matches kg without any
// re-layout; only the shadow is permuted (by fc1_make_shadow, once per optimiser step).
//   * forward   D[j, b]  = sum_k' W1s[j,k'] X[b,k']   A = W1s tile as K-major  SWIZZLE_NONE (LBO 2 KB, SBO 128 B)
//   * dgrad     D[k', b] = sum_j  W1s[k',j] G[b,j]    A = the SAME tile read as MN-major (SBO 2 KB, LBO 128 B)
//   * wgrad     D[j, k'] = sum_b  G[b,j] X[b,k']      B = X tile [kg][b][8] read as MN-major
// W1s tiles are single bulk copies (TMA engine); the small X / G operands are transposed into the canonical
// [k-group][row][8] form by the producer warp.  fp32 accumulation in TMEM; the forward writes split-K partials for the
// existing fused tail kernel, the data gradient fuses the ReLU mask and writes the two layouts the conv backward
// consumes, the weight gradient writes fp32 in the REFERENCE layout [128][K1] (what Adam / the all-reduce see).
#include <math.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pvb {

constexpr int kF1J = 128;       // fc1 output features (padded)
constexpr int kF1KG = 16;       // k-groups per tile  (128 feature columns)
constexpr int kF1Threads = 192; // warp 0 producer, warp 1 MMA, warps 2-5 epilogue
constexpr int kF1ThreadsWgrad = 448;  // + warps 6-13: the weight gradient's global-store warps

__device__ __forceinline__ uint32_t f2bf(float v) { return static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(v))); }

// ---- shadow: W1 fp32 [F1][K1] (reference layout, k = (cg*8+c8)*THW + pos) -> W1s bf16 [K1/8][128][8] -----------------
__global__ void __launch_bounds__(256)
fc1_make_shadow_kernel(const float* __restrict__ w, uint4* __restrict__ ws, int F1, long long THW, int Cg) {
  // one CTA = 32 positions x 32 features x one channel group; smem transpose so reads run along pos, writes along j
  __shared__ float tile[8][32][33];
  const long long pos0 = static_cast<long long>(blockIdx.x) * 32;
  const int j0 = blockIdx.y * 32;
  const int cg = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 warps
  for (int c8 = 0; c8 < 8; ++c8)
    for (int jj = ty; jj < 32; jj += 8) {
      const int j = j0 + jj;
      const long long pos = pos0 + tx;
      tile[c8][jj][tx] = (j < F1 && pos < THW) ? w[static_cast<long long>(j) * (Cg * 8 * THW) + (cg * 8 + c8) * THW + pos] : 0.f;
    }
  __syncthreads();
  for (int pp = ty; pp < 32; pp += 8) {
    const long long pos = pos0 + pp;
    if (pos >= THW) continue;
    const int j = j0 + tx;
    uint4 o;
    o.x = f2bf(tile[0][tx][pp]) | (f2bf(tile[1][tx][pp]) << 16);
    o.y = f2bf(tile[2][tx][pp]) | (f2bf(tile[3][tx][pp]) << 16);
    o.z = f2bf(tile[4][tx][pp]) | (f2bf(tile[5][tx][pp]) << 16);
    o.w = f2bf(tile[6][tx][pp]) | (f2bf(tile[7][tx][pp]) << 16);
    ws[(cg * THW + pos) * kF1J + j] = o;
  }
}

// ---- Adam on fc1.weight fused with the shadow refresh ------------------------------------------------------------
// Same arithmetic as adam_multi_kernel (loss_adam.cu; torch.optim.Adam single-tensor update, base_model.py:255-257),
// walked in the shadow's tile order so that the freshly updated weights are written out a second time as the bf16
// shadow [kg][128][8] in the same pass: the 565 MB master weight is not re-read by a separate conversion kernel.
struct AdamFc1Scalars {
  float beta1, beta2, one_minus_beta1, one_minus_beta2, eps, neg_step_size, bc2_sqrt, grad_scale;
};
// one CTA = 64 positions x 32 features x one channel group: 256-byte runs for p / g / m / v (float2 per lane) and
// 512-byte runs for the shadow (32 features x 16 B per position); the updated weights cross from "pos-major" to
// "feature-major" through a bf16 staging tile in shared memory (row padded to 66 elements: conflict-free both ways).
// All 32 x 4 loads of a thread are issued before the arithmetic so that enough bytes are in flight per SM.
constexpr int kAdamShPos = 64;
constexpr int kAdamShLd = 66;
// Row-sharded use (data parallel, optimiser state sharded by output feature): only the features [j_lo, j_hi) are
// updated and their bf16 copy goes to a contiguous shard [kg][j_hi - j_lo][8] (out_ld = j_hi - j_lo, out_j0 = j_lo)
// instead of the full shadow (out_ld = 128, out_j0 = 0).
__global__ void __launch_bounds__(256)
adam_fc1_shadow_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                       uint4* __restrict__ ws, int j_lo, int F1, int out_ld, int out_j0, long long THW, int Cg,
                       const AdamFc1Scalars s) {
  __shared__ __align__(16) uint16_t tile16[8 * 32 * kAdamShLd];  // [8 c8][32 j][66]
  const long long pos0 = static_cast<long long>(blockIdx.x) * kAdamShPos;
  const int j0 = j_lo + blockIdx.y * 32;
  const int cg = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long K1 = static_cast<long long>(Cg) * 8 * THW;
  const long long pos = pos0 + 2 * tx;
  // float2 accesses need an even element offset: rows start at multiples of THW, so only when THW is even; with an
  // odd THW (odd crops / odd sequence lengths) both positions of the lane go through scalar accesses instead
  const bool vec = (THW % 2 == 0) && (pos + 1 < THW);
  const bool ok1 = pos + 1 < THW;  // second position of the lane exists
  // not unrolled: the fully unrolled kernel was 120 KB of code and spent a third of its stall samples on
  // instruction fetch (ncu: no_instructions 31 %)
#pragma unroll 1
  for (int r = 0; r < 4; ++r) {
    const int jj = ty + 8 * r;
    const int j = j0 + jj;
    const bool ok = (j < F1) && (pos < THW);
    float2 pv[8], gv[8], mv[8], vv[8];
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      pv[c8] = gv[c8] = mv[c8] = vv[c8] = make_float2(0.f, 0.f);
      if (ok) {
        const long long i = static_cast<long long>(j) * K1 + (cg * 8 + c8) * THW + pos;
        if (vec) {
          gv[c8] = __ldcs(reinterpret_cast<const float2*>(g + i));
          mv[c8] = *reinterpret_cast<const float2*>(m + i);
          vv[c8] = *reinterpret_cast<const float2*>(v + i);
          pv[c8] = *reinterpret_cast<const float2*>(p + i);
        } else {
          gv[c8].x = g[i]; mv[c8].x = m[i]; vv[c8].x = v[i]; pv[c8].x = p[i];
          if (ok1) { gv[c8].y = g[i + 1]; mv[c8].y = m[i + 1]; vv[c8].y = v[i + 1]; pv[c8].y = p[i + 1]; }
        }
      }
    }
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      float ge[2] = {gv[c8].x * s.grad_scale, gv[c8].y * s.grad_scale};
      float me[2] = {mv[c8].x, mv[c8].y}, ve[2] = {vv[c8].x, vv[c8].y}, pe[2] = {pv[c8].x, pv[c8].y};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        me[e] = fmaf(ge[e] - me[e], s.one_minus_beta1, me[e]);
        ve[e] = fmaf(s.one_minus_beta2 * ge[e], ge[e], ve[e] * s.beta2);
        const float denom = __fdiv_rn(sqrtf(ve[e]), s.bc2_sqrt) + s.eps;
        pe[e] = fmaf(s.neg_step_size, __fdiv_rn(me[e], denom), pe[e]);
      }
      if (ok) {
        const long long i = static_cast<long long>(j) * K1 + (cg * 8 + c8) * THW + pos;
        if (vec) {
          *reinterpret_cast<float2*>(p + i) = make_float2(pe[0], pe[1]);
          *reinterpret_cast<float2*>(m + i) = make_float2(me[0], me[1]);
          *reinterpret_cast<float2*>(v + i) = make_float2(ve[0], ve[1]);
        } else {
          p[i] = pe[0]; m[i] = me[0]; v[i] = ve[0];
          if (ok1) { p[i + 1] = pe[1]; m[i + 1] = me[1]; v[i + 1] = ve[1]; }
        }
      } else {
        pe[0] = pe[1] = 0.f;
      }
      *reinterpret_cast<uint32_t*>(tile16 + (c8 * 32 + jj) * kAdamShLd + 2 * tx) = f2bf(pe[0]) | (f2bf(pe[1]) << 16);
    }
  }
  __syncthreads();
  for (int pp = ty; pp < kAdamShPos; pp += 8) {
    const long long q = pos0 + pp;
    if (q >= THW) continue;
    const uint16_t* src = tile16 + tx * kAdamShLd + pp;  // feature tx, position pp, channel stride 32*66
    uint4 o;
    o.x = src[0 * 32 * kAdamShLd] | (static_cast<uint32_t>(src[1 * 32 * kAdamShLd]) << 16);
    o.y = src[2 * 32 * kAdamShLd] | (static_cast<uint32_t>(src[3 * 32 * kAdamShLd]) << 16);
    o.z = src[4 * 32 * kAdamShLd] | (static_cast<uint32_t>(src[5 * 32 * kAdamShLd]) << 16);
    o.w = src[6 * 32 * kAdamShLd] | (static_cast<uint32_t>(src[7 * 32 * kAdamShLd]) << 16);
    if (j0 + tx < F1 || out_ld == kF1J) ws[(cg * THW + q) * out_ld + (j0 + tx - out_j0)] = o;
  }
}

// bf16 shards [nshards][KG][nrows][8] (all-gathered, one per rank) -> shadow [KG][128][8]; rows >= nshards*nrows untouched
__global__ void __launch_bounds__(256)
fc1_shadow_from_shards_kernel(const uint4* __restrict__ gathered, uint4* __restrict__ ws, int nshards, int nrows, long long KG) {
  const long long total = KG * nshards * nrows;
  const int F = nshards * nrows;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long kg = idx / F;
    const int j = static_cast<int>(idx - kg * F);
    const int r = j / nrows, jj = j - r * nrows;
    ws[kg * kF1J + j] = __ldcs(gathered + (static_cast<long long>(r) * KG + kg) * nrows + jj);
  }
}

struct Fc1Bf16Args {
  const uint4* ws;     // shadow [KG][128]
  const uint4* xb;     // features blocked [B][KG]
  const float* g1;     // [B][F1] fp32 (dgrad / wgrad)
  float* partial;      // forward: [S][B][F1]
  float* dw;           // wgrad: [F1][K1] fp32 reference layout
  uint4* gz_pad;       // dgrad out: [B][Cg][T+4][H+4][W+4] (zero border kept by the caller)
  uint4* gzw;          // dgrad out: [B][Cg][T][QP]
  int B, BP;           // batch, batch padded to a multiple of 16 (MMA N / K)
  int F1;
  long long KG;        // K1 / 8
  long long tiles;     // ceil(KG / 16)
  int Cg, T, H, W, QP;
  long long THW;
  int S;               // forward K splits
  long long tiles_per_split;
  int nst;             // pipeline stages that fit in shared memory (2..4)
  int nepi;            // weight gradient: staging buffers of the epilogue (1 or 2)
};

// transposing load of the X tile: smem [kgl][b][8] <- xb[b][kg0 + kgl]   (16 x BP chunks of 16 B).  Asynchronous
// (cp.async / LDGSTS, zero fill outside the tensor): the producer warp only issues the copies and lets each lane's
// completion arrive on the stage's `full` barrier, so it runs ahead by the depth of the ring instead of paying one
// memory latency per tile.
__device__ __forceinline__ void fc1_load_x_tile_async(uint4* dst, const Fc1Bf16Args& a, long long kg0, int lane, uint64_t* bar) {
  const uint32_t d0 = tc::smem_u32(dst);
  for (int b = lane; b < a.BP; b += 32) {
    const bool okb = b < a.B;
#pragma unroll
    for (int kgl = 0; kgl < kF1KG; ++kgl) {
      const long long kg = kg0 + kgl;
      const bool ok = okb && kg < a.KG;
      const uint4* src = a.xb + (ok ? static_cast<long long>(b) * a.KG + kg : 0);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d0 + static_cast<uint32_t>(kgl * a.BP + b) * 16u), "l"(src),
                   "r"(ok ? 16 : 0)
                   : "memory");
    }
  }
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// MODE 0: forward, 1: data gradient, 2: weight gradient
template <int MODE>
__global__ void __launch_bounds__(MODE == 2 ? kF1ThreadsWgrad : kF1Threads, 1) fc1_bf16_kernel(const Fc1Bf16Args a) {
  constexpr int kThr = MODE == 2 ? kF1ThreadsWgrad : kF1Threads;
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int kMaxSt = 4;
  const uint32_t NST = static_cast<uint32_t>(a.nst);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);   // [4]
  uint64_t* empty = full + kMaxSt;                      // [4]
  uint64_t* tfull = empty + kMaxSt;                     // [2]
  uint64_t* tempty = tfull + 2;                         // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* g_s = smem + 128;                            // G operand (dgrad / wgrad): BP*128 bf16 = up to 64 KB... sized BP*256 B
  const uint32_t g_bytes = (MODE == 0) ? 0u : static_cast<uint32_t>(a.BP) * 256u;
  const uint32_t w_bytes = (MODE == 2) ? 0u : kF1KG * kF1J * 16u;           // 32 KB W1s tile (not needed by wgrad)
  const uint32_t x_bytes = (MODE == 1) ? 0u : kF1KG * static_cast<uint32_t>(a.BP) * 16u;  // X tile (not needed by dgrad)
  const uint32_t stage_bytes = w_bytes + x_bytes;
  uint8_t* stage_s = g_s + ((g_bytes + 127u) & ~127u);
  // epilogue staging (transposes the accumulator tile so that global stores are wide and contiguous):
  //   dgrad: bf16 [BP][128 rows = (kgl, c8)]      wgrad: fp32 [128 cols = (c8, kgl)][129] (j fastest, padded)
  uint8_t* epi_s = stage_s + NST * stage_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // accumulator columns per tile: forward / dgrad N = BP, wgrad N = 128; double-buffered
  const uint32_t ncol = (MODE == 2) ? 128u : static_cast<uint32_t>(a.BP);
  const uint32_t tmem_cols = (2u * ncol <= 32u) ? 32u : (2u * ncol <= 64u) ? 64u : (2u * ncol <= 128u) ? 128u : (2u * ncol <= 256u) ? 256u : 512u;

  if (threadIdx.x == 0) {
    // full: lane 0's arrive (+ the W1s tile's bytes) and, where an X tile is staged, one cp.async arrival per producer lane
    for (int i = 0; i < kMaxSt; ++i) { tc::mbar_init(full + i, MODE == 1 ? 1 : 33); tc::mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(tfull + i, 1); tc::mbar_init(tempty + i, 4); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, tmem_cols);
  // stages start zeroed: a partial last tile leaves rows untouched, and stale bits must be finite (0 * NaN = NaN)
  for (uint32_t i = threadIdx.x; i < (NST * stage_bytes) / 16u; i += kThr) reinterpret_cast<uint4*>(stage_s)[i] = make_uint4(0, 0, 0, 0);
  tc::fence_proxy_async();
  if (MODE != 0) {
    // G operand, once per CTA.  dgrad: B-operand K-major [j/8][b][8 j]; wgrad: A-operand K-major [b/8][j][8 b]
    uint16_t* gs = reinterpret_cast<uint16_t*>(g_s);
    for (int idx = threadIdx.x; idx < a.BP * kF1J; idx += kThr) {
      const int b = idx / kF1J, j = idx - b * kF1J;
      const float v = (b < a.B && j < a.F1) ? a.g1[static_cast<long long>(b) * a.F1 + j] : 0.f;
      const int o = (MODE == 1) ? ((j >> 3) * a.BP + b) * 8 + (j & 7) : ((b >> 3) * kF1J + j) * 8 + (b & 7);
      gs[o] = static_cast<uint16_t>(f2bf(v));
    }
    tc::fence_proxy_async();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // tile range of this CTA
  long long t_begin, t_end;
  int split = 0;
  if (MODE == 0) {
    split = blockIdx.x;
    t_begin = split * a.tiles_per_split;
    t_end = min(a.tiles, t_begin + a.tiles_per_split);
  } else {
    t_begin = a.tiles * blockIdx.x / gridDim.x;
    t_end = a.tiles * (blockIdx.x + 1) / gridDim.x;
  }

  if (warp == 0) {
    // =============================== producer ===============================
    uint32_t seq = 0;
    for (long long t = t_begin; t < t_end; ++t, ++seq) {
      const uint32_t st = seq % NST;
      uint8_t* dst = stage_s + st * stage_bytes;
      const long long kg0 = t * kF1KG;
      if (lane == 0) tc::mbar_wait(empty + st, ((seq / NST) & 1u) ^ 1u);
      __syncwarp();
      if (MODE != 1) fc1_load_x_tile_async(reinterpret_cast<uint4*>(dst + w_bytes), a, kg0, lane, full + st);
      if (lane == 0) {
        if (MODE != 2) {
          const long long left = a.KG - kg0;
          const uint32_t nkg = static_cast<uint32_t>(left < kF1KG ? left : kF1KG);
          tc::mbar_arrive_expect_tx(full + st, nkg * kF1J * 16u);
          tc::bulk_g2s(dst, a.ws + kg0 * kF1J, nkg * kF1J * 16u, full + st);
        } else {
          tc::mbar_arrive(full + st);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const bool leader = tc::elect_one();
    const uint32_t hi_ver = 1u << 14;
    const uint32_t stage16 = tc::smem_u32(stage_s) >> 4;
    const uint32_t g16 = tc::smem_u32(g_s) >> 4;
    uint32_t seq = 0;
    if (MODE == 0) {
      // one accumulator for the whole K range of this split
      const uint32_t idesc = tc::umma_idesc(128, a.BP, 1, 0, 0);
      for (long long t = t_begin; t < t_end; ++t, ++seq) {
        const uint32_t st = seq % NST;
        tc::mbar_wait(full + st, (seq / NST) & 1u);
        tc::fence_proxy_async();  // the X tile was written by cp.async (generic proxy), the MMA reads through the async proxy
        tc::tc_fence_after();
        const uint32_t w16 = stage16 + st * (stage_bytes >> 4);
        const uint32_t x16 = w16 + (w_bytes >> 4);
#pragma unroll
        for (int ks = 0; ks < kF1KG / 2; ++ks) {
          // A = W1s tile K-major: k-groups 2 KB apart (LBO), 8-row groups 128 B apart (SBO)
          const uint32_t a_lo = ((2048u >> 4) << 16) | ((w16 + 2u * ks * 128u) & 0x3fffu);
          // B = X tile K-major: k-groups BP*16 B apart, 8-row groups 128 B apart
          const uint32_t b_lo = (((static_cast<uint32_t>(a.BP) * 16u) >> 4) << 16) | ((x16 + 2u * ks * a.BP) & 0x3fffu);
          if (leader) tc::umma_bf16_lohi(tmem_base, a_lo, (128u >> 4) | hi_ver, b_lo, (128u >> 4) | hi_ver, idesc, (seq | ks) ? 1u : 0u);
        }
        __syncwarp();
        if (leader) tc::umma_commit(empty + st);
        __syncwarp();
      }
      if (leader) tc::umma_commit(tfull);
      __syncwarp();
    } else {
      const uint32_t idesc = (MODE == 1) ? tc::umma_idesc(128, a.BP, 1, /*A MN*/ 1, /*B K*/ 0)
                                         : tc::umma_idesc(128, 128, 1, /*A K*/ 0, /*B MN*/ 1);
      const int nk = (MODE == 1) ? kF1J / 16 : a.BP / 16;  // K steps: over j (dgrad) or over b (wgrad)
      for (long long t = t_begin; t < t_end; ++t, ++seq) {
        const uint32_t st = seq % NST;
        const uint32_t acc = seq & 1u;
        tc::mbar_wait(full + st, (seq / NST) & 1u);
        tc::fence_proxy_async();
        tc::mbar_wait(tempty + acc, ((seq >> 1) & 1u) ^ 1u);
        tc::tc_fence_after();
        const uint32_t s16 = stage16 + st * (stage_bytes >> 4);
        const uint32_t d_tmem = tmem_base + acc * ncol;
        for (int ks = 0; ks < nk; ++ks) {
          uint32_t a_lo, a_hi, b_lo, b_hi;
          if (MODE == 1) {
            // A = W1s tile MN-major (M = k'): 8-j groups 128 B apart (LBO), k-groups 2 KB apart (SBO); step = 16 j = 256 B
            a_lo = ((128u >> 4) << 16) | ((s16 + 16u * ks) & 0x3fffu);
            a_hi = (2048u >> 4) | hi_ver;
            // B = G [j/8][b][8] K-major: k-groups BP*16 B apart (LBO), 8-b groups 128 B apart (SBO)
            b_lo = (((static_cast<uint32_t>(a.BP) * 16u) >> 4) << 16) | ((g16 + 2u * ks * a.BP) & 0x3fffu);
            b_hi = (128u >> 4) | hi_ver;
          } else {
            // A = G^T [b/8][j][8] K-major: k-groups 2 KB apart, 8-j groups 128 B apart; step = 16 b = 2 groups
            a_lo = ((2048u >> 4) << 16) | ((g16 + 2u * ks * 128u) & 0x3fffu);
            a_hi = (128u >> 4) | hi_ver;
            // B = X tile [kg][b][8] MN-major (N = k'): 8-b groups 128 B apart (LBO), k-groups BP*16 B apart (SBO)
            b_lo = ((128u >> 4) << 16) | ((s16 + 16u * ks) & 0x3fffu);
            b_hi = ((static_cast<uint32_t>(a.BP) * 16u) >> 4) | hi_ver;
          }
          if (leader) tc::umma_bf16_lohi(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, ks ? 1u : 0u);
        }
        __syncwarp();
        if (leader) { tc::umma_commit(empty + st); tc::umma_commit(tfull + acc); }
        __syncwarp();
      }
    }
  } else if (MODE == 2 && warp >= 6) {
    // =============================== weight gradient: global-store warps (6..13) ===============================
    // The accumulator tile arrives transposed in one of two staging buffers (named barriers 2/3 = buffer full,
    // 4/5 = buffer free, 128 staging + 256 storing threads); item = (j, c8, quad of positions): lanes = (quad, c8) for
    // one j -> eight 64-byte runs of dw[j][(cg*8+c8)*THW+pos] per store instruction.
    const int et = threadIdx.x - 192;  // 0..255
    const long long ntile = t_end - t_begin;
    const long long K1 = a.KG * 8;
    for (long long n = 0; n < ntile; ++n) {
      const long long t = t_begin + n;
      const int eb = (a.nepi == 2) ? static_cast<int>(n & 1) : 0;
      const float* es = reinterpret_cast<const float*>(epi_s) + eb * (128 * 129);
      if (eb) asm volatile("bar.sync 3, 384;" ::: "memory"); else asm volatile("bar.sync 2, 384;" ::: "memory");
      const long long kg0 = t * kF1KG;
      const int cg0 = static_cast<int>(kg0 / a.THW);
      const int pos0 = static_cast<int>(kg0 - cg0 * a.THW);
      const bool simple = (kg0 + kF1KG <= a.KG) && (pos0 + kF1KG <= a.THW) && (a.THW % 4 == 0) && (pos0 % 4 == 0);
      for (int item = et; item < 128 * 32; item += 256) {
        const int j = item >> 5, c8 = (item >> 2) & 7, q = item & 3;
        if (j >= a.F1) continue;
        const float* src = es + (c8 * kF1KG + 4 * q) * 129 + j;
        const float4 v4 = make_float4(src[0], src[129], src[258], src[387]);
        float* drow = a.dw + static_cast<long long>(j) * K1;
        if (simple) {
          __stcs(reinterpret_cast<float4*>(drow + (cg0 * 8 + c8) * a.THW + pos0 + 4 * q), v4);
        } else {
          const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
          for (int i = 0; i < 4; ++i) {
            const long long kg = kg0 + 4 * q + i;
            if (kg < a.KG) {
              const int cg = static_cast<int>(kg / a.THW);
              drow[(cg * 8 + c8) * a.THW + (kg - cg * a.THW)] = vv[i];
            }
          }
        }
      }
      if (n + a.nepi < ntile) {  // the staging warps wait for this buffer again nepi tiles later
        if (eb) asm volatile("bar.arrive 5, 384;" ::: "memory"); else asm volatile("bar.arrive 4, 384;" ::: "memory");
      }
    }
  } else {
    // =============================== epilogue (warps 2..5) ===============================
    const int qd = warp & 3;
    const int row = qd * 32 + lane;  // accumulator row owned by this thread
    if (MODE == 0) {
      tc::mbar_wait(tfull, 0);
      tc::tc_fence_after();
      const bool any = t_end > t_begin;
      for (int c0 = 0; c0 < a.BP; c0 += 16) {
        uint32_t v[16];
        tc::tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + c0, v);
        tc::tmem_ld_wait();
        if (row < a.F1) {
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (c0 + c < a.B) a.partial[(static_cast<long long>(split) * a.B + c0 + c) * a.F1 + row] = any ? __uint_as_float(v[c]) : 0.f;
        }
      }
    } else {
      uint32_t seq = 0;
      const long long plane_pad = static_cast<long long>(a.H + 4) * (a.W + 4);
      for (long long t = t_begin; t < t_end; ++t, ++seq) {
        const uint32_t acc = seq & 1u;
        const int et = threadIdx.x - 64;  // 0..127 within the epilogue warps
        // data gradient: this thread stores k-group (kg0 + et % 16) of the samples et / 16 + 8 i; the ReLU-mask
        // source (the activation itself) is fetched BEFORE waiting for the accumulator
        const long long dkg = t * kF1KG + (et & 15);
        const bool dkg_ok = dkg < a.KG;
        uint4 dmask[8];
        if (MODE == 1) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int b = (et >> 4) + 8 * i;
            dmask[i] = make_uint4(0, 0, 0, 0);
            if (b < a.B && dkg_ok && i * 8 < a.B) dmask[i] = __ldg(a.xb + static_cast<long long>(b) * a.KG + dkg);
          }
        }
        tc::mbar_wait(tfull + acc, (seq >> 1) & 1u);
        tc::tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + acc * ncol;
        if (MODE == 1) {
          // accumulator row = k' within the tile (k-group kg0 + row/8, channel row%8); columns = batch.
          // stage as bf16 [b][row] so that a (b, k-group) pair becomes one 16-byte vector
          uint16_t* es = reinterpret_cast<uint16_t*>(epi_s);
          for (int c0 = 0; c0 < a.BP; c0 += 16) {
            uint32_t v[16];
            tc::tmem_ld_32x16(taddr + c0, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) es[(c0 + c) * 128 + row] = static_cast<uint16_t>(f2bf(__uint_as_float(v[c])));
          }
          // accumulator fully read: hand it back to the MMA warp before the (slow) global phase
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(tempty + acc);
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (dkg_ok) {
            const int kgl = et & 15;
            const int cg = static_cast<int>(dkg / a.THW);
            const int pos = static_cast<int>(dkg - cg * a.THW);
            const int tt = pos / (a.H * a.W);
            const int hw = pos - tt * a.H * a.W;
            const int hh = hw / a.W, ww = hw - hh * a.W;
            // offsets of sample 0; a sample is Cg planes further
            const long long o_pad = (static_cast<long long>(cg) * (a.T + 4) + (tt + 2)) * plane_pad +
                                    static_cast<long long>(hh + 2) * (a.W + 4) + (ww + 2);
            const long long o_gzw = (static_cast<long long>(cg) * a.T + tt) * a.QP + static_cast<long long>(hh) * (a.W + 2) + ww;
            const long long s_pad = static_cast<long long>(a.Cg) * (a.T + 4) * plane_pad, s_gzw = static_cast<long long>(a.Cg) * a.T * a.QP;
            // keep a gradient lane where the activation (bf16, post-ReLU, finite) is > 0: sign clear and non-zero
            auto sel = [](uint32_t gv, uint32_t mv) {
              const uint32_t lo = ((mv & 0x8000u) == 0 && (mv & 0x7fffu) != 0) ? (gv & 0xffffu) : 0u;
              const uint32_t hi = ((mv & 0x80000000u) == 0 && (mv & 0x7fff0000u) != 0) ? (gv & 0xffff0000u) : 0u;
              return lo | hi;
            };
            for (int i0 = 0; i0 * 8 < a.B; i0 += 8) {
#pragma unroll
              for (int ii = 0; ii < 8; ++ii) {
                const int i = i0 + ii;
                const int b = (et >> 4) + 8 * i;
                if (b >= a.B) break;
                const uint4 g = *reinterpret_cast<const uint4*>(es + b * 128 + kgl * 8);
                const uint4 m = (i0 == 0) ? dmask[ii] : __ldg(a.xb + static_cast<long long>(b) * a.KG + dkg);
                const uint4 o = make_uint4(sel(g.x, m.x), sel(g.y, m.y), sel(g.z, m.z), sel(g.w, m.w));
                a.gz_pad[o_pad + b * s_pad] = o;
                if (a.gzw) a.gzw[o_gzw + b * s_gzw] = o;
              }
            }
          }
          asm volatile("bar.sync 2, 128;" ::: "memory");  // staging buffer free for the next tile
        } else {
          // accumulator row = feature j; columns = (kgl, c8).  Stage transposed as fp32 [(c8, kgl)][j] (row 129 floats:
          // conflict-free both ways) into the staging buffer seq % 2; the store warps write it out while this
          // warp group already stages the next tile.
          const uint32_t eb = (a.nepi == 2) ? (seq & 1u) : 0u;
          float* es = reinterpret_cast<float*>(epi_s) + eb * (128 * 129);
          if (seq >= static_cast<uint32_t>(a.nepi)) {
            if (eb) asm volatile("bar.sync 5, 384;" ::: "memory"); else asm volatile("bar.sync 4, 384;" ::: "memory");
          }
#pragma unroll
          for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tc::tmem_ld_32x32(taddr + c0, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int col = c0 + c, kgl = col >> 3, c8 = col & 7;
              es[(c8 * kF1KG + kgl) * 129 + row] = __uint_as_float(v[c]);
            }
          }
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(tempty + acc);
          if (eb) asm volatile("bar.arrive 3, 384;" ::: "memory"); else asm volatile("bar.arrive 2, 384;" ::: "memory");
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, tmem_cols);
}

static int fc1_bf16_fill(Fc1Bf16Args& a, int B, int F1, int Cg, int T, int H, int W) {
  PVB_REQUIRE(B > 0 && B <= 256, "fc1_bf16: batch %d not in 1..256 (split larger batches)", B);
  PVB_REQUIRE(F1 > 0 && F1 <= 128, "fc1_bf16: fc1_output_features=%d > 128 is not supported by the tensor-core path", F1);
  PVB_REQUIRE(Cg > 0 && T > 0 && H > 0 && W > 0, "fc1_bf16: bad feature geometry");
  a.B = B; a.BP = round_up(B, 16); a.F1 = F1;
  a.Cg = Cg; a.T = T; a.H = H; a.W = W;
  a.THW = static_cast<long long>(T) * H * W;
  a.KG = static_cast<long long>(Cg) * a.THW;
  a.tiles = ceil_div(a.KG, static_cast<long long>(kF1KG));
  a.QP = static_cast<int>(round_up(static_cast<long long>(H) * (W + 2), 128LL));
  return PVB200_OK;
}

template <int MODE>
static int fc1_bf16_launch(Fc1Bf16Args a, long long grid, cudaStream_t st) {
  const size_t g_bytes = (MODE == 0) ? 0 : static_cast<size_t>(a.BP) * 256;
  const size_t w_bytes = (MODE == 2) ? 0 : static_cast<size_t>(kF1KG) * kF1J * 16;
  const size_t x_bytes = (MODE == 1) ? 0 : static_cast<size_t>(kF1KG) * a.BP * 16;
  const size_t epi1 = (MODE == 1) ? static_cast<size_t>(a.BP) * 256 : (MODE == 2 ? static_cast<size_t>(128) * 129 * 4 : 0);
  const size_t fixed = 128 + round_up(g_bytes, static_cast<size_t>(128));
  const size_t cap = 227 * 1024;
  // as many pipeline stages (<= 4) and, for the weight gradient, epilogue staging buffers (<= 2) as fit
  a.nepi = (MODE == 2 && fixed + 2 * epi1 + 3 * (w_bytes + x_bytes) <= cap) ? 2 : 1;
  const size_t epi_bytes = a.nepi * epi1;
  PVB_REQUIRE(fixed + epi_bytes + 2 * (w_bytes + x_bytes) <= cap,
              "fc1_bf16: batch %d does not fit the tensor-core fc1 kernels' shared memory (use batches <= 128)", a.B);
  size_t nst = (cap - fixed - epi_bytes) / (w_bytes + x_bytes);
  a.nst = static_cast<int>(nst > 4 ? 4 : nst);
  const size_t smem = fixed + a.nst * (w_bytes + x_bytes) + epi_bytes;
  PVB_CUDA(cudaFuncSetAttribute(fc1_bf16_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fc1_bf16_kernel<MODE><<<static_cast<unsigned>(grid), MODE == 2 ? kF1ThreadsWgrad : kF1Threads, smem, st>>>(a);
  PVB_LAUNCHED("fc1_bf16");
  return PVB200_OK;
}

constexpr int kF1FwdSplits = 148;

}  // namespace pvb

extern "C" {

size_t pvb200_fc1_bf16_shadow_bytes(int Cg, int T, int H, int W) {
  return static_cast<size_t>(Cg) * T * H * W * pvb::kF1J * 16;
}

int pvb200_fc1_make_shadow_bf16(const float* w1, uint16_t* shadow, int F1, int Cg, int T, int H, int W,
                                pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(w1 && shadow && F1 > 0 && F1 <= kF1J && Cg > 0 && T > 0 && H > 0 && W > 0, "fc1_make_shadow: bad argument");
  const long long THW = static_cast<long long>(T) * H * W;
  dim3 grid(static_cast<unsigned>(ceil_div(THW, 32LL)), kF1J / 32, Cg);
  fc1_make_shadow_kernel<<<grid, 256, 0, as_stream(stream)>>>(w1, reinterpret_cast<uint4*>(shadow), F1, THW, Cg);
  PVB_LAUNCHED("fc1_make_shadow");
  return PVB200_OK;
}

int pvb200_adam_fc1_shadow(float* w1, const float* grad, float* exp_avg, float* exp_avg_sq, uint16_t* shadow, int F1, int Cg,
                           int T, int H, int W, float lr, float beta1, float beta2, float eps, int step, float grad_scale,
                           pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(w1 && grad && exp_avg && exp_avg_sq && shadow, "adam_fc1_shadow: null pointer");
  PVB_REQUIRE(F1 > 0 && F1 <= kF1J && Cg > 0 && T > 0 && H > 0 && W > 0 && step >= 1, "adam_fc1_shadow: bad argument");
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  AdamFc1Scalars s;
  s.beta1 = beta1; s.beta2 = beta2;
  s.one_minus_beta1 = static_cast<float>(1.0 - static_cast<double>(beta1));
  s.one_minus_beta2 = static_cast<float>(1.0 - static_cast<double>(beta2));
  s.eps = eps;
  s.neg_step_size = static_cast<float>(-(static_cast<double>(lr) / bc1));
  s.bc2_sqrt = static_cast<float>(sqrt(bc2));
  s.grad_scale = grad_scale;
  const long long THW = static_cast<long long>(T) * H * W;
  dim3 grid(static_cast<unsigned>(ceil_div(THW, static_cast<long long>(kAdamShPos))), kF1J / 32, Cg);
  adam_fc1_shadow_kernel<<<grid, 256, 0, as_stream(stream)>>>(w1, grad, exp_avg, exp_avg_sq, reinterpret_cast<uint4*>(shadow),
                                                                 0, F1, kF1J, 0, THW, Cg, s);
  PVB_LAUNCHED("adam_fc1_shadow");
  return PVB200_OK;
}

int pvb200_adam_fc1_shadow_rows(float* w1, const float* grad, float* exp_avg, float* exp_avg_sq, uint16_t* shard, int F1, int Cg,
                                int T, int H, int W, int row_lo, int nrows, float lr, float beta1, float beta2, float eps,
                                int step, float grad_scale, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(w1 && grad && exp_avg && exp_avg_sq && shard, "adam_fc1_shadow_rows: null pointer");
  PVB_REQUIRE(F1 > 0 && F1 <= kF1J && Cg > 0 && T > 0 && H > 0 && W > 0 && step >= 1, "adam_fc1_shadow_rows: bad argument");
  PVB_REQUIRE(row_lo >= 0 && nrows > 0 && nrows < kF1J && row_lo + nrows <= F1, "adam_fc1_shadow_rows: rows [%d, %d) outside [0, %d)",
              row_lo, row_lo + nrows, F1);
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  AdamFc1Scalars s;
  s.beta1 = beta1; s.beta2 = beta2;
  s.one_minus_beta1 = static_cast<float>(1.0 - static_cast<double>(beta1));
  s.one_minus_beta2 = static_cast<float>(1.0 - static_cast<double>(beta2));
  s.eps = eps;
  s.neg_step_size = static_cast<float>(-(static_cast<double>(lr) / bc1));
  s.bc2_sqrt = static_cast<float>(sqrt(bc2));
  s.grad_scale = grad_scale;
  const long long THW = static_cast<long long>(T) * H * W;
  dim3 grid(static_cast<unsigned>(ceil_div(THW, static_cast<long long>(kAdamShPos))), ceil_div(nrows, 32), Cg);
  adam_fc1_shadow_kernel<<<grid, 256, 0, as_stream(stream)>>>(w1, grad, exp_avg, exp_avg_sq, reinterpret_cast<uint4*>(shard),
                                                                 row_lo, row_lo + nrows, nrows, row_lo, THW, Cg, s);
  PVB_LAUNCHED("adam_fc1_shadow_rows");
  return PVB200_OK;
}

int pvb200_fc1_shadow_from_shards(const uint16_t* gathered, uint16_t* shadow, int nshards, int nrows, int Cg, int T, int H, int W,
                                  pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(gathered && shadow && nshards > 0 && nrows > 0 && nshards * nrows <= kF1J && Cg > 0 && T > 0 && H > 0 && W > 0,
              "fc1_shadow_from_shards: bad argument");
  const long long KG = static_cast<long long>(Cg) * T * H * W;
  fc1_shadow_from_shards_kernel<<<148 * 8, 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(gathered),
                                                                          reinterpret_cast<uint4*>(shadow), nshards, nrows, KG);
  PVB_LAUNCHED("fc1_shadow_from_shards");
  return PVB200_OK;
}

int pvb200_fc1_fwd_bf16_splits(void) { return pvb::kF1FwdSplits; }

/* partial[s][b][j] (s < pvb200_fc1_fwd_bf16_splits()) = split-K partial sums of xb . W1s^T; summed by pvb200_head_fwd_f32's
 * tail (pass partials as the head workspace with x = NULL) */
int pvb200_fc1_fwd_bf16(const uint16_t* xb, const uint16_t* shadow, float* partial, int B, int F1, int Cg, int T, int H,
                        int W, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(xb && shadow && partial, "fc1_fwd_bf16: null pointer");
  Fc1Bf16Args a{};
  int rc = fc1_bf16_fill(a, B, F1, Cg, T, H, W);
  if (rc) return rc;
  a.ws = reinterpret_cast<const uint4*>(shadow); a.xb = reinterpret_cast<const uint4*>(xb); a.partial = partial;
  a.S = kF1FwdSplits;
  a.tiles_per_split = ceil_div(a.tiles, static_cast<long long>(a.S));
  return fc1_bf16_launch<0>(a, a.S, as_stream(stream));
}

int pvb200_fc1_dgrad_bf16(const float* g1, const uint16_t* shadow, const uint16_t* xb, uint16_t* gz_pad, uint16_t* gzw,
                          int B, int F1, int Cg, int T, int H, int W, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(g1 && shadow && xb && gz_pad, "fc1_dgrad_bf16: null pointer");  // gzw may be NULL: not written
  Fc1Bf16Args a{};
  int rc = fc1_bf16_fill(a, B, F1, Cg, T, H, W);
  if (rc) return rc;
  a.ws = reinterpret_cast<const uint4*>(shadow); a.xb = reinterpret_cast<const uint4*>(xb); a.g1 = g1;
  a.gz_pad = reinterpret_cast<uint4*>(gz_pad); a.gzw = reinterpret_cast<uint4*>(gzw);
  const int sms = sm_count();
  PVB_REQUIRE(sms > 0, "fc1_dgrad_bf16: no CUDA device");
  return fc1_bf16_launch<1>(a, a.tiles < sms ? a.tiles : sms, as_stream(stream));
}

int pvb200_fc1_wgrad_bf16(const float* g1, const uint16_t* xb, float* dw1, int B, int F1, int Cg, int T, int H, int W,
                          pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(g1 && xb && dw1, "fc1_wgrad_bf16: null pointer");
  Fc1Bf16Args a{};
  int rc = fc1_bf16_fill(a, B, F1, Cg, T, H, W);
  if (rc) return rc;
  a.xb = reinterpret_cast<const uint4*>(xb); a.g1 = g1; a.dw = dw1;
  const int sms = sm_count();
  PVB_REQUIRE(sms > 0, "fc1_wgrad_bf16: no CUDA device");
  return fc1_bf16_launch<2>(a, a.tiles < sms ? a.tiles : sms, as_stream(stream));
}

}  // extern "C"
