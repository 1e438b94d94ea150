// fc1_bf16x3.cu -- a6/a11 in fp32 MODE on the tensor cores: the 128 x 1.1 M fc1 layer of the head (forward, data gradient,
// weight gradient) as weight-streaming GEMMs over a THREE-WAY bf16 split of the fp32 operands (fp32-class accuracy).
//
// Reference: self.fc1 = nn.Linear(cnn_output_size, 128) and F.relu(self.fc1(out)), predict_pv_yield/models/conv3d/
// model.py:92,125 (and their autograd).
//
// The round-1 fp32 kernels (head_f32.cu) run these GEMMs on the FMA pipe and are FMA bound at batch 32 (12.8 FLOP per byte
// against a ridge of 11): 0.35 of HBM speed for a layer whose every pass is a 565 MB stream.  Here the fp32 tiles are
// fetched by tiled TMA loads (tensor maps), eight split warps turn every value into three bf16 pieces (v = b0 + b1 + b2
// exactly; conv3d_wgrad_bf16x3.cu has the arithmetic) written in the tensor cores' operand layout, and six of the nine
// piece products run as kind::f16 MMAs with fp32 accumulation in TMEM.  The split pass is where the layout work happens
// for free: the W pieces are stored [k/8][j][8 k], which the forward reads as a K-major A (rows j) and the data gradient
// as an MN-major A (rows k) -- no transposed copy of the weight exists anywhere; X pieces [k/8][b][8 k] serve the forward
// (K-major B) and the weight gradient (MN-major A); G pieces [j/8][b][8 j] the data gradient (K-major B) and the weight
// gradient (MN-major B).
//   forward   D[j, b]  = sum_k W[j,k] X[b,k]      k tiles of 64; split-K over the CTAs; the accumulator is folded into
//                                                 fp32 registers every two tiles (the tensor core rounds toward zero)
//   dgrad     D[k, b]  = sum_j W[j,k] G[b,j]      k tiles of 128, fetched and accumulated in four quarters of 32 features;
//                                                 the ReLU-mask source x arrives by a tensor-map load into the buffer the
//                                                 result is written to in place, which leaves as ONE tensor-map store
//   wgrad     D[k, j]  = sum_b X[b,k] G[b,j]      k tiles of 128; dW goes through a shared-memory tile [j][64 k] per warp
//                                                 pair and leaves as tensor-map stores
// (Register stores of 4 bytes to rows 4.4 MB apart cost a translation per access: the first version of the weight
// gradient spent 230 of its 304 us on them.)  Tiles are dealt round-robin over the CTAs.
// Warp roles (448 threads): warp 0 producer, warp 1 MMA issuer + TMEM owner, warps 2-9 split, warps 10-13 epilogue.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pvb {

constexpr int kFxThreads = 448;
constexpr int kFxSplitWarps = 8;
constexpr int kFxKT = 64;                 // k per raw tile (256 B per row: one TMA box row)
constexpr uint32_t kFxWS = 128 * 16 + 16; // stride between 8-k groups of the W pieces (padded: fewer bank conflicts)

struct FxArgs {
  int flags;        // tools only: 1 = epilogues skip their stores, 4 = the data gradient issues no MMAs, 8 = ... no split arithmetic
  const float* x;   // [B][K1]
  const float* g;   // [B][F1] gradient w.r.t. the fc1 pre-activation (dgrad, wgrad)
  float* out;       // fwd: partial [S][B][F1]; dgrad: gx [B][K1]; wgrad: dW [F1][K1]
  int B, BP, F1;
  int Btot;         // fwd: samples in the whole partial buffer (B is this launch's chunk of them)
  int accumulate;   // wgrad: add this launch's tile sums to dW (a later chunk of the batch) instead of storing them
  long long K1, tiles;  // tiles: k tiles of 64 (fwd) / 128 (dgrad, wgrad)
};

__device__ __forceinline__ void fx_tma_2d(void* dst_smem, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   tc::smem_u32(dst_smem)),
               "l"(reinterpret_cast<uint64_t>(tm)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

// shared -> global tile store through a tensor map (rows / columns outside the tensor are clipped by the hardware)
__device__ __forceinline__ void fx_tma_store_2d(const CUtensorMap* tm, const void* src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tm)),
               "r"(tc::smem_u32(src_smem)), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
// the same as an element-wise ADD to global memory (cp.reduce: the TMA engine reads, adds and writes back)
__device__ __forceinline__ void fx_tma_add_2d(const CUtensorMap* tm, const void* src_smem, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(tm)),
               "r"(tc::smem_u32(src_smem)), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void fx_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fx_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fx_pair_sync(int pair) { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); }

// v = b0 + b1 + b2 exactly, two values at a time (cvt.rn.bf16x2.f32 = one full-rate F2FP)
__device__ __forceinline__ void fx_split2(float v0, float v1, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(v1), "f"(v0));
  const float r0 = v0 - __uint_as_float(p0 << 16), r1 = v1 - __uint_as_float(p0 & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(r1), "f"(r0));
  const float s0 = r0 - __uint_as_float(p1 << 16), s1 = r1 - __uint_as_float(p1 & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(s1), "f"(s0));
}
__device__ __forceinline__ void fx_split8(const float4 lo4, const float4 hi4, uint4& q0, uint4& q1, uint4& q2) {
  fx_split2(lo4.x, lo4.y, q0.x, q1.x, q2.x);
  fx_split2(lo4.z, lo4.w, q0.y, q1.y, q2.y);
  fx_split2(hi4.x, hi4.y, q0.z, q1.z, q2.z);
  fx_split2(hi4.z, hi4.w, q0.w, q1.w, q2.w);
}

__device__ __forceinline__ void fx_mma(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate));
}
// the six kept products of the three-way split, smallest first: (0,2) (2,0) (1,1) (1,0) (0,1) (0,0); `first` = the very first
// MMA into the accumulator overwrites it
__device__ __forceinline__ void fx_mma6(uint32_t d, uint32_t a, uint32_t a_piece16, uint32_t a_hi, uint32_t b, uint32_t b_piece16,
                                        uint32_t b_hi, uint32_t idesc, bool first) {
  fx_mma(d, a, a_hi, b + 2u * b_piece16, b_hi, idesc, first ? 0u : 1u);
  fx_mma(d, a + 2u * a_piece16, a_hi, b, b_hi, idesc, 1u);
  fx_mma(d, a + a_piece16, a_hi, b + b_piece16, b_hi, idesc, 1u);
  fx_mma(d, a + a_piece16, a_hi, b, b_hi, idesc, 1u);
  fx_mma(d, a, a_hi, b + b_piece16, b_hi, idesc, 1u);
  fx_mma(d, a, a_hi, b, b_hi, idesc, 1u);
}

// raw fp32 tile [rows][8 * KG k] -> three bf16 pieces [k8][row][8 k] at `stride` bytes per 8-k group, for the 8-k groups
// kg0 .. kg0+KG-1 of the piece buffers.  Lane mapping: KG lanes cover one row (conflict-free 16-byte reads and writes).
template <int KG>
__device__ __forceinline__ void fx_split_tile(const uint8_t* raw, int rows, uint8_t* p0, uint32_t piece_bytes, uint32_t stride, int kg0,
                                              int tid, int nthreads) {
  for (int i = tid; i < rows * KG; i += nthreads) {
    const int r = i / KG, kg = i % KG;
    const float4* src = reinterpret_cast<const float4*>(raw + r * (KG * 32) + kg * 32);
    const int sw = (kg >> 2) & 1;  // lanes 4-7 of a row read their second half first: conflict-free 16-byte loads
    const float4 u = src[sw], v = src[sw ^ 1];
    uint4 q0, q1, q2;
    fx_split8(sw ? v : u, sw ? u : v, q0, q1, q2);
    uint8_t* dst = p0 + static_cast<uint32_t>(kg0 + kg) * stride + static_cast<uint32_t>(r) * 16u;
    *reinterpret_cast<uint4*>(dst) = q0;
    *reinterpret_cast<uint4*>(dst + piece_bytes) = q1;
    *reinterpret_cast<uint4*>(dst + 2u * piece_bytes) = q2;
  }
}

// G [B][F1] fp32 (global) -> three pieces [j8 (16)][b (BP)][8 j] at gs bytes per 8-j group; rows b >= B and columns j >= F1 zero
__device__ __forceinline__ void fx_split_g(const float* g, int B, int BP, int F1, uint8_t* p0, uint32_t piece_bytes, uint32_t gs, int tid,
                                           int nthreads) {
  for (int i = tid; i < 16 * BP; i += nthreads) {
    const int b = i % BP, j8 = i / BP;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int j = j8 * 8 + e;
      f[e] = (b < B && j < F1) ? __ldg(g + static_cast<long long>(b) * F1 + j) : 0.f;
    }
    uint4 q0, q1, q2;
    fx_split8(make_float4(f[0], f[1], f[2], f[3]), make_float4(f[4], f[5], f[6], f[7]), q0, q1, q2);
    uint8_t* dst = p0 + static_cast<uint32_t>(j8) * gs + static_cast<uint32_t>(b) * 16u;
    *reinterpret_cast<uint4*>(dst) = q0;
    *reinterpret_cast<uint4*>(dst + piece_bytes) = q1;
    *reinterpret_cast<uint4*>(dst + 2u * piece_bytes) = q2;
  }
}

// =====================================================================================================================
// forward: partial[cta][b][j] = sum over this CTA's k tiles of W[j,k] X[b,k]
// smem: [0,128) barriers | raw stage s (2): W [128][64] fp32, X [BP][64] fp32 | piece set s (2): W pieces 3 x 8 x kFxWS, X pieces
// =====================================================================================================================
__global__ void __launch_bounds__(kFxThreads, 1) fc1x3_fwd_kernel(const FxArgs a, const __grid_constant__ CUtensorMap tm_w,
                                                                  const __grid_constant__ CUtensorMap tm_x) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* rfull = reinterpret_cast<uint64_t*>(smem);  // [2] raw tile landed
  uint64_t* rempty = rfull + 2;                          // [2] raw tile converted
  uint64_t* pfull = rempty + 2;                          // [2] pieces written
  uint64_t* pempty = pfull + 2;                          // [2] pieces consumed
  uint64_t* dfull = pempty + 2;                          // [2] accumulator window complete
  uint64_t* dempty = dfull + 2;                          // [2] accumulator window folded into registers
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(dempty + 2);
  const int BP = a.BP;
  const uint32_t xs = static_cast<uint32_t>(BP) * 16u + 16u;  // stride between 8-k groups of the X pieces
  const uint32_t raw_w = 128u * 256u, raw_x = static_cast<uint32_t>(BP) * 256u;
  const uint32_t raw_stage = raw_w + raw_x;
  const uint32_t wp = 8u * kFxWS, xp = (8u * xs + 127u) & ~127u;  // one piece of a tile
  const uint32_t set_bytes = 3u * wp + 3u * xp;
  uint8_t* raw_s = smem + 128;
  uint8_t* set_s = raw_s + 2u * raw_stage;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(rfull + i, 1); tc::mbar_init(rempty + i, kFxSplitWarps);
      tc::mbar_init(pfull + i, kFxSplitWarps); tc::mbar_init(pempty + i, 1);
      tc::mbar_init(dfull + i, 1); tc::mbar_init(dempty + i, 4);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // tiles are dealt round-robin: at any moment the CTAs stream ONE contiguous span of every matrix row (DRAM page locality)
  const long long t_begin = blockIdx.x, t_end = a.tiles, t_step = gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t seq = 0;
      for (long long t = t_begin; t < t_end; t += t_step, ++seq) {
        const uint32_t st = seq & 1u;
        tc::mbar_wait(rempty + st, ((seq >> 1) & 1u) ^ 1u);
        tc::mbar_arrive_expect_tx(rfull + st, raw_stage);
        fx_tma_2d(raw_s + st * raw_stage, &tm_w, static_cast<int>(t * kFxKT), 0, rfull + st);
        fx_tma_2d(raw_s + st * raw_stage + raw_w, &tm_x, static_cast<int>(t * kFxKT), 0, rfull + st);
      }
    }
  } else if (warp == 1) {
    const bool leader = tc::elect_one();
    const uint32_t a_hi = (128u >> 4) | (1u << 14), b_hi = a_hi;                    // SBO = 128 B (8-row groups), K-major
    const uint32_t a_lbo = ((kFxWS >> 4) << 16), b_lbo = ((xs >> 4) << 16);          // LBO = stride between the 8-k groups
    const uint32_t set16 = tc::smem_u32(set_s) >> 4, set_b16 = set_bytes >> 4, wp16 = wp >> 4, xp16 = xp >> 4;
    const uint32_t idesc = tc::umma_idesc(128, BP, /*BF16*/ 1, 0, 0);
    uint32_t seq = 0;
    for (long long t = t_begin; t < t_end; t += t_step, ++seq) {
      const uint32_t st = seq & 1u;
      const uint32_t win = (seq >> 1) & 1u;  // accumulator window: two tiles
      if ((seq & 1u) == 0) {
        tc::mbar_wait(dempty + win, ((seq >> 2) & 1u) ^ 1u);
        tc::tc_fence_after();
      }
      tc::mbar_wait(pfull + st, (seq >> 1) & 1u);
      tc::tc_fence_after();
      const uint32_t d = tmem_base + win * static_cast<uint32_t>(BP);
      const uint32_t a0 = a_lbo | ((set16 + st * set_b16) & 0x3fffu);
      const uint32_t b0 = b_lbo | ((set16 + st * set_b16 + 3u * wp16) & 0x3fffu);
      if (leader) {
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16)
          fx_mma6(d, a0 + static_cast<uint32_t>(2 * k16) * (kFxWS >> 4), wp16, a_hi, b0 + static_cast<uint32_t>(2 * k16) * (xs >> 4), xp16, b_hi,
                  idesc, (seq & 1u) == 0 && k16 == 0);
        tc::umma_commit(pempty + st);
        if ((seq & 1u) == 1u || t + t_step >= t_end) tc::umma_commit(dfull + win);
      }
      __syncwarp();
    }
  } else if (warp < 2 + kFxSplitWarps) {
    const int tid = threadIdx.x - 64;
    uint32_t seq = 0;
    for (long long t = t_begin; t < t_end; t += t_step, ++seq) {
      const uint32_t st = seq & 1u;
      tc::mbar_wait(pempty + st, ((seq >> 1) & 1u) ^ 1u);
      tc::mbar_wait(rfull + st, (seq >> 1) & 1u);
      uint8_t* set = set_s + st * set_bytes;
      fx_split_tile<8>(raw_s + st * raw_stage, 128, set, wp, kFxWS, 0, tid, kFxSplitWarps * 32);
      fx_split_tile<8>(raw_s + st * raw_stage + raw_w, BP, set + 3u * wp, xp, xs, 0, tid, kFxSplitWarps * 32);
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) { tc::mbar_arrive(pfull + st); tc::mbar_arrive(rempty + st); }
    }
  } else {
    const int qd = warp & 3, row = qd * 32 + lane;  // row = output feature j
    float acc[64];
#pragma unroll
    for (int c = 0; c < 64; ++c) acc[c] = 0.f;
    const long long ntile = (t_end - t_begin + t_step - 1) / t_step;
    const long long nwin = (ntile + 1) / 2;
    for (long long w = 0; w < nwin; ++w) {
      const uint32_t win = static_cast<uint32_t>(w & 1);
      tc::mbar_wait(dfull + win, static_cast<uint32_t>((w >> 1) & 1));
      tc::tc_fence_after();
      const uint32_t addr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + win * static_cast<uint32_t>(BP);
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        if (c0 < BP) {
          uint32_t v[16];
          tc::tmem_ld_32x16(addr + c0, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[c0 + c] += __uint_as_float(v[c]);
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(dempty + win);
    }
    if (row < a.F1) {
      float* dst = a.out + static_cast<long long>(blockIdx.x) * a.Btot * a.F1 + row;
#pragma unroll
      for (int b = 0; b < 64; ++b)
        if (b < a.B) dst[static_cast<long long>(b) * a.F1] = acc[b];
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================================
// data gradient: gx[b][k] = (x[b][k] > 0) * sum_j G[b][j] W[j][k], k tiles of 128.  A tile is fetched and accumulated in four
// quarters of 32 features j (the split of one quarter runs under the MMAs of another).  The x tile [b][128 k] of the ReLU
// mask arrives by a tensor-map load into the buffer the result is written to IN PLACE, which then leaves as one
// tensor-map store: 4-byte accesses from the registers to 32 rows 4.4 MB apart cost a translation per access and ran at a
// third of the speed.
// smem: [0,256) barriers | raw stage s (4): W [32 j][128 k] fp32 | W piece set s (3): 3 x 16 x kFxDS | G pieces 3 x 16 x gs |
//       x / gx tile (2): [BP][128 k] fp32
// =====================================================================================================================
constexpr uint32_t kFxDS = 32 * 16 + 16;  // stride between the 8-k groups of a quarter tile's W pieces
constexpr uint32_t kFxDSets = 3;          // piece sets in flight between the split warps and the MMA issuer

__global__ void __launch_bounds__(kFxThreads, 1) fc1x3_dgrad_kernel(const FxArgs a, const __grid_constant__ CUtensorMap tm_w,
                                                                    const __grid_constant__ CUtensorMap tm_x,
                                                                    const __grid_constant__ CUtensorMap tm_out) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* rfull = reinterpret_cast<uint64_t*>(smem);  // [4] raw quarter landed
  uint64_t* rempty = rfull + 4;                          // [4]
  uint64_t* pfull = rempty + 4;                          // [3] pieces written
  uint64_t* pempty = pfull + 3;                          // [3]
  uint64_t* dfull = pempty + 3;                          // [2] accumulator complete
  uint64_t* dempty = dfull + 2;                          // [2]
  uint64_t* xfull = dempty + 2;                          // [2] x tile landed
  uint64_t* xempty = xfull + 2;                          // [2] result tile read by its store
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(xempty + 2);
  const int BP = a.BP;
  const uint32_t gs = static_cast<uint32_t>(BP) * 16u + 16u;
  const uint32_t raw_w = 32u * 512u;
  const uint32_t wp = 16u * kFxDS, gp = (16u * gs + 127u) & ~127u;
  const uint32_t set_bytes = 3u * wp;
  const uint32_t xo_bytes = static_cast<uint32_t>(BP) * 512u;
  uint8_t* raw_s = smem + 256;
  uint8_t* w_s = raw_s + 4u * raw_w;
  uint8_t* g_s = w_s + kFxDSets * set_bytes;
  uint8_t* xo_s = g_s + 3u * gp;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { tc::mbar_init(rfull + i, 1); tc::mbar_init(rempty + i, kFxSplitWarps); }
    for (int i = 0; i < 3; ++i) { tc::mbar_init(pfull + i, kFxSplitWarps); tc::mbar_init(pempty + i, 1); }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(dfull + i, 1); tc::mbar_init(dempty + i, 4);
      tc::mbar_init(xfull + i, 1); tc::mbar_init(xempty + i, 1);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, 512);
  if (warp >= 2 && warp < 2 + kFxSplitWarps) fx_split_g(a.g, a.B, BP, a.F1, g_s, gp, gs, threadIdx.x - 64, kFxSplitWarps * 32);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // tiles are dealt round-robin: at any moment the CTAs stream ONE contiguous span of every matrix row (DRAM page locality)
  const long long t_begin = blockIdx.x, t_end = a.tiles, t_step = gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t seq = 0, tile = 0;
      for (long long t = t_begin; t < t_end; t += t_step, ++tile) {
        const uint32_t xb = tile & 1u;
        tc::mbar_wait(xempty + xb, ((tile >> 1) & 1u) ^ 1u);
        tc::mbar_arrive_expect_tx(xfull + xb, xo_bytes);
        fx_tma_2d(xo_s + xb * xo_bytes, &tm_x, static_cast<int>(t * 128), 0, xfull + xb);
        for (int q = 0; q < 4; ++q, ++seq) {
          const uint32_t st = seq & 3u;
          tc::mbar_wait(rempty + st, ((seq >> 2) & 1u) ^ 1u);
          tc::mbar_arrive_expect_tx(rfull + st, raw_w);
          fx_tma_2d(raw_s + st * raw_w, &tm_w, static_cast<int>(t * 128), q * 32, rfull + st);
        }
      }
    }
  } else if (warp == 1) {
    const bool leader = tc::elect_one();
    // A = W pieces read MN-major (M = k): 8-k groups kFxDS apart (SBO), the 8-j K groups 128 B apart (LBO)
    const uint32_t a_hi = (kFxDS >> 4) | (1u << 14), a_lbo = ((128u >> 4) << 16);
    // B = G pieces K-major (N = b, K = j): 8-row groups 128 B apart (SBO), 8-j groups gs apart (LBO)
    const uint32_t b_hi = (128u >> 4) | (1u << 14), b_lbo = ((gs >> 4) << 16);
    const uint32_t w16 = tc::smem_u32(w_s) >> 4, g16 = tc::smem_u32(g_s) >> 4, wp16 = wp >> 4, gp16 = gp >> 4, set16 = set_bytes >> 4;
    const uint32_t idesc = tc::umma_idesc(128, BP, /*BF16*/ 1, /*A MN-major*/ 1, /*B K-major*/ 0);
    uint32_t tile = 0, ps = 0, pphase = 0;
    for (long long t = t_begin; t < t_end; t += t_step, ++tile) {
      const uint32_t win = tile & 1u;
      tc::mbar_wait(dempty + win, ((tile >> 1) & 1u) ^ 1u);
      for (uint32_t q = 0; q < 4; ++q) {
        tc::mbar_wait(pfull + ps, pphase);
        tc::tc_fence_after();
        const uint32_t d = tmem_base + win * static_cast<uint32_t>(BP);
        const uint32_t a0 = a_lbo | ((w16 + ps * set16) & 0x3fffu), b0 = b_lbo | (g16 & 0x3fffu);
        if (leader) {
#pragma unroll
          for (uint32_t k16 = 0; k16 < 2; ++k16)  // K = j: 16 j per MMA = 256 B of the W pieces' row axis, two 8-j groups of G
            if (!(a.flags & 4))
              fx_mma6(d, a0 + k16 * 16u, wp16, a_hi, b0 + 2u * (q * 2u + k16) * (gs >> 4), gp16, b_hi, idesc, q == 0 && k16 == 0);
          tc::umma_commit(pempty + ps);
          if (q == 3) tc::umma_commit(dfull + win);
        }
        __syncwarp();
        if (++ps == kFxDSets) { ps = 0; pphase ^= 1u; }
      }
    }
  } else if (warp < 2 + kFxSplitWarps) {
    const int tid = threadIdx.x - 64;
    uint32_t seq = 0, ps = 0, pphase = 0;
    for (long long t = t_begin; t < t_end; t += t_step)
      for (int q = 0; q < 4; ++q, ++seq) {
        const uint32_t st = seq & 3u;
        tc::mbar_wait(pempty + ps, pphase ^ 1u);
        tc::mbar_wait(rfull + st, (seq >> 2) & 1u);
        if (!(a.flags & 8)) fx_split_tile<16>(raw_s + st * raw_w, 32, w_s + ps * set_bytes, wp, kFxDS, 0, tid, kFxSplitWarps * 32);
        tc::fence_proxy_async();
        __syncwarp();
        if (lane == 0) { tc::mbar_arrive(pfull + ps); tc::mbar_arrive(rempty + st); }
        if (++ps == kFxDSets) { ps = 0; pphase ^= 1u; }
      }
  } else {
    const int qd = warp & 3, row = qd * 32 + lane;  // row = k within the tile
    uint32_t tile = 0;
    for (long long t = t_begin; t < t_end; t += t_step, ++tile) {
      const uint32_t win = tile & 1u;
      float* xo = reinterpret_cast<float*>(xo_s + win * xo_bytes);
      tc::mbar_wait(xfull + win, (tile >> 1) & 1u);
      tc::mbar_wait(dfull + win, (tile >> 1) & 1u);
      tc::tc_fence_after();
      const uint32_t addr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + win * static_cast<uint32_t>(BP);
#pragma unroll 1
      for (int c0 = 0; c0 < BP; c0 += 16) {
        uint32_t v[16];
        tc::tmem_ld_32x16(addr + c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          float* p = xo + (c0 + c) * 128 + row;
          *p = (*p > 0.f) ? __uint_as_float(v[c]) : 0.f;
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(dempty + win);
      tc::fence_proxy_async();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 2 + kFxSplitWarps && lane == 0) {
        if (!(a.flags & 1)) fx_tma_store_2d(&tm_out, xo, static_cast<int>(t * 128), 0);
        fx_store_wait_read();
        tc::mbar_arrive(xempty + win);
      }
    }
    if (warp == 2 + kFxSplitWarps && lane == 0) fx_store_wait_all();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================================
// weight gradient: dW[j][k] = sum_b G[b][j] X[b][k], k tiles of 128
// smem: [0,256) barriers | raw stage s (4): X [BP][64] fp32 (half a tile) | X piece set s (2): 3 x 16 x xs | G pieces 3 x 16 x gs
// =====================================================================================================================
__global__ void __launch_bounds__(kFxThreads, 1) fc1x3_wgrad_kernel(const FxArgs a, const __grid_constant__ CUtensorMap tm_x,
                                                                    const __grid_constant__ CUtensorMap tm_out) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* rfull = reinterpret_cast<uint64_t*>(smem);  // [4]
  uint64_t* rempty = rfull + 4;                          // [4]
  uint64_t* pfull = rempty + 4;                          // [2]
  uint64_t* pempty = pfull + 2;                          // [2]
  uint64_t* dfull = pempty + 2;                          // [2]
  uint64_t* dempty = dfull + 2;                          // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(dempty + 2);
  const int BP = a.BP;
  const uint32_t xs = static_cast<uint32_t>(BP) * 16u + 16u, gs = xs;
  const uint32_t raw_x = static_cast<uint32_t>(BP) * 256u;
  const uint32_t xp = (16u * xs + 127u) & ~127u, gp = xp;
  uint8_t* raw_s = smem + 256;
  uint8_t* x_s = raw_s + 4u * raw_x;   // [2 sets][3 pieces]
  uint8_t* g_s = x_s + 2u * 3u * xp;
  uint8_t* o_s = g_s + 3u * gp;  // [2 warp pairs][128 j][64 k] fp32
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { tc::mbar_init(rfull + i, 1); tc::mbar_init(rempty + i, kFxSplitWarps); }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(pfull + i, kFxSplitWarps); tc::mbar_init(pempty + i, 1);
      tc::mbar_init(dfull + i, 1); tc::mbar_init(dempty + i, 4);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, 512);
  if (warp >= 2 && warp < 2 + kFxSplitWarps) fx_split_g(a.g, a.B, BP, a.F1, g_s, gp, gs, threadIdx.x - 64, kFxSplitWarps * 32);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // tiles are dealt round-robin: at any moment the CTAs stream ONE contiguous span of every matrix row (DRAM page locality)
  const long long t_begin = blockIdx.x, t_end = a.tiles, t_step = gridDim.x;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t seq = 0;
      for (long long t = t_begin; t < t_end; t += t_step)
        for (int hf = 0; hf < 2; ++hf, ++seq) {
          const uint32_t st = seq & 3u;
          tc::mbar_wait(rempty + st, ((seq >> 2) & 1u) ^ 1u);
          tc::mbar_arrive_expect_tx(rfull + st, raw_x);
          fx_tma_2d(raw_s + st * raw_x, &tm_x, static_cast<int>(t * 128 + hf * kFxKT), 0, rfull + st);
        }
    }
  } else if (warp == 1) {
    const bool leader = tc::elect_one();
    // A = X pieces MN-major (M = k, K = b): 8-k groups xs apart (SBO), 8-b groups 128 B apart (LBO)
    // B = G pieces MN-major (N = j, K = b): 8-j groups gs apart (SBO), 8-b groups 128 B apart (LBO)
    const uint32_t a_hi = (xs >> 4) | (1u << 14), b_hi = (gs >> 4) | (1u << 14), lbo = ((128u >> 4) << 16);
    const uint32_t x16 = tc::smem_u32(x_s) >> 4, g16 = tc::smem_u32(g_s) >> 4, xp16 = xp >> 4, gp16 = gp >> 4;
    const uint32_t idesc = tc::umma_idesc(128, 128, /*BF16*/ 1, 1, 1);
    const int k16n = BP >> 4;
    uint32_t seq = 0;
    for (long long t = t_begin; t < t_end; t += t_step, ++seq) {
      const uint32_t st = seq & 1u;
      tc::mbar_wait(dempty + st, ((seq >> 1) & 1u) ^ 1u);
      tc::mbar_wait(pfull + st, (seq >> 1) & 1u);
      tc::tc_fence_after();
      const uint32_t d = tmem_base + st * 128u;
      const uint32_t a0 = lbo | ((x16 + st * 3u * xp16) & 0x3fffu), b0 = lbo | (g16 & 0x3fffu);
      if (leader) {
#pragma unroll
        for (int k16 = 0; k16 < 4; ++k16) {
          if (k16 >= k16n) break;
          fx_mma6(d, a0 + static_cast<uint32_t>(k16) * 16u, xp16, a_hi, b0 + static_cast<uint32_t>(k16) * 16u, gp16, b_hi, idesc, k16 == 0);
        }
        tc::umma_commit(pempty + st);
        tc::umma_commit(dfull + st);
      }
      __syncwarp();
    }
  } else if (warp < 2 + kFxSplitWarps) {
    const int tid = threadIdx.x - 64;
    uint32_t seq = 0, tile = 0;
    for (long long t = t_begin; t < t_end; t += t_step, ++tile) {
      const uint32_t ps = tile & 1u;
      tc::mbar_wait(pempty + ps, ((tile >> 1) & 1u) ^ 1u);
      for (int hf = 0; hf < 2; ++hf, ++seq) {
        const uint32_t st = seq & 3u;
        tc::mbar_wait(rfull + st, (seq >> 2) & 1u);
        fx_split_tile<8>(raw_s + st * raw_x, BP, x_s + ps * 3u * xp, xp, xs, hf * 8, tid, kFxSplitWarps * 32);
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(rempty + st);
      }
      tc::fence_proxy_async();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(pfull + ps);
    }
  } else {
    const int qd = warp & 3, row = qd * 32 + lane;  // row = k within the tile, columns = j
    const int pair = qd >> 1, r64 = row & 63;
    float* out_s = reinterpret_cast<float*>(o_s + static_cast<uint32_t>(pair) * 32768u);
    uint32_t seq = 0;
    for (long long t = t_begin; t < t_end; t += t_step, ++seq) {
      const uint32_t st = seq & 1u;
      const long long k = t * 128 + row;
      tc::mbar_wait(dfull + st, (seq >> 1) & 1u);
      tc::tc_fence_after();
      const uint32_t addr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + st * 128u;
      // the accumulator goes through a shared-memory tile [j][64 k] per warp pair and leaves as ONE tensor-map store
      if (lane == 0 && (qd & 1) == 0) fx_store_wait_read();  // the pair's previous store has read the staging tile
      fx_pair_sync(pair);
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32(addr + c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) out_s[(c0 + c) * 64 + r64] = __uint_as_float(v[c]);
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(dempty + st);
      tc::fence_proxy_async();
      fx_pair_sync(pair);
      if (lane == 0 && (qd & 1) == 0 && !(a.flags & 1)) {
        if (a.accumulate) fx_tma_add_2d(&tm_out, out_s, static_cast<int>(t * 128 + pair * 64), 0);
        else fx_tma_store_2d(&tm_out, out_s, static_cast<int>(t * 128 + pair * 64), 0);
      }
    }
    if (lane == 0 && (qd & 1) == 0) fx_store_wait_all();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

// ---- host ---------------------------------------------------------------------------------------------------------
int g_fc1x3 = 1;
int g_fc1x3_flags = 0;  // tools may switch the tensor-core fc1 off through pvb200_debug_set_fc1x3

// 2-D fp32 tensor map over a row-major matrix [rows][K1]; box = (64 k, box_rows); rows beyond `rows` arrive as zeros
static int fx_make_map(CUtensorMap* tm, const float* base, long long K1, int rows, int box_rows, int box_k = kFxKT) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return -1;
    encode = reinterpret_cast<EncodeFn>(fn);
  }
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(K1), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstr[1] = {static_cast<cuuint64_t>(K1) * 4};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_k), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t es[2] = {1, 1};
  return static_cast<int>(encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
}

bool fc1x3_ok(int B, int max_b, int F1, long long K1, const void* x, const void* w) {
  return g_fc1x3 && B >= 1 && B <= max_b && F1 >= 1 && F1 <= 128 && K1 >= 64 && K1 % 4 == 0 && K1 < (1LL << 31) - 256 &&
         reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(w) % 16 == 0;
}

// the K split of the forward is one slice per SM of the WHOLE device (not of the SMs left by pvb200_reserve_sms): a
// forecast does not depend on what else is running
int fc1x3_fwd_ctas(long long K1) {
  static int sms = 0;
  if (sms <= 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      return -1;
    sms = n;
  }
  const long long tiles = ceil_div(K1, static_cast<long long>(kFxKT));
  return static_cast<int>(tiles < sms ? tiles : sms);
}

// partial [S = fc1x3_fwd_ctas(K1)][B][F1].  Batches larger than 48 go through in chunks of 48 samples (each streams W once):
// a sample's row of the result does not depend on the batch it rides in, whatever its size.
int fc1x3_fwd(const float* x, const float* w, float* partial, int S, int B, int F1, long long K1, cudaStream_t st) {
  constexpr int kChunk = 48;
  CUtensorMap tw;
  PVB_REQUIRE(fx_make_map(&tw, w, K1, F1, 128) == 0, "fc1x3_fwd: cuTensorMapEncodeTiled failed");
  for (int b0 = 0; b0 < B; b0 += kChunk) {
    FxArgs a;
    a.flags = g_fc1x3_flags; a.x = x + b0 * K1; a.g = nullptr; a.out = partial + static_cast<long long>(b0) * F1;
    a.B = B - b0 < kChunk ? B - b0 : kChunk; a.BP = round_up(a.B, 16); a.F1 = F1; a.Btot = B; a.accumulate = 0; a.K1 = K1;
    a.tiles = ceil_div(K1, static_cast<long long>(kFxKT));
    CUtensorMap tx;
    PVB_REQUIRE(fx_make_map(&tx, a.x, K1, a.B, a.BP) == 0, "fc1x3_fwd: cuTensorMapEncodeTiled failed");
    const uint32_t xs = a.BP * 16 + 16;
    const size_t smem = 128 + 2 * (128 * 256 + a.BP * 256) + 2 * (3 * 8 * kFxWS + 3 * round_up(8 * xs, 128u));
    PVB_CUDA(cudaFuncSetAttribute(fc1x3_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fc1x3_fwd_kernel<<<S, kFxThreads, smem, st>>>(a, tw, tx);
    PVB_LAUNCHED("fc1x3_fwd");
  }
  return PVB200_OK;
}

// Batches larger than 48 go through in chunks of 48 samples, each streaming W (dgrad) / writing dW (wgrad) once: the data
// gradient's samples are independent; the weight gradient's later chunks ADD their sums to dW through the TMA engine
// (cp.reduce.async.bulk.tensor ... .add), launch after launch in stream order -- deterministic.
constexpr int kFxBwdChunk = 48;

int fc1x3_dgrad(const float* g, const float* w, const float* x, float* gx, int B, int F1, long long K1, cudaStream_t st) {
  const int sms = sm_count();
  PVB_REQUIRE(sms > 0, "fc1x3_dgrad: no CUDA device");
  CUtensorMap tw;
  PVB_REQUIRE(fx_make_map(&tw, w, K1, F1, 32, 128) == 0, "fc1x3_dgrad: cuTensorMapEncodeTiled failed");
  for (int b0 = 0; b0 < B; b0 += kFxBwdChunk) {
    FxArgs a;
    a.flags = g_fc1x3_flags; a.x = x + b0 * K1; a.g = g + static_cast<long long>(b0) * F1; a.out = gx + b0 * K1;
    a.B = B - b0 < kFxBwdChunk ? B - b0 : kFxBwdChunk; a.BP = round_up(a.B, 16); a.F1 = F1; a.Btot = B; a.accumulate = 0; a.K1 = K1;
    a.tiles = ceil_div(K1, 128LL);
    CUtensorMap tx, to;
    PVB_REQUIRE(fx_make_map(&tx, a.x, K1, a.B, a.BP, 128) == 0 && fx_make_map(&to, a.out, K1, a.B, a.BP, 128) == 0,
                "fc1x3_dgrad: cuTensorMapEncodeTiled failed");
    const uint32_t gs = a.BP * 16 + 16;
    const size_t smem = 256 + 4 * 32 * 512 + kFxDSets * 3 * 16 * kFxDS + 3 * round_up(16 * gs, 128u) + 2 * a.BP * 512;
    const long long grid = a.tiles < sms ? a.tiles : sms;
    PVB_CUDA(cudaFuncSetAttribute(fc1x3_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fc1x3_dgrad_kernel<<<static_cast<unsigned>(grid), kFxThreads, smem, st>>>(a, tw, tx, to);
    PVB_LAUNCHED("fc1x3_dgrad");
  }
  return PVB200_OK;
}

int fc1x3_wgrad(const float* g, const float* x, float* dw, int B, int F1, long long K1, cudaStream_t st) {
  const int sms = sm_count();
  PVB_REQUIRE(sms > 0, "fc1x3_wgrad: no CUDA device");
  CUtensorMap to;
  PVB_REQUIRE(fx_make_map(&to, dw, K1, F1, 128) == 0, "fc1x3_wgrad: cuTensorMapEncodeTiled failed");
  for (int b0 = 0; b0 < B; b0 += kFxBwdChunk) {
    FxArgs a;
    a.flags = g_fc1x3_flags; a.x = x + b0 * K1; a.g = g + static_cast<long long>(b0) * F1; a.out = dw;
    a.B = B - b0 < kFxBwdChunk ? B - b0 : kFxBwdChunk; a.BP = round_up(a.B, 16); a.F1 = F1; a.Btot = B; a.accumulate = b0 > 0; a.K1 = K1;
    a.tiles = ceil_div(K1, 128LL);
    CUtensorMap tx;
    PVB_REQUIRE(fx_make_map(&tx, a.x, K1, a.B, a.BP) == 0, "fc1x3_wgrad: cuTensorMapEncodeTiled failed");
    const uint32_t xs = a.BP * 16 + 16;
    const size_t smem = 256 + 4 * a.BP * 256 + 3 * 3 * round_up(16 * xs, 128u) + 2 * 32768;
    const long long grid = a.tiles < sms ? a.tiles : sms;
    PVB_CUDA(cudaFuncSetAttribute(fc1x3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fc1x3_wgrad_kernel<<<static_cast<unsigned>(grid), kFxThreads, smem, st>>>(a, tx, to);
    PVB_LAUNCHED("fc1x3_wgrad");
  }
  return PVB200_OK;
}

}  // namespace pvb

extern "C" {
/* tools only (not declared in pvb200.h): 0 = fp32 head on the FMA-pipe kernels, 1 = tensor-core fc1 where it applies */
void pvb200_debug_set_fc1x3(int on) { pvb::g_fc1x3 = on & 1; pvb::g_fc1x3_flags = on >> 1; }
}
