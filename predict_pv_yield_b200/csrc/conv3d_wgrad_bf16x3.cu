// conv3d_wgrad_bf16x3.cu -- a11 in fp32 MODE on the tensor cores: the Conv3d 3x3x3 weight (and bias) gradient as
// tcgen05.mma.kind::f16 GEMMs over a split of the fp32 operands into 16-bit pieces (fp32-class accuracy): the TWO-WAY
// fp16 split of scaled operands (three products per fp32 product; pvb200_conv3d_wgrad_f16x2, what the model runs) or the
// THREE-WAY bf16 split (six products; pvb200_conv3d_wgrad_bf16x3, needs no scaling) -- one kernel template.
//
// Reference: autograd of nn.Conv3d, predict_pv_yield/models/conv3d/model.py:80-90,117-120:
//   dW[co][ci][kt][kh][kw] = sum_{b,t,h,w} gz[b][co][t][h][w] * x[b][ci][t+kt][h+kh][w+kw],   db[co] = sum gz.
//
// Why not 3xTF32 like the first forward / data gradient (conv3d_igemm_tf32x3.cu): the reduction runs over output positions,
// so positions are the K dimension and both operands must be read "MN-major" (channels contiguous, positions strided) from
// the blocked layout.  kind::tf32 returns ZEROS for MN-major SWIZZLE_NONE operands (tools/probe/tf32_mn_probe.cu; CUTLASS:
// "for mn-major tf32 operands, SW128_32B is the only available smem layout", a 32-bit-granular swizzle bulk copies cannot
// produce).  kind::f16 has no such restriction.  Three-way bf16 split: an fp32 value is EXACTLY the sum of three bf16 values
// (v = b0 + b1 + b2, 8 + 8 + 8 significand bits, round-to-nearest residuals), so
//   x . g = x0 g0 + x0 g1 + x1 g0 + x1 g1 + x0 g2 + x2 g0      (the dropped terms are <= 2^-24 relative)
// costs six K = 16 MMAs per 16 positions.  Two-way fp16 split: s v = h0 + h1 with 11 + 11 significand bits after scaling
// by the power of two that brings the tensor's largest magnitude into fp16's range (w3_scale_exp), x . g = x1 g0 + x0 g1 +
// x0 g0 (dropped: 2^-22): three MMAs, half the operand bytes -- see w3_split2h.
//
// GEMM shape.  One step of the kernel = one input plane p of one sample and one output row h:
//   * two tiled TMA loads through tensor maps stage the raw fp32 rows (the three input rows h..h+2 of every channel group --
//     contiguous in memory; row h of the gradient planes p, p-1, p-2); eight "split" warps turn them into the 16-bit
//     pieces in the operand layout [group of 8 channels][row][position][8] (two fp32 channel groups merge into one group).
//   * A (M = 128): pieces of the input rows as [g][kh][Wi]: M-group m = 3 g + kh sits at the uniform stride Wi * 16 B, so
//     the kh taps cost no copies; M-group 3 G8 is a constant row of ones (bias gradient for free); the remaining M rows
//     read whatever follows inside the buffer and are never looked at.
//   * the kw taps are the descriptor START address (kw * 16 B): three accumulators, no copies.
//   * B (N = 96): pieces of the gradient rows as [kt][g][WP] (WP = Wo rounded up to 16, the padding stays zero): plane p
//     contributes to the three time taps at once; taps that fall outside the output are cut off by narrowing N.
//   => per step 3 (kw) x WP/16 x 3 (or 6) MMAs of 128 x 96 x 16 against 48 KB of fp32 operands from L2.
// Step order (w3_step): blocks of 6 output rows x all input planes, so that the re-reads of a gradient plane hit L2; the
// ranges of steps a CTA works on come through a shared-memory ring -- static split or chunks from an atomic counter.
// Accuracy.  The tensor core's fp32 accumulator rounds toward zero (tools/probe/tf32_acc_probe.cu): every kFlush steps the
// three accumulators are drained and added -- in fp32 round-to-nearest, red.global.add.f32 -- to the CTA's private partial
// in global memory (L2 resident, 147 KB per CTA).  A second kernel reduces the per-CTA partials in fixed order into
// dW [Co][Ci][3][3][3] and db, and takes the operand scales out.
// Warp roles (448 threads): warp 0 producer (TMA loads, step ranges), warp 1 MMA issuer + TMEM owner, warps 2-9 split,
// warps 10-13 accumulator drain.  Pipeline: raw staging (two buffers; one for the three-way split) -> pieces (two buffers) -> MMA.
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, libcuda is not linked)
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pvb {

constexpr int kW3Threads = 448;   // 2 + 8 split + 4 drain warps: 146 registers per thread (704 threads left 80 and the split spilled)
constexpr int kW3SplitWarps = 8;
constexpr int kW3Stages = 2;    // piece buffers
constexpr int kW3Flush = 16;    // steps between two drains of the accumulators
constexpr int kW3AccCols = 96;  // TMEM columns per kw accumulator
constexpr int kW3Pairs = 6;     // products of the three-way split that are kept
constexpr int kW3RowBlock = 6;  // output rows per block of the step order (see w3_step)
constexpr int kW3Items = 3;     // items (8 channels of one position) a split thread converts per operand and step, at most

struct W3Args {
  int gz_pad;       // zero padding of the gradient tensor on T, H, W (coordinates of the gradient's tensor map are shifted by it)
  float* partial;   // [grid][3 kw][96 columns][128 rows]
  int B, G, Ti, Hi, Wi;  // G: fp32 groups of 4 channels (even)
  int GOr;          // gradient fp32 channel groups present in memory (even)
  int CoP;          // 16 or 32: columns per time tap (N = 3 * CoP)
  int To, Ho, Wo, WP;
  int plane_off;    // input plane of output t, tap kt: t + kt + plane_off
  long long steps;  // B * Ti * Ho
  int dbg_flags;    // tools only: 1 = no bulk copies, 2 = no split arithmetic, 4 = no MMAs, 8 = no drain traffic
  int* sched;       // dynamic step ranges: global counter of claimed chunks (zeroed before the launch), or null = static split
  const float* amax_x;  // two-way fp16 split only: max |x| and max |gz| of the two tensors (device scalars)
  const float* amax_g;
};

__device__ __forceinline__ void w3_mma(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
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

// one tiled TMA load of a 5-D box (SASS UTMALDG): coordinates innermost first; out-of-bounds elements arrive as zeros
__device__ __forceinline__ void w3_tma_5d(void* dst_smem, const CUtensorMap* tm, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          tc::smem_u32(dst_smem)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// v = b0 + b1 + b2 exactly (round-to-nearest pieces: every residual is exact in fp32 and fits the next piece), two values
// at a time: cvt.rn.bf16x2.f32 is one full-rate instruction (SASS F2FP.BF16.F32.PACK_AB; the scalar F2F.BF16.F32 runs
// on the quarter-rate conversion pipe and made the split warps the bottleneck of the kernel)
__device__ __forceinline__ void w3_split2(float v0, float v1, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(v1), "f"(v0));  // high half <- first source
  const float r0 = v0 - __uint_as_float(p0 << 16), r1 = v1 - __uint_as_float(p0 & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(r1), "f"(r0));
  const float s0 = r0 - __uint_as_float(p1 << 16), s1 = r1 - __uint_as_float(p1 & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p2) : "f"(s1), "f"(s0));
}
// two fp32 channel groups (8 channels of one position) -> the three bf16 pieces (16 bytes each)
__device__ __forceinline__ void w3_split8(const float4 lo4, const float4 hi4, uint4& q0, uint4& q1, uint4& q2) {
  w3_split2(lo4.x, lo4.y, q0.x, q1.x, q2.x);
  w3_split2(lo4.z, lo4.w, q0.y, q1.y, q2.y);
  w3_split2(hi4.x, hi4.y, q0.z, q1.z, q2.z);
  w3_split2(hi4.z, hi4.w, q0.w, q1.w, q2.w);
}

// ---- two-way fp16 split: s v = h0 + h1 + O(2^-22 |s v|), s a power of two that brings the tensor's largest magnitude to
// [2^14, 2^15) (w3_scale_exp): 11 + 11 significand bits, three of the four piece products -- the accuracy class of the 3xTF32
// forward at HALF the MMAs of the three-way bf16 split.  Values below 2^-18 of the tensor's maximum lose relative (not
// absolute) precision: their second piece is an fp16 subnormal with an absolute step of 2^-39 of the maximum.
__device__ __forceinline__ int w3_scale_exp(float amax) {
  if (!(amax > 0.f) || !(amax <= 3.0e38f)) return 0;
  const int e = 14 - (static_cast<int>((__float_as_uint(amax) >> 23) & 0xffu) - 127);
  return e < -100 ? -100 : (e > 100 ? 100 : e);
}
__device__ __forceinline__ float w3_exp2i(int e) { return __uint_as_float(static_cast<uint32_t>(e + 127) << 23); }
__device__ __forceinline__ void w3_split2h(float v0, float v1, uint32_t& p0, uint32_t& p1) {
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(v1), "f"(v0));  // high half <- first source
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&p0));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(v1 - f.y), "f"(v0 - f.x));
}
__device__ __forceinline__ void w3_split8h(const float4 lo4, const float4 hi4, float s, uint4& q0, uint4& q1) {
  w3_split2h(lo4.x * s, lo4.y * s, q0.x, q1.x);
  w3_split2h(lo4.z * s, lo4.w * s, q0.y, q1.y);
  w3_split2h(hi4.x * s, hi4.y * s, q0.z, q1.z);
  w3_split2h(hi4.z * s, hi4.w * s, q0.w, q1.w);
}

__device__ __forceinline__ float w3_ld_keep(const float* p, uint64_t policy) {
  float v;
  asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(policy) : "memory");
  return v;
}
__device__ __forceinline__ void w3_st_keep(float* p, float v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(policy) : "memory");
}

__device__ __forceinline__ void w3_red_keep(float* p, float v, uint64_t policy) {
  asm volatile("red.global.add.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(policy) : "memory");
}

struct W3Step {
  int b, p, h, kt_lo, kt_hi;  // time taps kt_lo..kt_hi of plane p fall on existing outputs (kt_lo > kt_hi: none)
  int h0, hb;                 // the row block [h0, h0 + hb) the step belongs to
};
__device__ __forceinline__ void w3_taps(W3Step& r, const W3Args& a) {
  // t = p - kt - plane_off in [0, To)
  const int tmax = r.p - a.plane_off;  // kt = 0
  const int lo = tmax - (a.To - 1);
  r.kt_lo = lo > 0 ? lo : 0;
  r.kt_hi = tmax < 2 ? tmax : 2;
}
// Step order inside a sample: blocks of kW3RowBlock output rows, all input planes of a block, the rows of the block.  A
// gradient plane is read by the steps of three consecutive input planes: with the rows of a whole plane in between (the
// first order: 60 steps x 148 CTAs x 48 KB = 4x the L2) every one of those reads came from DRAM -- 1.25 GB per launch
// against 0.59 GB of tensors; now they are 2 x kW3RowBlock steps apart and hit L2, at the price of re-reading the two halo
// rows of x per block.  Decoded once per role (64-bit divisions), then advanced incrementally -- the per-step divisions sat
// on the MMA warp's issue path and left the tensor pipe idle at every step boundary.
__device__ __forceinline__ W3Step w3_step(long long s, const W3Args& a) {
  W3Step r;
  const long long per_b = static_cast<long long>(a.Ti) * a.Ho;
  r.b = static_cast<int>(s / per_b);
  const int rs = static_cast<int>(s - r.b * per_b);
  const int k = rs / (kW3RowBlock * a.Ti);  // every block before the last one is full
  r.h0 = k * kW3RowBlock;
  r.hb = a.Ho - r.h0 < kW3RowBlock ? a.Ho - r.h0 : kW3RowBlock;
  const int q = rs - k * kW3RowBlock * a.Ti;
  r.p = q / r.hb;
  r.h = r.h0 + q % r.hb;
  w3_taps(r, a);
  return r;
}
__device__ __forceinline__ void w3_next(W3Step& r, const W3Args& a) {
  if (++r.h == r.h0 + r.hb) {
    r.h = r.h0;
    if (++r.p == a.Ti) {
      r.p = 0;
      r.h0 += r.hb;
      if (r.h0 >= a.Ho) { r.h0 = 0; ++r.b; }
      r.hb = a.Ho - r.h0 < kW3RowBlock ? a.Ho - r.h0 : kW3RowBlock;
      r.h = r.h0;
    }
    w3_taps(r, a);
  }
}

// smem layout (bytes): [0,256) barriers | [256,384) the ring of step ranges | [384,512) a zero core matrix | piece buffers: stage s = A pieces 0..NP-1 (a_piece
// bytes each), for s = 0,1, then stage s = B pieces (b_piece bytes each) | raw fp32 staging: RB x A rows, RB x B rows (the
// two-way split leaves room for a second raw buffer: the loads of step s+1 fly while step s is converted)
// ---- step ranges of a CTA -------------------------------------------------------------------------------------------------
// Static: one contiguous range per CTA (steps * cta / grid).  Dynamic (a.sched != null; switched on under data parallelism,
// pvb200_set_dynamic_tiles): chunks of kW3Chunk steps claimed with an atomic counter -- when NCCL's kernels occupy some SMs
// while this persistent grid is launched, the displaced CTAs start late (or never): with the static split the whole kernel
// then waits for their ranges (a grid tail, +26 % measured on two GPUs), with chunks the resident CTAs simply take more.
// The producer warp claims a range and publishes it through a small shared-memory ring; the other roles read the same
// sequence.  A chunk is a whole number of flush windows, so windows never straddle ranges.
constexpr int kW3Ring = 4;
constexpr int kW3Chunk = kW3Flush;  // steps per claimed chunk
struct W3Ranges {
  long long* lo;     // [kW3Ring]
  long long* hi;     // [kW3Ring]   lo >= hi: no more work
  uint64_t* full;    // [kW3Ring]   range published (1 arrival)
  uint64_t* empty;   // [kW3Ring]   range read by every consumer warp
};
// producer warp (all lanes): claim and publish range number ci
__device__ __forceinline__ void w3_publish_range(const W3Ranges& r, uint32_t ci, const W3Args& a, long long& lo, long long& hi) {
  const uint32_t slot = ci % kW3Ring;
  if ((threadIdx.x & 31) == 0) {
    tc::mbar_wait(r.empty + slot, ((ci / kW3Ring) & 1u) ^ 1u);
    if (a.sched) {
      const long long id = atomicAdd(a.sched, 1);
      lo = id * kW3Chunk;
      hi = lo + kW3Chunk < a.steps ? lo + kW3Chunk : a.steps;
    } else if (ci == 0) {
      lo = a.steps * blockIdx.x / gridDim.x;
      hi = a.steps * (blockIdx.x + 1) / gridDim.x;
    } else {
      lo = hi = a.steps;
    }
    r.lo[slot] = lo;
    r.hi[slot] = hi;
    tc::mbar_arrive(r.full + slot);  // release: the range is visible to the warps that acquire the barrier
  }
  lo = __shfl_sync(0xffffffffu, lo, 0);
  hi = __shfl_sync(0xffffffffu, hi, 0);
}
// consumer warps (all lanes)
__device__ __forceinline__ void w3_next_range(const W3Ranges& r, uint32_t ci, long long& lo, long long& hi) {
  const uint32_t slot = ci % kW3Ring;
  tc::mbar_wait(r.full + slot, (ci / kW3Ring) & 1u);
  lo = r.lo[slot];
  hi = r.hi[slot];
  __syncwarp();
  if ((threadIdx.x & 31) == 0) tc::mbar_arrive(r.empty + slot);
}

// NP = 3: three-way bf16 split (six products); NP = 2: two-way fp16 split of the scaled operands (three products)
template <int NP>
__global__ void __launch_bounds__(kW3Threads, 1) conv3d_wgrad_bf16x3_kernel(const W3Args a, const __grid_constant__ CUtensorMap tm_x,
                                                                            const __grid_constant__ CUtensorMap tm_gz) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr uint32_t RB = NP == 2 ? 2u : 1u;                // raw staging buffers
  uint64_t* raw_full = reinterpret_cast<uint64_t*>(smem);  // [2][RB] raw rows landed: [0] input rows, [1] gradient rows
  uint64_t* raw_empty = raw_full + 2 * RB;                  // [2][RB] raw rows converted (each half is refilled as soon as it is read)
  uint64_t* ready = raw_empty + 2 * RB;                     // [2] pieces written
  uint64_t* empty = ready + kW3Stages;                      // [2] pieces consumed
  uint64_t* afull = empty + kW3Stages;                      // [3] accumulator kw complete (flush window closed)
  uint64_t* aempty = afull + 3;                             // [3] accumulator kw drained
  uint64_t* rfull = aempty + 3;                             // [4] step range published
  uint64_t* rempty = rfull + kW3Ring;                       // [4] step range read by the consumer warps
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(rempty + kW3Ring);
  const W3Ranges ranges = {reinterpret_cast<long long*>(smem + 256), reinterpret_cast<long long*>(smem + 256) + kW3Ring, rfull, rempty};
  const int G = a.G, G8 = a.G / 2, Wi = a.Wi, WP = a.WP, CoP = a.CoP;
  const int GP8 = CoP / 8;                                                   // gradient groups of 8 per time tap in the pieces
  const uint32_t a_piece = ((static_cast<uint32_t>(3 * G8 + 4) * Wi * 16u) + 127u) & ~127u;  // staged rows + ones rows + slack
  const uint32_t b_piece = static_cast<uint32_t>(3 * GP8 * WP) * 16u;
  const uint32_t a_raw_bytes = static_cast<uint32_t>(G * 3 * Wi) * 16u;        // box [G][3 rows][Wi] of x
  const uint32_t a_raw_span = (a_raw_bytes + 127u) & ~127u;
  const uint32_t b_raw_bytes = static_cast<uint32_t>(a.GOr * 3 * a.Wo) * 16u;  // box [GOr][3 planes][Wo] of gz
  const uint32_t b_raw_span = (b_raw_bytes + 127u) & ~127u;
  uint8_t* a_s = smem + 512;                          // [stage][piece]
  uint8_t* b_s = a_s + kW3Stages * NP * a_piece;      // [stage][piece]
  uint8_t* a_raw = b_s + kW3Stages * NP * b_piece;
  uint8_t* b_raw = a_raw + RB * a_raw_span;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // zero everything once: padding positions / groups, the slack rows and the zero core matrix stay zero for the whole kernel
  {
    const uint32_t total16 = (128u + kW3Stages * NP * (a_piece + b_piece) + RB * (a_raw_span + b_raw_span)) >> 4;
    uint4* z = reinterpret_cast<uint4*>(smem + 384);
    for (uint32_t i = threadIdx.x; i < total16; i += kW3Threads) z[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  // the ones rows of piece 0 (M-group 3 G8 = "row" 3 G8 of the operand layout); pieces 1, 2 keep zeros there
  for (int s = 0; s < kW3Stages; ++s) {
    uint4* ones = reinterpret_cast<uint4*>(a_s + (s * NP) * a_piece) + static_cast<uint32_t>(3 * G8) * Wi;
    const uint32_t one2 = NP == 3 ? 0x3f803f80u : 0x3c003c00u;  // 1.0 in bf16 / fp16, twice
    for (int i = threadIdx.x; i < 3 * Wi; i += kW3Threads) ones[i] = make_uint4(one2, one2, one2, one2);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * static_cast<int>(RB); ++i) { tc::mbar_init(raw_full + i, 1); tc::mbar_init(raw_empty + i, kW3SplitWarps); }
    for (int i = 0; i < kW3Stages; ++i) { tc::mbar_init(ready + i, kW3SplitWarps); tc::mbar_init(empty + i, 1); }
    for (int i = 0; i < 3; ++i) { tc::mbar_init(afull + i, 1); tc::mbar_init(aempty + i, 4); }
    for (int i = 0; i < kW3Ring; ++i) { tc::mbar_init(rfull + i, 1); tc::mbar_init(rempty + i, kW3Threads / 32 - 1); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, 512);
  tc::fence_proxy_async();  // the zero / ones fill is read by the tensor core
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  long long s_begin = 0, s_end = 0;
  if (warp == 0) {
    // =============================== producer ===============================
    uint32_t seq = 0;
    for (uint32_t ci = 0;; ++ci) {
    w3_publish_range(ranges, ci, a, s_begin, s_end);
    if (s_begin >= s_end) break;
    W3Step st = w3_step(s_begin, a);
    for (long long s = s_begin; s < s_end; ++s, ++seq, w3_next(st, a)) {
      // two tiled TMA loads per step (tensor maps, SASS UTMALDG) instead of 32 small bulk copies: the three input rows
      // of every channel group, and row h of the three gradient planes t = p - off - 2 .. p - off of every group (planes
      // outside the tensor arrive as zeros: out-of-bounds fill / the tensor's own zero padding)
      // the two halves of the raw staging have their own barriers: the input rows of step s+1 are fetched while the split
      // warps still convert the gradient rows of step s
      if (lane == 0) {
        const uint32_t rb = seq % RB, ph = ((seq / RB) & 1u) ^ 1u;
        tc::mbar_wait(raw_empty + rb, ph);
        if (a.dbg_flags & 1) {
          tc::mbar_arrive(raw_full + rb);
        } else {
          tc::mbar_arrive_expect_tx(raw_full + rb, a_raw_bytes);
          w3_tma_5d(a_raw + rb * a_raw_span, &tm_x, 0, st.h, st.p, 0, st.b, raw_full + rb);
        }
        tc::mbar_wait(raw_empty + RB + rb, ph);
        if (a.dbg_flags & 1) {
          tc::mbar_arrive(raw_full + RB + rb);
        } else {
          tc::mbar_arrive_expect_tx(raw_full + RB + rb, b_raw_bytes);
          w3_tma_5d(b_raw + rb * b_raw_span, &tm_gz, a.gz_pad * 4, st.h + a.gz_pad, st.p - a.plane_off - 2 + a.gz_pad, 0, st.b,
                    raw_full + RB + rb);
        }
      }
      __syncwarp();
    }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const bool leader = tc::elect_one();
    // MN-major SWIZZLE_NONE descriptors: SBO = stride between groups of 8 channels, LBO = stride between the two
    // 8-position K groups of a K = 16 step (contiguous positions: 128 B)
    const uint32_t a_hi_word = ((static_cast<uint32_t>(Wi) * 16u) >> 4) | (1u << 14);
    const uint32_t b_hi_word = ((static_cast<uint32_t>(WP) * 16u) >> 4) | (1u << 14);
    const uint32_t lbo_word = (128u >> 4) << 16;
    const uint32_t a_addr16 = tc::smem_u32(a_s) >> 4, b_addr16 = tc::smem_u32(b_s) >> 4;
    const uint32_t a_piece16 = a_piece >> 4, b_piece16 = b_piece >> 4;
    const uint32_t idesc0 = tc::umma_idesc(128, 0, /*BF16 : F16*/ NP == 3 ? 1 : 0, /*A MN-major*/ 1, /*B MN-major*/ 1);
    const uint32_t idesc_blk = static_cast<uint32_t>(CoP >> 3) << 17;
    // B operand of the MMA that opens a flush window: every N group reads the same all-zero core matrices (SBO = 0)
    const uint32_t zero_lo = ((tc::smem_u32(smem + 384) >> 4) & 0x3fffu);  // LBO = 0 as well: both K groups read it
    const uint32_t zero_hi = (1u << 14);
    const int k16n = WP >> 4;
    uint32_t seq = 0;
    uint32_t nwin = 0;  // flush windows closed so far (phase of the accumulator barriers)
    for (uint32_t ci = 0;; ++ci) {
    w3_next_range(ranges, ci, s_begin, s_end);
    if (s_begin >= s_end) break;
    W3Step st = w3_step(s_begin, a);
    for (long long s = s_begin; s < s_end; ++s, ++seq, w3_next(st, a)) {
      const uint32_t stage = seq % kW3Stages;
      const bool win_first = (seq % kW3Flush) == 0;
      const bool win_last = ((seq + 1) % kW3Flush) == 0 || (s + 1 == s_end);
      tc::mbar_wait(ready + stage, (seq / kW3Stages) & 1u);
      tc::tc_fence_after();
      const int nb = st.kt_hi - st.kt_lo + 1;
      const uint32_t a0 = lbo_word | ((a_addr16 + (stage * NP) * a_piece16) & 0x3fffu);
      const uint32_t b0 = lbo_word | ((b_addr16 + (stage * NP) * b_piece16 + static_cast<uint32_t>(st.kt_lo * GP8 * WP)) & 0x3fffu);
      const uint32_t idesc = idesc0 + static_cast<uint32_t>(nb > 0 ? nb : 1) * idesc_blk;
#pragma unroll 1
      for (int kw = 0; kw < 3; ++kw) {
        if (win_first) {
          tc::mbar_wait(aempty + kw, (nwin & 1u) ^ 1u);  // the previous window of this accumulator has been drained
          tc::tc_fence_after();
        }
        const uint32_t d = tmem_base + static_cast<uint32_t>(kw * kW3AccCols + st.kt_lo * CoP);
        // a window opens with an MMA against the zero matrix that overwrites ALL three time-tap blocks of the accumulator
        // (a narrowed step only touches its own columns: the others must not keep the drained window's values)
        if (leader && win_first)
          w3_mma(tmem_base + static_cast<uint32_t>(kw * kW3AccCols), a0, a_hi_word, zero_lo, zero_hi, idesc0 + 3u * idesc_blk, 0u);
        if (leader && nb > 0 && !(a.dbg_flags & 4)) {
          // fully unrolled (rows of at most 64 positions): the rolled loop issued one MMA per ~75 clk instead of ~55
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
            if (k16 >= k16n) break;
            const uint32_t ao = static_cast<uint32_t>(k16 * 16 + kw), bo = static_cast<uint32_t>(k16 * 16);
            // smallest products first: x0 g2, x2 g0, x1 g1, x1 g0, x0 g1, x0 g0 (two-way split: x1 g0, x0 g1, x0 g0)
            if (NP == 3) {
              w3_mma(d, a0 + ao, a_hi_word, b0 + 2u * b_piece16 + bo, b_hi_word, idesc, 1u);
              w3_mma(d, a0 + 2u * a_piece16 + ao, a_hi_word, b0 + bo, b_hi_word, idesc, 1u);
              w3_mma(d, a0 + a_piece16 + ao, a_hi_word, b0 + b_piece16 + bo, b_hi_word, idesc, 1u);
            }
            w3_mma(d, a0 + a_piece16 + ao, a_hi_word, b0 + bo, b_hi_word, idesc, 1u);
            w3_mma(d, a0 + ao, a_hi_word, b0 + b_piece16 + bo, b_hi_word, idesc, 1u);
            w3_mma(d, a0 + ao, a_hi_word, b0 + bo, b_hi_word, idesc, 1u);
          }
        }
        if (win_last && leader) tc::umma_commit(afull + kw);
      }
      if (leader) tc::umma_commit(empty + stage);
      if (win_last) ++nwin;
      __syncwarp();
    }
    }
  } else if (warp < 2 + kW3SplitWarps) {
    // =============================== split warps: fp32 rows -> three bf16 pieces in the operand layout ==============
    const int tid = threadIdx.x - 64;
    constexpr int NT = kW3SplitWarps * 32;
    // first item and per-iteration increments of this thread in the two item spaces (A: rows of 3 Wi, B: rows of WP)
    const int a_g0 = tid / (3 * Wi), a_r0 = tid % (3 * Wi), a_dg = NT / (3 * Wi), a_dr = NT % (3 * Wi);
    const int b_r0 = tid / WP, b_w0 = tid % WP, b_dr = NT / WP, b_dw = NT % WP;
    const int go8 = a.GOr / 2;
    const float sx = NP == 2 ? w3_exp2i(w3_scale_exp(__ldg(a.amax_x))) : 1.f;
    const float sg = NP == 2 ? w3_exp2i(w3_scale_exp(__ldg(a.amax_g))) : 1.f;
    uint32_t seq = 0;
    for (uint32_t ci = 0;; ++ci) {
    w3_next_range(ranges, ci, s_begin, s_end);
    if (s_begin >= s_end) break;
    W3Step st = w3_step(s_begin, a);
    for (long long s = s_begin; s < s_end; ++s, ++seq, w3_next(st, a)) {
      const uint32_t stage = seq % kW3Stages;
      tc::mbar_wait(empty + stage, ((seq / kW3Stages) & 1u) ^ 1u);  // the MMAs of the step that used these piece buffers are done
      const uint32_t rb = seq % RB, rph = (seq / RB) & 1u;
      tc::mbar_wait(raw_full + rb, rph);
      const float4* ar = reinterpret_cast<const float4*>(a_raw + rb * a_raw_span);
      uint4* ap0 = reinterpret_cast<uint4*>(a_s + (stage * NP) * a_piece);
      uint4* ap1 = reinterpret_cast<uint4*>(a_s + (stage * NP + 1u) * a_piece);
      uint4* ap2 = reinterpret_cast<uint4*>(a_s + (stage * NP + (NP - 1u)) * a_piece);
      const int row = 3 * Wi;  // elements of one fp32 channel group (three rows)
      // item i = (g8, r): no divisions in the loop -- (g8, r) advance incrementally from the thread's first item
      // (at most kW3Items items per thread: G8 <= 4 groups of 3 rows of <= 64 positions over 256 threads; all loads of a
      // thread are issued before the first conversion -- two warps per scheduler cannot hide a load-convert-store chain)
      {
        int g8 = a_g0, r = a_r0;
        const int n_items = (a.dbg_flags & 2) ? 0 : G8 * row;
        float4 lo[kW3Items], hi[kW3Items];
#pragma unroll
        for (int u = 0; u < kW3Items; ++u) {
          if (tid + u * NT < n_items) {
            lo[u] = ar[(2 * g8) * row + r];
            hi[u] = ar[(2 * g8 + 1) * row + r];
          }
          r += a_dr; g8 += a_dg;
          if (r >= row) { r -= row; ++g8; }
        }
#pragma unroll
        for (int u = 0; u < kW3Items; ++u) {
          const int i = tid + u * NT;
          if (i < n_items) {
            uint4 q0, q1, q2;
            if (NP == 3) {
              w3_split8(lo[u], hi[u], q0, q1, q2);
              ap0[i] = q0; ap1[i] = q1; ap2[i] = q2;
            } else {
              w3_split8h(lo[u], hi[u], sx, q0, q1);
              ap0[i] = q0; ap1[i] = q1;
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(raw_empty + rb);  // the input rows have been read: the next step's may land
      tc::mbar_wait(raw_full + RB + rb, rph);
      const int nb = st.kt_hi - st.kt_lo + 1;
      if (nb > 0 && !(a.dbg_flags & 2)) {
        const float4* br = reinterpret_cast<const float4*>(b_raw + rb * b_raw_span);
        uint4* bp0 = reinterpret_cast<uint4*>(b_s + (stage * NP) * b_piece);
        uint4* bp1 = reinterpret_cast<uint4*>(b_s + (stage * NP + 1u) * b_piece);
        uint4* bp2 = reinterpret_cast<uint4*>(b_s + (stage * NP + (NP - 1u)) * b_piece);
        const int nrow = nb * go8;  // (kt, g8) rows of WP positions
        int rw = b_r0, w = b_w0;
        float4 lo[kW3Items], hi[kW3Items];
        int o[kW3Items];
#pragma unroll
        for (int u = 0; u < kW3Items; ++u) {
          o[u] = -1;
          if (rw < nrow) {
            const int ktl = rw / go8;  // go8 <= 4, nb <= 3: a handful of values, the compiler turns this into compares
            const int g8 = rw - ktl * go8, kt = st.kt_lo + ktl;
            o[u] = (kt * GP8 + g8) * WP + w;
            lo[u] = hi[u] = make_float4(0.f, 0.f, 0.f, 0.f);  // positions beyond the row: the K padding of the operand
            if (w < a.Wo) {
              lo[u] = br[((2 * g8) * 3 + (2 - kt)) * a.Wo + w];
              hi[u] = br[((2 * g8 + 1) * 3 + (2 - kt)) * a.Wo + w];
            }
          }
          w += b_dw; rw += b_dr;
          if (w >= WP) { w -= WP; ++rw; }
        }
#pragma unroll
        for (int u = 0; u < kW3Items; ++u) {
          if (o[u] >= 0) {
            uint4 q0, q1, q2;
            if (NP == 3) {
              w3_split8(lo[u], hi[u], q0, q1, q2);
              bp0[o[u]] = q0; bp1[o[u]] = q1; bp2[o[u]] = q2;
            } else {
              w3_split8h(lo[u], hi[u], sg, q0, q1);
              bp0[o[u]] = q0; bp1[o[u]] = q1;
            }
          }
        }
      }
      tc::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) {
        tc::mbar_arrive(ready + stage);
        tc::mbar_arrive(raw_empty + RB + rb);  // the gradient rows have been read
      }
    }
    }
  } else {
    // =============================== accumulator drain (warps 10..13, one per TMEM lane quadrant) ===============================
    // per window and accumulator: tcgen05.ld, the accumulator goes back to the MMA warp at once, and the window's sums are
    // added to the CTA's private partial by fire-and-forget red.global.add.f32 (the first window stores): no read-modify-write
    // round trip, so four warps are enough (the first version needed a group of four per accumulator)
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
    float* mine = a.partial + static_cast<size_t>(blockIdx.x) * 3 * 128 * kW3AccCols + row;  // [kw][column][row]
    // the partials (22 MB over the grid) are touched every window while ~100 MB of operands stream through L2 in between:
    // without a retention hint they were evicted and every update went to DRAM
    uint64_t keep;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
    // every thread zeroes exactly the elements it later adds to (a CTA that never gets a range still owes the reduce zeros)
#pragma unroll 1
    for (int c = 0; c < 3 * kW3AccCols; ++c) w3_st_keep(mine + static_cast<size_t>(c) * 128, 0.f, keep);
    long long wdx = 0;
    for (uint32_t ci = 0;; ++ci) {
      w3_next_range(ranges, ci, s_begin, s_end);
      if (s_begin >= s_end) break;
      const long long nwin_range = (s_end - s_begin + kW3Flush - 1) / kW3Flush;
      for (long long wr = 0; wr < nwin_range; ++wr, ++wdx) {
#pragma unroll 1
        for (int kw = 0; kw < 3; ++kw) {
          tc::mbar_wait(afull + kw, static_cast<uint32_t>(wdx & 1));
          tc::tc_fence_after();
          uint32_t v[3][32];
          if (!(a.dbg_flags & 8)) {
#pragma unroll
            for (int c = 0; c < 3; ++c) tc::tmem_ld_32x32(lane_addr + static_cast<uint32_t>(kw * kW3AccCols + c * 32), v[c]);
            tc::tmem_ld_wait();
          }
          tc::tc_fence_before();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(aempty + kw);
          if (a.dbg_flags & (8 | 16)) continue;  // 16: accumulators read and released, no partial traffic
          // partial layout [kw][column][row]: a warp's 32 rows of one column are 128 contiguous bytes
          float* dst = mine + static_cast<size_t>(kw) * 128 * kW3AccCols;
#pragma unroll
          for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int j = 0; j < 32; ++j) w3_red_keep(dst + static_cast<size_t>(c * 32 + j) * 128, __uint_as_float(v[c][j]), keep);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

// dW[co][ci][kt][kh][kw] = sum over CTAs of partial[cta][kw][kt*CoP + co][(3*(ci/8) + kh)*8 + ci%8];
// db[co] = sum over CTAs of partial[cta][0][kt_bias*CoP + co][(3*G8)*8]  (the row of ones against the time tap kt_bias = pad_t,
// the one tap for which every output plane t = p meets an existing input plane exactly once)
// amax != null (two-way fp16 split): the sums are in the scaled domain, 2^ex x 2^eg times too large (db: 2^eg)
__global__ void wgrad_bf16x3_reduce_kernel(const float* __restrict__ partial, int ncta, float* __restrict__ dw, float* __restrict__ db,
                                           int Ci, int Co, int G8, int CoP, int kt_bias, const float* __restrict__ amax_x,
                                           const float* __restrict__ amax_g) {
  const float ux = amax_x ? w3_exp2i(-w3_scale_exp(__ldg(amax_x))) : 1.f, ug = amax_g ? w3_exp2i(-w3_scale_exp(__ldg(amax_g))) : 1.f;
  // one thread per element of the partial layout [kw][column][row]: a warp reads 32 consecutive rows = 128 contiguous bytes
  // of every CTA's partial (indexing by dW element instead cost 32 sectors per load: 15 us per layer, L2 bound)
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 3 * kW3AccCols * 128) return;
  const int row = idx & 127, col = (idx >> 7) % kW3AccCols, kw = idx / (128 * kW3AccCols);
  const int kt = col / CoP, co = col - kt * CoP;
  const int mg = row >> 3, c8 = row & 7;  // M-group 3 g8 + kh, or the group of ones
  const bool is_bias = (mg == 3 * G8) && c8 == 0 && kw == 0 && kt == kt_bias && db != nullptr;
  const int g8 = mg / 3, kh = mg - 3 * g8, ci = g8 * 8 + c8;
  const bool is_w = mg < 3 * G8 && ci < Ci;
  if (co >= Co || kt >= 3 || !(is_w || is_bias)) return;
  const size_t per_cta = static_cast<size_t>(3) * 128 * kW3AccCols;
  const float* src = partial + idx;
  float s = 0.f;
  for (int c = 0; c < ncta; ++c) s += src[c * per_cta];  // fixed order: deterministic
  if (is_w) dw[((static_cast<size_t>(co) * Ci + ci) * 3 + kt) * 9 + kh * 3 + kw] = (s * ux) * ug;
  else db[co] = s * ug;
}

int g_w3_dbg_flags = 0;  // set through pvb200_debug_set_wgrad_flags (tools only)
extern int g_dynamic_tiles;  // runtime.cu: pvb200_set_dynamic_tiles

// 5-D fp32 tensor map, dense (strides follow the dimensions), no swizzle / interleave, zero fill outside the tensor.  The
// encoder is a host-only driver function, fetched through the runtime so that the library keeps linking only libcudart.
static int w3_make_tensor_map(CUtensorMap* tm, const void* base, const unsigned long long dims[5], const unsigned box[5]) {
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
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  unsigned long long stride = 4;
  for (int i = 0; i < 5; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    stride *= dims[i];
    if (i < 4) gstr[i] = stride;  // byte stride of dimension i + 1
  }
  return static_cast<int>(encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(base), gdim, gstr, bx, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
}

static int w3_groups(int C) { return 2 * ceil_div(C, 8); }
static int w3_cop(int Co) { return Co <= 16 ? 16 : 32; }

static size_t w3_smem_bytes(int G, int GOr, int Wi, int Wo, int CoP, int NP = 3) {
  const size_t WP = round_up(Wo, 16);
  const size_t a_piece = round_up(static_cast<size_t>(3 * (G / 2) + 4) * Wi * 16, static_cast<size_t>(128));
  const size_t b_piece = static_cast<size_t>(3) * (CoP / 8) * WP * 16;
  const size_t RB = NP == 2 ? 2 : 1;
  return 512 + kW3Stages * NP * (a_piece + b_piece) +
         RB * (round_up(static_cast<size_t>(G) * 3 * Wi * 16, static_cast<size_t>(128)) + round_up(static_cast<size_t>(3) * GOr * Wo * 16, static_cast<size_t>(128)));
}

static int w3_supported(int Cin, int Cout, int Hi, int Wi, int NP) {
  if (Cin <= 0 || Cout <= 0 || Cin > 32 || Cout > 32 || Hi < 3 || Wi < 3) return 0;
  const int G = w3_groups(Cin);
  // the M = 128 instruction reads 16 row groups at stride Wi*16 from the start of an A piece: it must stay inside the allocation
  const size_t smem = w3_smem_bytes(G, w3_groups(Cout), Wi, Wi - 2, w3_cop(Cout), NP);
  const size_t a_piece = round_up(static_cast<size_t>(3 * (G / 2) + 4) * Wi * 16, static_cast<size_t>(128));
  const size_t last_a = 512 + (NP * kW3Stages - 1) * a_piece;
  const size_t reach = last_a + static_cast<size_t>(16) * Wi * 16 + static_cast<size_t>(round_up(Wi, 16) + 16) * 16;
  return (smem <= 227 * 1024 && reach <= smem && Wi <= 64) ? 1 : 0;  // Wi * 4 floats = the 256-element limit of a TMA box dimension
}

static int w3_launch(const char* who, int NP, const float* amax_x, const float* amax_g, const float* xb, const float* gzb, int gz_pad, float* dw, float* db,
                     void* workspace, size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout, int pad_t,
                     cudaStream_t stream) {
  PVB_REQUIRE(xb && gzb && dw, "%s: null pointer", who);
  PVB_REQUIRE(B > 0 && gz_pad >= 0 && (pad_t == 0 || pad_t == 1), "%s: bad argument", who);
  PVB_REQUIRE(w3_supported(Cin, Cout, Hi, Wi, NP), "%s: Cin=%d Cout=%d plane %dx%d is not supported by the tensor-core weight "
              "gradient (use pvb200_conv3d_wgrad_f32)", who, Cin, Cout, Hi, Wi);
  W3Args a;
  a.B = B; a.G = w3_groups(Cin); a.Ti = Ti; a.Hi = Hi; a.Wi = Wi;
  a.GOr = w3_groups(Cout); a.CoP = w3_cop(Cout);
  a.To = Ti + 2 * pad_t - 2; a.Ho = Hi - 2; a.Wo = Wi - 2; a.WP = round_up(a.Wo, 16);
  PVB_REQUIRE(a.To > 0, "%s: input too short", who);
  a.plane_off = -pad_t;
  a.gz_pad = gz_pad;
  a.steps = static_cast<long long>(B) * Ti * a.Ho;
  a.dbg_flags = g_w3_dbg_flags;
  a.amax_x = amax_x; a.amax_g = amax_g;
  const int sms = sm_count();
  PVB_REQUIRE(sms > 0, "%s: no CUDA device", who);
  long long grid = a.steps < sms ? a.steps : sms;
  const size_t need = static_cast<size_t>(grid) * 3 * 128 * kW3AccCols * sizeof(float) + 64;  // partials + the chunk counter
  if (!workspace || workspace_bytes < need) {
    set_error("%s: workspace too small (%zu < %zu bytes)", who, workspace_bytes, need);
    return PVB200_ERR_WORKSPACE;
  }
  a.sched = nullptr;
  if (g_dynamic_tiles) {
    a.sched = reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + need - 64);
    PVB_CUDA(cudaMemsetAsync(a.sched, 0, sizeof(int), stream));
  }
  PVB_REQUIRE(reinterpret_cast<uintptr_t>(xb) % 16 == 0 && reinterpret_cast<uintptr_t>(gzb) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(workspace) % 16 == 0, "%s: pointers must be 16-byte aligned", who);
  a.partial = static_cast<float*>(workspace);
  const size_t smem = w3_smem_bytes(a.G, a.GOr, Wi, a.Wo, a.CoP, NP);
  // tensor maps (fp32 elements, innermost dimension = one row of 4-channel elements): x [B][G][Ti][Hi][Wi*4] with the box
  // [1][G][1][3][Wi*4]; gz [B][GOr][Tz][Hz][Wz*4] with the box [1][GOr][3][1][Wo*4]
  CUtensorMap tm_x, tm_gz;
  {
    const unsigned long long Tz = a.To + 2 * gz_pad, Hz = a.Ho + 2 * gz_pad, Wz = a.Wo + 2 * gz_pad;
    const unsigned long long xd[5] = {static_cast<unsigned long long>(Wi) * 4, static_cast<unsigned long long>(Hi), static_cast<unsigned long long>(Ti),
                                      static_cast<unsigned long long>(a.G), static_cast<unsigned long long>(B)};
    const unsigned xb_[5] = {static_cast<unsigned>(Wi) * 4, 3, 1, static_cast<unsigned>(a.G), 1};
    const unsigned long long gd[5] = {Wz * 4, Hz, Tz, static_cast<unsigned long long>(a.GOr), static_cast<unsigned long long>(B)};
    const unsigned gb_[5] = {static_cast<unsigned>(a.Wo) * 4, 1, 3, static_cast<unsigned>(a.GOr), 1};
    const int r1 = w3_make_tensor_map(&tm_x, xb, xd, xb_);
    const int r2 = w3_make_tensor_map(&tm_gz, gzb, gd, gb_);
    PVB_REQUIRE(r1 == 0 && r2 == 0, "%s: cuTensorMapEncodeTiled failed (%d, %d)", who, r1, r2);
  }
  if (NP == 3) {
    PVB_CUDA(cudaFuncSetAttribute(conv3d_wgrad_bf16x3_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3d_wgrad_bf16x3_kernel<3><<<static_cast<unsigned>(grid), kW3Threads, smem, stream>>>(a, tm_x, tm_gz);
  } else {
    PVB_CUDA(cudaFuncSetAttribute(conv3d_wgrad_bf16x3_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3d_wgrad_bf16x3_kernel<2><<<static_cast<unsigned>(grid), kW3Threads, smem, stream>>>(a, tm_x, tm_gz);
  }
  PVB_LAUNCHED(who);
  wgrad_bf16x3_reduce_kernel<<<3 * kW3AccCols, 128, 0, stream>>>(a.partial, static_cast<int>(grid), dw, db, Cin, Cout, a.G / 2,
                                                                       a.CoP, pad_t, NP == 2 ? amax_x : nullptr, NP == 2 ? amax_g : nullptr);
  PVB_LAUNCHED("wgrad_bf16x3_reduce");
  return PVB200_OK;
}

// largest magnitude of a tensor, as the bit pattern of a non-negative float (atomicMax on the unsigned view)
__global__ void __launch_bounds__(256) absmax_f32_kernel(const float4* __restrict__ x, long long n4, const float* __restrict__ tail,
                                                         int ntail, unsigned* __restrict__ out) {
  float m = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * 256) {
    const float4 v = __ldg(x + i);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  if (blockIdx.x == 0 && static_cast<int>(threadIdx.x) < ntail) m = fmaxf(m, fabsf(tail[threadIdx.x]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));
}

}  // namespace pvb

extern "C" {

/* tools only (not declared in pvb200.h): switch parts of the weight-gradient kernel off to time the others */
void pvb200_debug_set_wgrad_flags(int f) { pvb::g_w3_dbg_flags = f; }

/* 1 when the tensor-core weight gradient takes this layer (channel counts <= 32, rows that fit shared memory) */
int pvb200_conv3d_wgrad_bf16x3_supported(int Cin, int Cout, int Hi, int Wi) { return pvb::w3_supported(Cin, Cout, Hi, Wi, 3); }

size_t pvb200_conv3d_wgrad_bf16x3_workspace_bytes(void) {
  int sms = pvb::sm_count();
  if (sms <= 0) sms = 148;
  return static_cast<size_t>(sms) * 3 * 128 * pvb::kW3AccCols * sizeof(float) + 64;
}

/* dw [Cout][Cin][3][3][3], db [Cout] (or null) from x blocked fp32 [B][G(Cin)][Ti][Hi][Wi][4] and the pre-activation gradient
 * gz blocked fp32, zero-padded by gz_pad on T, H, W ([B][G(Cout)][To+2p][Ho+2p][Wo+2p][4], To = Ti + 2 pad_t - 2): the
 * padded tensor the data gradient reads (gz_pad = 2) or a plain one (gz_pad = 0) */
int pvb200_conv3d_wgrad_bf16x3(const float* xb, const float* gzb, int gz_pad, float* dw, float* db, void* workspace,
                               size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout, int pad_t,
                               pvb200_stream_t stream) {
  return pvb::w3_launch("conv3d_wgrad_bf16x3", 3, nullptr, nullptr, xb, gzb, gz_pad, dw, db, workspace, workspace_bytes, B, Cin, Ti, Hi, Wi, Cout,
                        pad_t, pvb::as_stream(stream));
}

/* the same through the two-way fp16 split (three products instead of six).  amax_x, amax_gz: device scalars max |x|, max |gz|
 * of the two tensors (pvb200_absmax_f32 or the amax_out of the kernel that wrote them): the operands are scaled by powers of
 * two into fp16's range inside the kernel and the result is scaled back; workspace and shape limits as the bf16x3 entry */
int pvb200_conv3d_wgrad_f16x2(const float* xb, const float* gzb, int gz_pad, const float* amax_x, const float* amax_gz, float* dw,
                              float* db, void* workspace,
                              size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout, int pad_t,
                              pvb200_stream_t stream) {
  PVB_REQUIRE(amax_x && amax_gz, "conv3d_wgrad_f16x2: null amax");
  return pvb::w3_launch("conv3d_wgrad_f16x2", 2, amax_x, amax_gz, xb, gzb, gz_pad, dw, db, workspace, workspace_bytes, B, Cin, Ti, Hi, Wi, Cout,
                        pad_t, pvb::as_stream(stream));
}

/* *out = max(*out, max |x[i]|) for a device scalar the caller has zeroed (several tensors may share it) */
int pvb200_absmax_f32(const float* x, long long n, float* out, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(x && out && n > 0, "absmax_f32: bad argument");
  PVB_REQUIRE(reinterpret_cast<uintptr_t>(x) % 16 == 0, "absmax_f32: x must be 16-byte aligned");
  const int sms = sm_count();
  PVB_REQUIRE(sms > 0, "absmax_f32: no CUDA device");
  const long long n4 = n / 4;
  long long grid = ceil_div(n4 > 0 ? n4 : 1LL, 256LL * 8);
  if (grid > 8LL * sms) grid = 8LL * sms;
  absmax_f32_kernel<<<static_cast<unsigned>(grid), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x), n4, x + n4 * 4,
                                                                                 static_cast<int>(n - n4 * 4),
                                                                                 reinterpret_cast<unsigned*>(out));
  PVB_LAUNCHED("absmax_f32");
  return PVB200_OK;
}

}  // extern "C"
