// conv3d_igemm_tf32x3.cu -- a3/a4/a11 in fp32 MODE on the tensor cores: Conv3d 3x3x3 forward and data gradient as an
// implicit GEMM of split-precision tcgen05 MMAs, fp32-class accuracy (<= 1e-5): the 3xTF32 split (kind::tf32) described
// first, and -- in the CTA-pair kernel, when the caller passes the input tensor's largest magnitude (amax_in) -- the
// two-way fp16 split of scaled operands (kind::f16, K = 16 channels per MMA, half the MMAs; see t3_scale_exp below), which
// is what the model runs for Cout > 16.  The layout kernels at the end report the largest magnitude they write (amax_out).
//
// Reference call sites: predict_pv_yield/models/conv3d/model.py:80-90,117-120 (and their autograd).
//
// Arithmetic (probed on the B200, tools/probe/tf32_acc_probe.cu -> profiles/tf32_probe_r02.txt):
//   * the tensor core TRUNCATES 32-bit operands to TF32 (10 mantissa bits), sums the K = 8 products of one MMA exactly
//     and adds them to the fp32 accumulator in TMEM with rounding TOWARD ZERO.
//   * so x = x_hi + x_lo with x_hi = the value as stored (the hardware truncates it) and x_lo = x - trunc(x), exact in
//     fp32; the same for w; and  x.w ~= x_hi.w_lo + x_lo.w_hi + x_hi.w_hi  (the dropped x_lo.w_lo is 2^-22 relative).
//   * the toward-zero accumulator is a BIAS of half an ulp of the accumulator per MMA: a chain of 324 accumulating MMAs
//     (K = 864, three terms) costs 1e-5 per layer (tools/tf32x3_study.py).  Two measures bring it to ~1e-6:
//       - the three time taps of a plane are NOT summed in the tensor core: every input plane owns a fresh accumulator
//         block (48 columns: kt x 16 output channels), the three blocks an output plane needs are added in the epilogue
//         in fp32 round-to-nearest;
//       - inside a block the two correction terms of ALL channels and taps are issued FIRST, while the accumulator
//         still holds correction-sized values (their half-ulps are 2^-11 smaller), the 36 main-term MMAs last.
//     The longest chain of full-size roundings is therefore 36 (9 taps x 4 channel steps), three of them per output.
//
// Structure (the scatter form of conv3d_igemm_bf16.cu, see there for the layout argument):
//   * activations are blocked fp32 [B][G][T][H][W][4] (16 bytes = 4 channels innermost, G even): positions flattened with
//     the input pitch make the A operand of tap (kh, kw) the contiguous run starting at q0 + kh*Wi + kw of a plane -- the
//     SWIZZLE_NONE K-major canonical layout as stored; a tap is a descriptor start address, K = 8 = two channel groups.
//   * the loop runs over INPUT planes: plane p contributes D[r, kt*16 + co] to the outputs t = p - kt (N = 48).
//   * hi + lo weights of all 27 taps are 221 KB -- more than shared memory -- so a CTA owns HALF of the output channels
//     (2 x 55 KB resident); even / odd CTAs of the persistent grid walk the same tile list for the two halves at the same
//     time, the second read of the activations hits L2.
//   * the TMA engine (cp.async.bulk, mbarrier complete_tx) stages one (plane, 8-channel step) segment per ring slot; four
//     "split" warps derive x_lo next to it in shared memory (generic proxy -> fence.proxy.async -> mbarrier); one elected
//     thread issues the 27 x KS x 3 MMAs of a plane; four epilogue warps add the three blocks of an output plane from
//     TMEM (tcgen05.ld), bias / ReLU (or the ReLU mask of the data gradient) and store the blocked and / or NCDHW copy.
// Warp roles (320 threads): warp 0 producer, warp 1 MMA issuer + TMEM owner, warps 2-5 split, warps 6-9 epilogue.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pvb {

constexpr int kT3Threads = 320;
constexpr int kT3MaxSlots = 12;  // ring of (plane, channel step) segments
constexpr int kT3TileM = 128;    // MMA rows = output positions per tile
constexpr int kT3Blocks = 10;    // accumulator blocks (one per input plane) resident in TMEM: 10 x 48 of 512 columns
constexpr int kT3N = 48;         // MMA N = 3 time taps x 16 output channels
constexpr int kT3Half = 16;      // output channels per CTA

struct T3Args {
  const uint4* x;     // [B][G][Ti][Hi][Wi] 16-byte elements (4 fp32 channels)
  const uint4* wq;    // [half][hi, lo][9 (kh,kw)][KS][2][48] 16-byte elements: 4 input channels of one (kt, output channel)
  const float* bias;  // [Co] or null
  const uint4* mask;  // blocked [B][GO][To][Ho][Wo] or null (ReLU-mask source of the data gradient)
  uint4* y_blk;       // [B][GO][To+2p][Ho+2p][Wo+2p] or null
  float* y_nc;        // [B][Co][To][Ho][Wo] or null
  const float* amax_in;  // two-way fp16 split only: max |x| of the input tensor and max |w| (device scalars)
  const float* amax_w;
  unsigned* amax_out; // or null: atomicMax of the bit pattern of max |y| (a non-negative float) over everything written
  int B, G, Ti, Hi, Wi;
  int Co, GO, To, Ho, Wo;
  int nhalf;        // 1 (Co <= 16) or 2
  int out_pad, relu;
  int zero_planes;  // the first / last `zero_planes` input time planes are all zero (padded gz): never staged, no MMAs
  int plane_off;    // input plane of output t, tap kt: t + kt + plane_off
  int NP;           // staged positions per (plane, channel group)
  int tiles_q;      // q tiles per output plane
  int nslot;        // ring slots that fit in shared memory
  long long tiles;  // B * tiles_q * To
};

__device__ __forceinline__ float t3_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

// largest magnitude a warp has seen -> one atomicMax per warp on the unsigned view of a non-negative float
__device__ __forceinline__ void t3_publish_amax(unsigned* out, float am) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
  if ((threadIdx.x & 31) == 0 && am > 0.f) atomicMax(out, __float_as_uint(am));
}

// ---- two-way fp16 split (CTA-pair kernel, F16 = true) -------------------------------------------------------------------
// s v = h0 + h1 + O(2^-22 |s v|) with s the power of two that brings the tensor's largest magnitude to [2^14, 2^15): the
// same 11 + 11 significand bits as the TF32 split, but both pieces are 16-bit operands of kind::f16 -- K = 16 channels per
// MMA instead of 8, half the MMAs and half the operand bytes per plane.  fp16 has 5 exponent bits, hence the scaling; the
// epilogue multiplies the sums by 2^-(ex + ew) (exact).  Values below 2^-18 of the tensor's maximum lose relative (not
// absolute) precision: their second piece is an fp16 subnormal with an absolute step of 2^-39 of the maximum.
__device__ __forceinline__ int t3_scale_exp(float amax) {
  if (!(amax > 0.f) || !(amax <= 3.0e38f)) return 0;
  const int e = 14 - (static_cast<int>((__float_as_uint(amax) >> 23) & 0xffu) - 127);
  return e < -100 ? -100 : (e > 100 ? 100 : e);
}
__device__ __forceinline__ float t3_exp2i(int e) { return __uint_as_float(static_cast<uint32_t>(e + 127) << 23); }
__device__ __forceinline__ void t3_split2h(float v0, float v1, uint32_t& p0, uint32_t& p1) {
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(v1), "f"(v0));  // high half <- first source
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&p0));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(v1 - f.y), "f"(v0 - f.x));
}
__device__ __forceinline__ void t3_split8h(const float4 lo4, const float4 hi4, float s, uint4& q0, uint4& q1) {
  t3_split2h(lo4.x * s, lo4.y * s, q0.x, q1.x);
  t3_split2h(lo4.z * s, lo4.w * s, q0.y, q1.y);
  t3_split2h(hi4.x * s, hi4.y * s, q0.z, q1.z);
  t3_split2h(hi4.z * s, hi4.w * s, q0.w, q1.w);
}

// the same for a whole block of 256 threads that ALL call it (layout / normalise kernels: one atomic per block, not per warp)
__device__ __forceinline__ void t3_publish_amax_block(unsigned* out, float am) {
  __shared__ unsigned red[8];
  const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(am));  // non-negative floats order like their bit patterns
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned r = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) r = r > red[i] ? r : red[i];
    if (r) atomicMax(out, r);
  }
}

// weights fp32 [Co][Ci][27] -> fp16 pieces [rank][piece][(kh,kw)][ks16][kg][48 columns of the N = 96 = (kt, co) operand][8 ci],
// scaled by the power of two of max |w| -- which every block computes for itself (27 K values) and block 0 leaves at
// `amax_w` for the main kernel's epilogue
__global__ void __launch_bounds__(256) t3_weight_prep_f16_kernel(const float* __restrict__ w, uint32_t* __restrict__ wq, int Ci_role,
                                                                 int Co_role, int KS16, long long s_co, long long s_ci, int flip,
                                                                 int nw, float* __restrict__ amax_w) {
  __shared__ float red[8];
  float m = 0.f;
  for (int i = threadIdx.x; i < nw; i += 256) m = fmaxf(m, fabsf(__ldg(w + i)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  if (blockIdx.x == 0 && threadIdx.x == 0) *amax_w = m;
  const float sw = t3_exp2i(t3_scale_exp(m));
  const int per_piece = 9 * KS16 * 2 * kT3N * 4;  // 32-bit words (two channels each) of one piece of one rank
  const int total = 2 * 2 * per_piece;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int e2 = idx & 3;  // channel pair inside the 8-channel group
    const int n = (idx >> 2) % kT3N;
    const int kg = (idx / (4 * kT3N)) & 1;
    const int ks = (idx / (8 * kT3N)) % KS16;
    const int hw = (idx / (8 * kT3N * KS16)) % 9;
    const int piece = (idx / per_piece) & 1;
    const int rank = idx / (2 * per_piece);
    const int col = rank * kT3N + n;  // column of the 96-wide operand
    const int kt = col >> 5, co = col & 31;
    const int tap = kt * 9 + hw;
    float v[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int ci = (ks * 2 + kg) * 8 + e2 * 2 + e;
      v[e] = (co < Co_role && ci < Ci_role) ? w[co * s_co + ci * s_ci + (flip ? 26 - tap : tap)] * sw : 0.f;
    }
    uint32_t p0, p1;
    t3_split2h(v[0], v[1], p0, p1);
    wq[idx] = piece ? p1 : p0;
  }
}

// weights fp32 [Co][Ci][27] -> [half][term][(kh,kw)][ks][kg][kt*16 + c][4 ci]; term 0 = the value as stored (the tensor
// core truncates it to TF32), term 1 = the exact residual w - trunc(w); flipped / transposed roles for the data gradient
// pair = 1 (CTA-pair kernel, Cout = 32): "half" is the CTA rank and holds columns 48*rank .. 48*rank+47 of the N = 96 = (kt, co)
// operand instead of all three time taps of 16 output channels
__global__ void t3_weight_prep_kernel(const float* __restrict__ w, float* __restrict__ wq, int Ci_role, int Co_role, int KS,
                                      int nhalf, long long s_co, long long s_ci, int flip, int pair) {
  const int per_term = 9 * KS * 2 * kT3N * 4;
  const int total = nhalf * 2 * per_term;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int e4 = idx & 3;
    const int n = (idx >> 2) % kT3N;
    const int kg = (idx / (4 * kT3N)) & 1;
    const int ks = (idx / (8 * kT3N)) % KS;
    const int hw = (idx / (8 * kT3N * KS)) % 9;
    const int term = (idx / per_term) & 1;
    const int half = idx / (2 * per_term);
    int kt = n / kT3Half, c = n - kt * kT3Half;
    int co = half * kT3Half + c;
    if (pair) {
      const int col = half * kT3N + n;  // column of the 96-wide operand
      kt = col >> 5;
      co = col & 31;
    }
    const int ci = (ks * 2 + kg) * 4 + e4;
    const int tap = kt * 9 + hw;
    float v = 0.f;
    if (co < Co_role && ci < Ci_role) v = w[co * s_co + ci * s_ci + (flip ? 26 - tap : tap)];
    wq[idx] = term ? (v - t3_trunc(v)) : v;
  }
}

__device__ __forceinline__ void t3_mma(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate));
}

// the CTA-pair MMA of kind::f16 (see tc::umma_tf32_pair_lohi)
__device__ __forceinline__ void t3_mma_f16_pair(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate));
}

// A "run" = consecutive output time steps of one (sample, q-tile) column handled by one CTA.  Output tiles are numbered
// g = (b * tiles_q + qt) * To + t and split evenly over the persistent CTAs of a half; every role walks the same sequence
// of runs.  A run of n outputs consumes the n + 2 input planes i = 0 .. n+1 (absolute plane t0 + i + plane_off).
struct T3Run {
  int b, qt, t0, n;
};
__device__ __forceinline__ T3Run t3_run(long long g, long long g_end, int To, int tiles_q) {
  T3Run r;
  const long long col = g / To;
  r.t0 = static_cast<int>(g - col * To);
  r.qt = static_cast<int>(col % tiles_q);
  r.b = static_cast<int>(col / tiles_q);
  const long long left = g_end - g;
  r.n = static_cast<int>(left < (To - r.t0) ? left : (To - r.t0));
  return r;
}

template <int KS>
__global__ void __launch_bounds__(kT3Threads, 1) conv3d_igemm_tf32x3_kernel(const T3Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [12] segment landed (TMA)
  uint64_t* ready = full + kT3MaxSlots;                 // [12] x_lo written next to it (split warps)
  uint64_t* empty = ready + kT3MaxSlots;                // [12] segment consumed (MMA commit)
  uint64_t* wfull = empty + kT3MaxSlots;                // [1]  weights landed
  uint64_t* bfull = wfull + 1;                          // [10] accumulator block complete
  uint64_t* bempty = bfull + kT3Blocks;                 // [10] accumulator block read out
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bempty + kT3Blocks);
  float* bias_s = reinterpret_cast<float*>(smem + 512);  // [16] of this half
  uint8_t* w_s = smem + 640;
  constexpr uint32_t w_term_bytes = 9u * KS * 2u * kT3N * 16u;  // one of hi / lo
  constexpr uint32_t w_bytes = 2u * w_term_bytes;
  const uint32_t slot_term = 2u * static_cast<uint32_t>(a.NP) * 16u;  // hi (or lo) part of a slot: two channel groups
  const uint32_t slot_bytes = 2u * slot_term;
  uint8_t* slot_s = w_s + ((w_bytes + 127u) & ~127u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = a.nhalf == 2 ? static_cast<int>(blockIdx.x & 1u) : 0;
  const long long cta = a.nhalf == 2 ? (blockIdx.x >> 1) : blockIdx.x;
  const long long ncta = a.nhalf == 2 ? (gridDim.x >> 1) : gridDim.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kT3MaxSlots; ++i) { tc::mbar_init(full + i, 1); tc::mbar_init(ready + i, 4); tc::mbar_init(empty + i, 1); }
    tc::mbar_init(wfull, 1);
    for (int i = 0; i < kT3Blocks; ++i) { tc::mbar_init(bfull + i, 1); tc::mbar_init(bempty + i, 4); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, 512);
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kT3Half) {
    const int c = half * kT3Half + (threadIdx.x - 64);
    bias_s[threadIdx.x - 64] = (a.bias && c < a.Co) ? __ldg(a.bias + c) : 0.f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const long long g_begin = a.tiles * cta / ncta;
  const long long g_end = a.tiles * (cta + 1) / ncta;
  const long long in_plane = static_cast<long long>(a.Hi) * a.Wi;
  const uint32_t nslot = static_cast<uint32_t>(a.nslot);

  if (warp == 0) {
    // =============================== producer ===============================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(wfull, w_bytes);
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wq) + static_cast<size_t>(half) * w_bytes;
      for (uint32_t off = 0; off < w_bytes; off += 32768u) {
        const uint32_t n = (w_bytes - off < 32768u) ? (w_bytes - off) : 32768u;
        tc::bulk_g2s(w_s + off, wsrc + off, n, wfull);
      }
      uint32_t seq = 0;
      for (long long g = g_begin; g < g_end;) {
        const T3Run r = t3_run(g, g_end, a.To, a.tiles_q);
        const int q0 = r.qt * kT3TileM;
        const long long avail = in_plane - q0;
        const uint32_t npos = static_cast<uint32_t>(avail < a.NP ? avail : a.NP);
        for (int i = 0; i < r.n + 2; ++i) {
          const int pa = r.t0 + i + a.plane_off;
          if (pa < a.zero_planes || pa >= a.Ti - a.zero_planes) continue;
#pragma unroll 1
          for (int ks = 0; ks < KS; ++ks) {
            const uint32_t slot = seq % nslot;
            tc::mbar_wait(empty + slot, ((seq / nslot) & 1u) ^ 1u);
            ++seq;
            tc::mbar_arrive_expect_tx(full + slot, npos * 32u);
#pragma unroll
            for (int kg = 0; kg < 2; ++kg) {
              const uint4* src = a.x + ((static_cast<long long>(r.b) * a.G + (ks * 2 + kg)) * a.Ti + pa) * in_plane + q0;
              tc::bulk_g2s(slot_s + slot * slot_bytes + static_cast<uint32_t>(kg) * a.NP * 16u, src, npos * 16u, full + slot);
            }
          }
        }
        g += r.n;
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const bool leader = tc::elect_one();
    const uint32_t a_lbo = static_cast<uint32_t>(a.NP) * 16u;  // between the two channel groups of a K = 8 step
    const uint32_t b_lbo = static_cast<uint32_t>(kT3N) * 16u;
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);  // SBO = 128 B | descriptor version
    const uint32_t a_lo_base = ((a_lbo >> 4) << 16);
    const uint32_t b_lo_base = ((b_lbo >> 4) << 16) | ((tc::smem_u32(w_s) >> 4) & 0x3fffu);
    const uint32_t slot_addr16 = tc::smem_u32(slot_s) >> 4;
    const uint32_t slot_16 = slot_bytes >> 4, slot_term16 = slot_term >> 4;
    constexpr uint32_t w_term16 = w_term_bytes >> 4;
    constexpr uint32_t b_step16 = 2u * kT3N;  // one (tap, channel step) tile = 2 x 48 x 16 B
    uint32_t tap16[9];
#pragma unroll
    for (int hw = 0; hw < 9; ++hw) tap16[hw] = static_cast<uint32_t>((hw / 3) * a.Wi + (hw % 3));
    const uint32_t idesc = tc::umma_idesc(128, kT3N, /*TF32*/ 2, /*K-major*/ 0, 0);
    tc::mbar_wait(wfull, 0);
    uint32_t seq = 0;     // staged segments consumed so far
    uint32_t blk_ph = 0;  // bit b = uses of accumulator block b so far, mod 2
    uint32_t pc0 = 0;     // run-plane counter at the start of the run (block of run plane i = (pc0 + i) % kT3Blocks)
    for (long long g = g_begin; g < g_end;) {
      const T3Run r = t3_run(g, g_end, a.To, a.tiles_q);
      for (int i = 0; i < r.n + 2; ++i) {
        const int pa = r.t0 + i + a.plane_off;
        if (pa < a.zero_planes || pa >= a.Ti - a.zero_planes) continue;
        const uint32_t blk = (pc0 + static_cast<uint32_t>(i)) % kT3Blocks;
        tc::mbar_wait(bempty + blk, ((blk_ph >> blk) & 1u) ^ 1u);
        tc::tc_fence_after();
        const uint32_t d = tmem_base + blk * kT3N;
        uint32_t a_seg[KS];
        // correction terms first: x_hi.w_lo and x_lo.w_hi of every tap and channel step
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const uint32_t s = seq + ks;
          const uint32_t slot = s % nslot;
          tc::mbar_wait(ready + slot, (s / nslot) & 1u);
          tc::tc_fence_after();
          a_seg[ks] = a_lo_base | ((slot_addr16 + slot * slot_16) & 0x3fffu);
          if (leader) {
#pragma unroll
            for (int hw = 0; hw < 9; ++hw) {
              const uint32_t b_t = b_lo_base + static_cast<uint32_t>(hw * KS + ks) * b_step16;
              t3_mma(d, a_seg[ks] + tap16[hw], desc_hi, b_t + w_term16, desc_hi, idesc, (ks | hw) ? 1u : 0u);
              t3_mma(d, a_seg[ks] + slot_term16 + tap16[hw], desc_hi, b_t, desc_hi, idesc, 1u);
            }
          }
        }
        // main term last
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          if (leader) {
#pragma unroll
            for (int hw = 0; hw < 9; ++hw)
              t3_mma(d, a_seg[ks] + tap16[hw], desc_hi, b_lo_base + static_cast<uint32_t>(hw * KS + ks) * b_step16, desc_hi, idesc, 1u);
            tc::umma_commit(empty + (seq + ks) % nslot);  // the segment is consumed
          }
        }
        if (leader) tc::umma_commit(bfull + blk);
        blk_ph ^= 1u << blk;
        seq += KS;
        __syncwarp();
      }
      pc0 += static_cast<uint32_t>(r.n + 2);
      g += r.n;
    }
  } else if (warp < 6) {
    // =============================== split warps: x_lo = x - trunc(x) ===============================
    const int tid = threadIdx.x - 64;
    uint32_t seq = 0;
    for (long long g = g_begin; g < g_end;) {
      const T3Run r = t3_run(g, g_end, a.To, a.tiles_q);
      const long long avail = in_plane - r.qt * kT3TileM;
      const int npos = static_cast<int>(avail < a.NP ? avail : a.NP);
      for (int i = 0; i < r.n + 2; ++i) {
        const int pa = r.t0 + i + a.plane_off;
        if (pa < a.zero_planes || pa >= a.Ti - a.zero_planes) continue;
#pragma unroll 1
        for (int ks = 0; ks < KS; ++ks) {
          const uint32_t slot = seq % nslot;
          tc::mbar_wait(full + slot, (seq / nslot) & 1u);
          ++seq;
          float4* hi = reinterpret_cast<float4*>(slot_s + slot * slot_bytes);
          float4* lo = reinterpret_cast<float4*>(slot_s + slot * slot_bytes + slot_term);
#pragma unroll
          for (int kg = 0; kg < 2; ++kg)
            for (int p = tid; p < npos; p += 128) {
              const float4 v = hi[kg * a.NP + p];
              lo[kg * a.NP + p] = make_float4(v.x - t3_trunc(v.x), v.y - t3_trunc(v.y), v.z - t3_trunc(v.z), v.w - t3_trunc(v.w));
            }
          tc::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async proxy
          __syncwarp();
          if (lane == 0) tc::mbar_arrive(ready + slot);
        }
      }
      g += r.n;
    }
  } else {
    // =============================== epilogue (warps 6..9) ===============================
    const int qd = warp & 3;  // the TMEM lane quadrant this warp may access
    const int row = qd * 32 + lane;
    const int GO = a.GO;
    const int Top = a.To + 2 * a.out_pad, Hop = a.Ho + 2 * a.out_pad, Wop = a.Wo + 2 * a.out_pad;
    const long long oplane = static_cast<long long>(Hop) * Wop;
    const long long mplane = static_cast<long long>(a.Ho) * a.Wo;
    const long long o_cg = static_cast<long long>(Top) * oplane, m_cg = static_cast<long long>(a.To) * mplane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
    uint32_t blk_ph = 0, pc0 = 0;
    float am = 0.f;  // largest magnitude this thread has written
    for (long long g = g_begin; g < g_end;) {
      const T3Run r = t3_run(g, g_end, a.To, a.tiles_q);
      const int q = r.qt * kT3TileM + row;
      const int ho = q / a.Wi, wo = q - ho * a.Wi;
      const bool valid = (ho < a.Ho) && (wo < a.Wo);
      long long o_off = (static_cast<long long>(r.b) * GO * Top + (r.t0 + a.out_pad)) * oplane +
                        static_cast<long long>(ho + a.out_pad) * Wop + (wo + a.out_pad);
      // unpadded offset of (b, channel 0 / group 0, t0, ho, wo): the NCDHW copy and the blocked mask source share it
      long long m_off = (static_cast<long long>(r.b) * GO * a.To + r.t0) * mplane + static_cast<long long>(ho) * a.Wo + wo;
      long long n_off = (static_cast<long long>(r.b) * a.Co * a.To + r.t0) * mplane + static_cast<long long>(ho) * a.Wo + wo;
      for (int j = 0; j < r.n; ++j, o_off += oplane, m_off += mplane, n_off += mplane) {
        uint4 mk[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        if (a.mask && valid) {
#pragma unroll
          for (int gi = 0; gi < 4; ++gi)
            if (half * 4 + gi < GO) mk[gi] = __ldg(a.mask + m_off + (half * 4 + gi) * m_cg);
        }
        // the three blocks of output j: run planes j, j+1, j+2 (time taps 0, 1, 2)
        bool have[3];
        uint32_t v[3][16];
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          const int pa = r.t0 + j + kt + a.plane_off;
          have[kt] = !(pa < a.zero_planes || pa >= a.Ti - a.zero_planes);
          if (have[kt]) {
            const uint32_t blk = (pc0 + static_cast<uint32_t>(j + kt)) % kT3Blocks;
            tc::mbar_wait(bfull + blk, (blk_ph >> blk) & 1u);
          }
        }
        tc::tc_fence_after();
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          if (have[kt]) {
            const uint32_t blk = (pc0 + static_cast<uint32_t>(j + kt)) % kT3Blocks;
            tc::tmem_ld_32x16(lane_addr + blk * kT3N + static_cast<uint32_t>(kt * kT3Half), v[kt]);
          } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) v[kt][c] = 0u;
          }
        }
        tc::tmem_ld_wait();
        // every TMEM read of this warp for output j is done.  Run plane j has no later reader (outputs j-1, j-2 came
        // before): release its block; the last output of the run also releases the two trailing planes.
        tc::tc_fence_before();
        __syncwarp();
        {
          const int last = (j == r.n - 1) ? 2 : 0;
          for (int d = 0; d <= last; ++d) {
            const int pa = r.t0 + j + d + a.plane_off;
            if (pa < a.zero_planes || pa >= a.Ti - a.zero_planes) continue;
            const uint32_t blk = (pc0 + static_cast<uint32_t>(j + d)) % kT3Blocks;
            if (lane == 0) tc::mbar_arrive(bempty + blk);
            blk_ph ^= 1u << blk;
          }
        }
        if (valid) {
          float f[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            float s = __uint_as_float(v[0][c]) + __uint_as_float(v[1][c]);
            s += __uint_as_float(v[2][c]);
            s += bias_s[c];
            if (a.relu) s = fmaxf(s, 0.f);
            f[c] = s;
          }
          if (a.mask) {
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) {
              const uint32_t mw[4] = {mk[gi].x, mk[gi].y, mk[gi].z, mk[gi].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) f[gi * 4 + e] = (__uint_as_float(mw[e]) > 0.f) ? f[gi * 4 + e] : 0.f;
            }
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) am = fmaxf(am, fabsf(f[c]));
          if (a.y_blk) {
#pragma unroll
            for (int gi = 0; gi < 4; ++gi)
              if (half * 4 + gi < GO)
                a.y_blk[o_off + (half * 4 + gi) * o_cg] = make_uint4(__float_as_uint(f[gi * 4]), __float_as_uint(f[gi * 4 + 1]),
                                                                     __float_as_uint(f[gi * 4 + 2]), __float_as_uint(f[gi * 4 + 3]));
          }
          if (a.y_nc) {
#pragma unroll
            for (int c = 0; c < 16; ++c)
              if (half * kT3Half + c < a.Co) a.y_nc[n_off + (half * kT3Half + c) * m_cg] = f[c];
          }
        }
      }
      pc0 += static_cast<uint32_t>(r.n + 2);
      g += r.n;
    }
    if (a.amax_out) t3_publish_amax(a.amax_out, am);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

// ---- CTA-pair version (cta_group::2), Cout = 32 ----------------------------------------------------------------------
// Two CTAs of a cluster (two SMs of a TPC) run ONE tcgen05.mma over M = 256 positions: each CTA stages the input planes of
// its own position tile and holds HALF of the N = 96 (kt, co) weight columns (the 110 KB that fit beside the plane ring);
// the tensor cores of both SMs read both halves.  Per plane the pair issues the same 27 x KS x 3 MMAs as one CTA of the
// single-CTA kernel, but for 256 positions x 32 output channels instead of 128 x 16: twice the work per issued MMA, and
// the MMA issue rate (one per ~55 clk) is what bounds these kernels.  Only the leader CTA (rank 0) issues MMAs; barriers
// the leader waits on (segment ready, accumulator block free) collect arrivals of both CTAs through shared::cluster
// addresses, barriers both CTAs wait on (segment consumed, block complete) are signalled by multicast commits.
constexpr int kT3PairBlocks = 5;   // accumulator blocks of 96 columns (480 of 512 TMEM columns)
constexpr int kT3PairN = 96;

struct T3PairRun {
  int col, t0, n;
};
__device__ __forceinline__ T3PairRun t3_pair_run(long long g, long long g_end, int To) {
  T3PairRun r;
  const long long pc = g / To;
  r.t0 = static_cast<int>(g - pc * To);
  r.col = static_cast<int>(pc);
  const long long left = g_end - g;
  r.n = static_cast<int>(left < (To - r.t0) ? left : (To - r.t0));
  return r;
}

// F16 = false: 3xTF32, KS = 8-channel steps (two blocked groups).  F16 = true: two-way fp16 split, KS = 16-channel steps (four
// blocked groups): a ring slot is [raw fp32: 4 groups][piece 0: 2 groups of 8 channels][piece 1], the split warps convert the
// whole segment, and the three products of a (tap, step) are x0.w1, x1.w0 (corrections, first) and x0.w0.
template <int KS, bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kT3Threads, 1) conv3d_igemm_tf32x3_pair_kernel(const T3Args a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [12] segment landed (local TMA)
  uint64_t* ready = full + kT3MaxSlots;                 // [12] LEADER's: x_lo written in both CTAs (8 arrivals)
  uint64_t* empty = ready + kT3MaxSlots;                // [12] segment consumed (multicast commit)
  uint64_t* wfull = empty + kT3MaxSlots;                // [1]  weights landed (local)
  uint64_t* wready = wfull + 1;                         // [1]  LEADER's: weights landed in both CTAs (2 arrivals)
  uint64_t* bfull = wready + 1;                         // [5]  accumulator block complete (multicast commit)
  uint64_t* bempty = bfull + kT3PairBlocks;             // [5]  LEADER's: block read out in both CTAs (8 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bempty + kT3PairBlocks);
  float* bias_s = reinterpret_cast<float*>(smem + 512);  // [32]
  uint8_t* w_s = smem + 640;
  constexpr uint32_t w_term_bytes = 9u * KS * 2u * kT3N * 16u;  // one of hi / lo, this CTA's 48 columns
  constexpr uint32_t w_bytes = 2u * w_term_bytes;
  const uint32_t slot_term = 2u * static_cast<uint32_t>(a.NP) * 16u;
  const uint32_t slot_bytes = (F16 ? 4u : 2u) * slot_term;
  constexpr int kRawGroups = F16 ? 4 : 2;  // blocked groups of 4 channels per segment
  uint8_t* slot_s = w_s + ((w_bytes + 127u) & ~127u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();
  const long long cta = blockIdx.x >> 1;
  const long long ncta = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kT3MaxSlots; ++i) { tc::mbar_init(full + i, 1); tc::mbar_init(ready + i, 8); tc::mbar_init(empty + i, 1); }
    tc::mbar_init(wfull, 1);
    tc::mbar_init(wready, 2);
    for (int i = 0; i < kT3PairBlocks; ++i) { tc::mbar_init(bfull + i, 1); tc::mbar_init(bempty + i, 8); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc_pair(tmem_ptr, 512);
  if (threadIdx.x >= 64 && threadIdx.x < 96) bias_s[lane] = (a.bias && lane < a.Co) ? __ldg(a.bias + lane) : 0.f;
  tc::tc_fence_before();
  tc::cluster_sync();  // barriers of both CTAs initialised before any remote arrival / multicast commit
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // pair tiles: g = pc * To + t over column PAIRS pc; the leader takes column 2 pc, the peer column 2 pc + 1
  const int ncol = a.B * a.tiles_q;
  const long long npc = (ncol + 1) / 2;
  const long long tiles = npc * a.To;
  const long long g_begin = tiles * cta / ncta;
  const long long g_end = tiles * (cta + 1) / ncta;
  const long long in_plane = static_cast<long long>(a.Hi) * a.Wi;
  const uint32_t nslot = static_cast<uint32_t>(a.nslot);

  if (warp == 0) {
    // =============================== producer (each CTA its own planes) ===============================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(wfull, w_bytes);
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wq) + static_cast<size_t>(rank) * w_bytes;
      for (uint32_t off = 0; off < w_bytes; off += 32768u) {
        const uint32_t n = (w_bytes - off < 32768u) ? (w_bytes - off) : 32768u;
        tc::bulk_g2s(w_s + off, wsrc + off, n, wfull);
      }
      uint32_t seq = 0;
      bool told = false;
      for (long long g = g_begin; g < g_end;) {
        const T3PairRun r = t3_pair_run(g, g_end, a.To);
        int col = 2 * r.col + static_cast<int>(rank);
        if (col >= ncol) col = ncol - 1;  // odd column count: the peer repeats the last column (its stores are skipped)
        const int b = col / a.tiles_q, qt = col - b * a.tiles_q;
        const int q0 = qt * kT3TileM;
        const long long avail = in_plane - q0;
        const uint32_t npos = static_cast<uint32_t>(avail < a.NP ? avail : a.NP);
        for (int i = 0; i < r.n + 2; ++i) {
          const int pa = r.t0 + i + a.plane_off;
          if (pa < a.zero_planes || pa >= a.Ti - a.zero_planes) continue;
#pragma unroll 1
          for (int ks = 0; ks < KS; ++ks) {
            const uint32_t slot = seq % nslot;
            tc::mbar_wait(empty + slot, ((seq / nslot) & 1u) ^ 1u);
            ++seq;
            tc::mbar_arrive_expect_tx(full + slot, npos * 16u * kRawGroups);
#pragma unroll
            for (int kg = 0; kg < kRawGroups; ++kg) {
              const uint4* src = a.x + ((static_cast<long long>(b) * a.G + (ks * kRawGroups + kg)) * a.Ti + pa) * in_plane + q0;
              tc::bulk_g2s(slot_s + slot * slot_bytes + static_cast<uint32_t>(kg) * a.NP * 16u, src, npos * 16u, full + slot);
            }
          }
          if (!told) {  // the first planes are in flight: now tell the leader that this CTA's weights have landed
            tc::mbar_wait(wfull, 0);
            tc::mbar_arrive_cluster(tc::map_to_cta(wready, 0));
            told = true;
          }
        }
        g += r.n;
      }
      if (!told) {
        tc::mbar_wait(wfull, 0);
        tc::mbar_arrive_cluster(tc::map_to_cta(wready, 0));
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA only) ===============================
    if (rank == 0) {
      const bool leader = tc::elect_one();
      const uint32_t a_lbo = static_cast<uint32_t>(a.NP) * 16u;
      const uint32_t b_lbo = static_cast<uint32_t>(kT3N) * 16u;  // 48 rows of B per CTA
      const uint32_t desc_hi = (128u >> 4) | (1u << 14);
      const uint32_t a_lo_base = ((a_lbo >> 4) << 16);
      const uint32_t b_lo_base = ((b_lbo >> 4) << 16) | ((tc::smem_u32(w_s) >> 4) & 0x3fffu);
      const uint32_t slot_addr16 = tc::smem_u32(slot_s) >> 4;
      const uint32_t slot_16 = slot_bytes >> 4, slot_term16 = slot_term >> 4;
      constexpr uint32_t w_term16 = w_term_bytes >> 4;
      constexpr uint32_t b_step16 = 2u * kT3N;
      uint32_t tap16[9];
#pragma unroll
      for (int hw = 0; hw < 9; ++hw) tap16[hw] = static_cast<uint32_t>((hw / 3) * a.Wi + (hw % 3));
      const uint32_t idesc = tc::umma_idesc(256, kT3PairN, /*F16 : TF32*/ F16 ? 0 : 2, /*K-major*/ 0, 0);
      // operand pieces inside a slot: 3xTF32 reads the raw segment as x_hi (the hardware truncates) and x_lo behind it;
      // the fp16 split reads piece 0 / piece 1 behind the raw segment
      const uint32_t a_main16 = F16 ? 2u * slot_term16 : 0u, a_corr16 = F16 ? 3u * slot_term16 : slot_term16;
      tc::mbar_wait_cluster(wready, 0);
      uint32_t seq = 0, blk_ph = 0, pc0 = 0;
      for (long long g = g_begin; g < g_end;) {
        const T3PairRun r = t3_pair_run(g, g_end, a.To);
        for (int i = 0; i < r.n + 2; ++i) {
          const int pa = r.t0 + i + a.plane_off;
          if (pa < a.zero_planes || pa >= a.Ti - a.zero_planes) continue;
          const uint32_t blk = (pc0 + static_cast<uint32_t>(i)) % kT3PairBlocks;
          tc::mbar_wait_cluster(bempty + blk, ((blk_ph >> blk) & 1u) ^ 1u);
          tc::tc_fence_after();
          const uint32_t d = tmem_base + blk * kT3PairN;
          uint32_t a_seg[KS];
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            const uint32_t s = seq + ks;
            const uint32_t slot = s % nslot;
            tc::mbar_wait_cluster(ready + slot, (s / nslot) & 1u);
            tc::tc_fence_after();
            a_seg[ks] = a_lo_base | ((slot_addr16 + slot * slot_16) & 0x3fffu);
            if (leader) {
#pragma unroll
              for (int hw = 0; hw < 9; ++hw) {
                const uint32_t b_t = b_lo_base + static_cast<uint32_t>(hw * KS + ks) * b_step16;
                if (F16) {
                  t3_mma_f16_pair(d, a_seg[ks] + a_main16 + tap16[hw], desc_hi, b_t + w_term16, desc_hi, idesc, (ks | hw) ? 1u : 0u);
                  t3_mma_f16_pair(d, a_seg[ks] + a_corr16 + tap16[hw], desc_hi, b_t, desc_hi, idesc, 1u);
                } else {
                  tc::umma_tf32_pair_lohi(d, a_seg[ks] + tap16[hw], desc_hi, b_t + w_term16, desc_hi, idesc, (ks | hw) ? 1u : 0u);
                  tc::umma_tf32_pair_lohi(d, a_seg[ks] + slot_term16 + tap16[hw], desc_hi, b_t, desc_hi, idesc, 1u);
                }
              }
            }
          }
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            if (leader) {
#pragma unroll
              for (int hw = 0; hw < 9; ++hw) {
                if (F16)
                  t3_mma_f16_pair(d, a_seg[ks] + a_main16 + tap16[hw], desc_hi, b_lo_base + static_cast<uint32_t>(hw * KS + ks) * b_step16,
                                  desc_hi, idesc, 1u);
                else
                  tc::umma_tf32_pair_lohi(d, a_seg[ks] + tap16[hw], desc_hi, b_lo_base + static_cast<uint32_t>(hw * KS + ks) * b_step16,
                                          desc_hi, idesc, 1u);
              }
              tc::umma_commit_pair(empty + (seq + ks) % nslot, 3);  // the segment is consumed in both CTAs
            }
          }
          if (leader) tc::umma_commit_pair(bfull + blk, 3);
          blk_ph ^= 1u << blk;
          seq += KS;
          __syncwarp();
        }
        pc0 += static_cast<uint32_t>(r.n + 2);
        g += r.n;
      }
    }
  } else if (warp < 6) {
    // =============================== split warps: x_lo = x - trunc(x) ===============================
    const int tid = threadIdx.x - 64;
    uint32_t seq = 0;
    const uint32_t ready0 = tc::map_to_cta(ready, 0);  // the leader's barrier array
    const float sx = F16 ? t3_exp2i(t3_scale_exp(__ldg(a.amax_in))) : 1.f;
    for (long long g = g_begin; g < g_end;) {
      const T3PairRun r = t3_pair_run(g, g_end, a.To);
      int col = 2 * r.col + static_cast<int>(rank);
      if (col >= ncol) col = ncol - 1;
      const int qt = col % a.tiles_q;
      const long long avail = in_plane - qt * kT3TileM;
      const int npos = static_cast<int>(avail < a.NP ? avail : a.NP);
      for (int i = 0; i < r.n + 2; ++i) {
        const int pa = r.t0 + i + a.plane_off;
        if (pa < a.zero_planes || pa >= a.Ti - a.zero_planes) continue;
#pragma unroll 1
        for (int ks = 0; ks < KS; ++ks) {
          const uint32_t slot = seq % nslot;
          tc::mbar_wait(full + slot, (seq / nslot) & 1u);
          ++seq;
          float4* hi = reinterpret_cast<float4*>(slot_s + slot * slot_bytes);
          if (F16) {
            uint4* p0 = reinterpret_cast<uint4*>(slot_s + slot * slot_bytes + 2u * slot_term);
            uint4* p1 = reinterpret_cast<uint4*>(slot_s + slot * slot_bytes + 3u * slot_term);
#pragma unroll
            for (int g8 = 0; g8 < 2; ++g8)
              for (int p = tid; p < npos; p += 128) {
                uint4 q0, q1;
                t3_split8h(hi[(2 * g8) * a.NP + p], hi[(2 * g8 + 1) * a.NP + p], sx, q0, q1);
                p0[g8 * a.NP + p] = q0;
                p1[g8 * a.NP + p] = q1;
              }
          } else {
            float4* lo = reinterpret_cast<float4*>(slot_s + slot * slot_bytes + slot_term);
#pragma unroll
            for (int kg = 0; kg < 2; ++kg)
              for (int p = tid; p < npos; p += 128) {
                const float4 v = hi[kg * a.NP + p];
                lo[kg * a.NP + p] = make_float4(v.x - t3_trunc(v.x), v.y - t3_trunc(v.y), v.z - t3_trunc(v.z), v.w - t3_trunc(v.w));
              }
          }
          tc::fence_proxy_async();
          __syncwarp();
          if (lane == 0) tc::mbar_arrive_cluster(ready0 + slot * 8u);
        }
      }
      g += r.n;
    }
  } else {
    // =============================== epilogue (warps 6..9), all 32 output channels of this CTA's positions ========
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const int GO = a.GO;
    const int Top = a.To + 2 * a.out_pad, Hop = a.Ho + 2 * a.out_pad, Wop = a.Wo + 2 * a.out_pad;
    const long long oplane = static_cast<long long>(Hop) * Wop;
    const long long mplane = static_cast<long long>(a.Ho) * a.Wo;
    const long long o_cg = static_cast<long long>(Top) * oplane, m_cg = static_cast<long long>(a.To) * mplane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16);
    const uint32_t bempty0 = tc::map_to_cta(bempty, 0);
    // two-way fp16 split: the sums carry the scales of both operands (two exact multiplications: 2^-(ex + ew) may not be a float)
    const float unscale_x = F16 ? t3_exp2i(-t3_scale_exp(__ldg(a.amax_in))) : 1.f;
    const float unscale_w = F16 ? t3_exp2i(-t3_scale_exp(__ldg(a.amax_w))) : 1.f;
    uint32_t blk_ph = 0, pc0 = 0;
    float am = 0.f;  // largest magnitude this thread has written
    for (long long g = g_begin; g < g_end;) {
      const T3PairRun r = t3_pair_run(g, g_end, a.To);
      int col = 2 * r.col + static_cast<int>(rank);
      const bool dup = col >= ncol;
      if (dup) col = ncol - 1;
      const int b = col / a.tiles_q, qt = col - b * a.tiles_q;
      const int q = qt * kT3TileM + row;
      const int ho = q / a.Wi, wo = q - ho * a.Wi;
      const bool valid = (ho < a.Ho) && (wo < a.Wo) && !dup;
      long long o_off = (static_cast<long long>(b) * GO * Top + (r.t0 + a.out_pad)) * oplane + static_cast<long long>(ho + a.out_pad) * Wop +
                        (wo + a.out_pad);
      long long m_off = (static_cast<long long>(b) * GO * a.To + r.t0) * mplane + static_cast<long long>(ho) * a.Wo + wo;
      long long n_off = (static_cast<long long>(b) * a.Co * a.To + r.t0) * mplane + static_cast<long long>(ho) * a.Wo + wo;
      for (int j = 0; j < r.n; ++j, o_off += oplane, m_off += mplane, n_off += mplane) {
        float f[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) f[c] = 0.f;
        // data gradient: the ReLU-mask source of this position is fetched BEFORE the accumulator blocks are waited for (the
        // loads sat behind the tensor-memory reads: 47 % of the stall samples, tensor pipe at 53 %)
        uint4 mk[8];
        if (a.mask && valid) {
#pragma unroll
          for (int gi = 0; gi < 8; ++gi)
            if (gi < GO) mk[gi] = __ldg(a.mask + m_off + gi * m_cg);
        }
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          const int pa = r.t0 + j + kt + a.plane_off;
          if (pa < a.zero_planes || pa >= a.Ti - a.zero_planes) continue;  // warp-uniform
          const uint32_t blk = (pc0 + static_cast<uint32_t>(j + kt)) % kT3PairBlocks;
          tc::mbar_wait(bfull + blk, (blk_ph >> blk) & 1u);
          tc::tc_fence_after();
          uint32_t v[32];
          tc::tmem_ld_32x32(lane_addr + blk * kT3PairN + static_cast<uint32_t>(kt * 32), v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) f[c] += __uint_as_float(v[c]);
        }
        tc::tc_fence_before();
        __syncwarp();
        {
          const int last = (j == r.n - 1) ? 2 : 0;
          for (int d = 0; d <= last; ++d) {
            const int pa = r.t0 + j + d + a.plane_off;
            if (pa < a.zero_planes || pa >= a.Ti - a.zero_planes) continue;
            const uint32_t blk = (pc0 + static_cast<uint32_t>(j + d)) % kT3PairBlocks;
            if (lane == 0) tc::mbar_arrive_cluster(bempty0 + blk * 8u);
            blk_ph ^= 1u << blk;
          }
        }
        if (valid) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float s = F16 ? fmaf(f[c] * unscale_x, unscale_w, bias_s[c]) : f[c] + bias_s[c];
            if (a.relu) s = fmaxf(s, 0.f);
            f[c] = s;
          }
          if (a.mask) {
#pragma unroll
            for (int gi = 0; gi < 8; ++gi) {
              if (gi >= GO) continue;
              const uint32_t mw[4] = {mk[gi].x, mk[gi].y, mk[gi].z, mk[gi].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) f[gi * 4 + e] = (__uint_as_float(mw[e]) > 0.f) ? f[gi * 4 + e] : 0.f;
            }
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) am = fmaxf(am, fabsf(f[c]));
          if (a.y_blk) {
#pragma unroll
            for (int gi = 0; gi < 8; ++gi)
              if (gi < GO)
                a.y_blk[o_off + gi * o_cg] = make_uint4(__float_as_uint(f[gi * 4]), __float_as_uint(f[gi * 4 + 1]),
                                                        __float_as_uint(f[gi * 4 + 2]), __float_as_uint(f[gi * 4 + 3]));
          }
          if (a.y_nc) {
#pragma unroll
            for (int c = 0; c < 32; ++c)
              if (c < a.Co) a.y_nc[n_off + c * m_cg] = f[c];
          }
        }
      }
      pc0 += static_cast<uint32_t>(r.n + 2);
      g += r.n;
    }
    if (a.amax_out) t3_publish_amax(a.amax_out, am);
  }
  tc::tc_fence_before();
  tc::cluster_sync();  // both CTAs are done with tensor memory and with each other's barriers
  if (warp == 1) tc::tmem_dealloc_pair(tmem_base, 512);
}

// ---- layout kernels ------------------------------------------------------------------------------------------
// [B][C][T][H][W] fp32 -> blocked fp32 [B][G][T+2p][H+2p][W+2p][4] interior (channels >= C are zero)
// grid: x = chunks of 256 positions of a plane, y = (sample, channel group); a thread walks the time planes of its position:
// one 32-bit division per thread (the grid-stride form over the flat index paid three 64-bit divisions per element), one
// atomicMax per warp
__global__ void __launch_bounds__(256) nc_to_blocked_f32_kernel(const float* __restrict__ x, float4* __restrict__ y, int C, int G, int T,
                                                                int H, int W, int pad, unsigned* __restrict__ amax_out) {
  float am = 0.f;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const int g = blockIdx.y % G, b = blockIdx.y / G;
  if (p < H * W) {
    const long long hw = static_cast<long long>(H) * W;
    const int h = p / W, w = p - h * W;
    const int Hp = H + 2 * pad, Wp = W + 2 * pad, Tp = T + 2 * pad;
    const float* xs = x + (static_cast<long long>(b) * C + g * 4) * T * hw + p;
    float4* ys = y + ((static_cast<long long>(blockIdx.y) * Tp + pad) * Hp + (h + pad)) * Wp + (w + pad);
#pragma unroll 2
    for (int t = 0; t < T; ++t) {
      float f[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) f[j] = (g * 4 + j < C) ? xs[(static_cast<long long>(j) * T + t) * hw] : 0.f;
      ys[static_cast<long long>(t) * Hp * Wp] = make_float4(f[0], f[1], f[2], f[3]);
      am = fmaxf(fmaxf(am, fmaxf(fabsf(f[0]), fabsf(f[1]))), fmaxf(fabsf(f[2]), fabsf(f[3])));
    }
  }
  if (amax_out) t3_publish_amax_block(amax_out, am);  // (every thread of the grid reaches this line)
}

// blocked fp32 [B][G][T][H][W][4] -> [B][C][T][H][W] fp32
__global__ void blocked_f32_to_nc_kernel(const float4* __restrict__ x, float* __restrict__ y, int C, int G, long long thw,
                                         long long total) {
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pos = idx % thw;
    const long long r = idx / thw;
    const int g = static_cast<int>(r % G);
    const long long b = r / G;
    const float4 v = x[idx];
    const float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = g * 4 + j;
      if (c < C) y[(b * C + c) * thw + pos] = f[j];
    }
  }
}

// int16 [B][C][T][H][W] -> normalised blocked fp32 [B][G][T][H][W][4] (a1 fused with the layout change; the arithmetic
// is sat_norm of common.cuh: bit-identical to the reference)
// grid: x = chunks of 1024 positions, y = (sample, channel group): no 64-bit division per element (the first version spent
// 0.38 ms on them for 180 MB of traffic)
__global__ void __launch_bounds__(256) sat_normalise_blocked_f32_kernel(const int16_t* __restrict__ x, float4* __restrict__ y,
                                                                        const float* __restrict__ mean, const float* __restrict__ stdv,
                                                                        int C, int G, int thw, unsigned* __restrict__ amax_out) {
  float am = 0.f;
  const int g = blockIdx.y % G, b = blockIdx.y / G;
  const int16_t* xs = x + (static_cast<long long>(b) * C + g * 4) * thw;
  float4* ys = y + static_cast<long long>(blockIdx.y) * thw;
  float mu[4], sd[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool ok = g * 4 + j < C;
    mu[j] = ok ? __ldg(mean + g * 4 + j) : 0.f;
    sd[j] = ok ? __ldg(stdv + g * 4 + j) : 1.f;
  }
  const int p0 = blockIdx.x * 1024 + threadIdx.x;
  int16_t v[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pos = p0 + i * 256;
#pragma unroll
    for (int j = 0; j < 4; ++j) v[i][j] = (pos < thw && g * 4 + j < C) ? xs[static_cast<long long>(j) * thw + pos] : int16_t(0);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pos = p0 + i * 256;
    if (pos >= thw) break;
    float f[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) f[j] = (g * 4 + j < C) ? sat_norm(v[i][j], mu[j], sd[j]) : 0.f;
    ys[pos] = make_float4(f[0], f[1], f[2], f[3]);
    am = fmaxf(fmaxf(am, fmaxf(fabsf(f[0]), fabsf(f[1]))), fmaxf(fabsf(f[2]), fabsf(f[3])));
  }
  if (amax_out) t3_publish_amax_block(amax_out, am);
}

int g_t3_pair = 1;  // CTA-pair kernel for Cout > 16 (tools may switch it off through pvb200_debug_set_tf32x3_pair)

static int t3_groups(int C) { return 2 * ceil_div(C, 8); }  // channel groups of 4, padded to an even count (UMMA K = 8)
static int t3_halves(int Co) { return Co <= kT3Half ? 1 : 2; }

static size_t t3_ws_bytes(int Ci_role, int Co_role) {
  return static_cast<size_t>(t3_halves(Co_role)) * 2 * 9 * (t3_groups(Ci_role) / 2) * 2 * kT3N * 16;
}

static int launch_t3(const void* xb, const float* w, long long s_co, long long s_ci, int flip, const float* bias, const void* mask,
                     void* y_blk, float* y_nc, void* ws, size_t ws_bytes, int B, int Ci, int Ti, int Hi, int Wi, int Co, int out_pad,
                     int relu, int zero_planes, int To, int plane_off, const float* amax_in, float* amax_out, cudaStream_t stream) {
  T3Args a;
  a.x = static_cast<const uint4*>(xb);
  a.bias = bias;
  a.mask = static_cast<const uint4*>(mask);
  a.y_blk = static_cast<uint4*>(y_blk);
  a.y_nc = y_nc;
  a.amax_out = reinterpret_cast<unsigned*>(amax_out);
  a.B = B; a.G = t3_groups(Ci); a.Ti = Ti; a.Hi = Hi; a.Wi = Wi;
  PVB_REQUIRE(Co <= 32, "conv3d_tf32x3: Cout=%d > 32 is not supported by the tensor-core path", Co);
  PVB_REQUIRE(Ci <= 32, "conv3d_tf32x3: Cin=%d > 32 is not supported by the tensor-core path", Ci);
  a.Co = Co; a.GO = t3_groups(Co); a.nhalf = t3_halves(Co);
  a.To = To; a.Ho = Hi - 2; a.Wo = Wi - 2;
  a.plane_off = plane_off;
  PVB_REQUIRE(a.To > 0 && a.Ho > 0 && a.Wo > 0, "conv3d_tf32x3: input %dx%dx%d too small", Ti, Hi, Wi);
  PVB_REQUIRE(y_blk || y_nc, "conv3d_tf32x3: no output requested");
  a.out_pad = out_pad; a.relu = relu;
  a.zero_planes = zero_planes;
  a.NP = round_up(kT3TileM + 2 * Wi + 2, 8);
  const int Qtot = (a.Ho - 1) * Wi + a.Wo;
  a.tiles_q = ceil_div(Qtot, kT3TileM);
  const int sms = sm_count();
  PVB_REQUIRE(sms > 0, "conv3d_tf32x3: no CUDA device");
  a.tiles = static_cast<long long>(B) * a.tiles_q * a.To;
  const size_t need = t3_ws_bytes(Ci, Co);
  if (!ws || ws_bytes < need) {
    set_error("conv3d_tf32x3: workspace too small (%zu < %zu bytes)", ws_bytes, need);
    return PVB200_ERR_WORKSPACE;
  }
  PVB_REQUIRE(reinterpret_cast<uintptr_t>(xb) % 16 == 0 && reinterpret_cast<uintptr_t>(y_blk) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(mask) % 16 == 0 && reinterpret_cast<uintptr_t>(ws) % 16 == 0,
              "conv3d_tf32x3: pointers must be 16-byte aligned");
  a.wq = static_cast<const uint4*>(ws);
  const int KS = a.G / 2;
  const bool pair = g_t3_pair && a.nhalf == 2 && sms >= 2;
  a.amax_in = amax_in;
  a.amax_w = nullptr;
  if (amax_in && pair && a.G % 4 == 0) {
    // two-way fp16 split (CTA pair, 16-channel steps): half the weights, half the MMAs
    const int KS16 = a.G / 4;
    const size_t w_bytes = static_cast<size_t>(2) * 9 * KS16 * 2 * kT3N * 16;  // per CTA: two pieces
    float* amax_w = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + 2 * w_bytes);  // behind the weights of both ranks (<= need - 16)
    a.amax_w = amax_w;
    t3_weight_prep_f16_kernel<<<ceil_div(static_cast<int>(2 * w_bytes / 4), 256), 256, 0, stream>>>(
        w, static_cast<uint32_t*>(ws), Ci, Co, KS16, s_co, s_ci, flip, Co * Ci * 27, amax_w);
    PVB_LAUNCHED("t3_weight_prep_f16");
    const size_t fixed = 640 + round_up(w_bytes, static_cast<size_t>(128));
    const size_t slot_bytes = static_cast<size_t>(8) * a.NP * 16;
    long long nslot = (227 * 1024 - static_cast<long long>(fixed)) / static_cast<long long>(slot_bytes);
    if (nslot > kT3MaxSlots) nslot = kT3MaxSlots;
    PVB_REQUIRE(nslot >= KS16, "conv3d_f16x2: Cin=%d Cout=%d width=%d does not fit in shared memory", Ci, Co, Wi);
    a.nslot = static_cast<int>(nslot);
    const size_t smem = fixed + nslot * slot_bytes;
    const long long pair_tiles = ((static_cast<long long>(B) * a.tiles_q + 1) / 2) * a.To;
    long long npairs = sms / 2;
    if (npairs > pair_tiles) npairs = pair_tiles;
    if (KS16 == 1) {
      PVB_CUDA(cudaFuncSetAttribute(conv3d_igemm_tf32x3_pair_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      conv3d_igemm_tf32x3_pair_kernel<1, true><<<static_cast<unsigned>(2 * npairs), kT3Threads, smem, stream>>>(a);
    } else {
      PVB_CUDA(cudaFuncSetAttribute(conv3d_igemm_tf32x3_pair_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      conv3d_igemm_tf32x3_pair_kernel<2, true><<<static_cast<unsigned>(2 * npairs), kT3Threads, smem, stream>>>(a);
    }
    PVB_LAUNCHED("conv3d_igemm_f16x2_pair");
    return PVB200_OK;
  }
  {
    const int total = static_cast<int>(need / 4);
    t3_weight_prep_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(w, static_cast<float*>(ws), Ci, Co, KS, a.nhalf, s_co, s_ci, flip,
                                                                    pair ? 1 : 0);
    PVB_LAUNCHED("t3_weight_prep");
  }
  const size_t w_bytes = static_cast<size_t>(2) * 9 * KS * 2 * kT3N * 16;
  const size_t fixed = 640 + round_up(w_bytes, static_cast<size_t>(128));
  const size_t slot_bytes = static_cast<size_t>(4) * a.NP * 16;
  long long nslot = (227 * 1024 - static_cast<long long>(fixed)) / static_cast<long long>(slot_bytes);
  if (nslot > kT3MaxSlots) nslot = kT3MaxSlots;
  PVB_REQUIRE(nslot >= KS, "conv3d_tf32x3: Cin=%d Cout=%d width=%d does not fit in shared memory", Ci, Co, Wi);
  a.nslot = static_cast<int>(nslot);
  const size_t smem = fixed + nslot * slot_bytes;
  if (pair) {
    const long long pair_tiles = ((static_cast<long long>(B) * a.tiles_q + 1) / 2) * a.To;
    long long npairs = sms / 2;
    if (npairs > pair_tiles) npairs = pair_tiles;
#define PVB_T3P_LAUNCH(KSV)                                                                                                      \
  do {                                                                                                                          \
    PVB_CUDA(cudaFuncSetAttribute(conv3d_igemm_tf32x3_pair_kernel<KSV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    conv3d_igemm_tf32x3_pair_kernel<KSV, false><<<static_cast<unsigned>(2 * npairs), kT3Threads, smem, stream>>>(a);                  \
  } while (0)
    switch (KS) {
      case 1: PVB_T3P_LAUNCH(1); break;
      case 2: PVB_T3P_LAUNCH(2); break;
      case 3: PVB_T3P_LAUNCH(3); break;
      default: PVB_T3P_LAUNCH(4); break;
    }
#undef PVB_T3P_LAUNCH
    PVB_LAUNCHED("conv3d_igemm_tf32x3_pair");
    return PVB200_OK;
  }
  long long per_half = sms / a.nhalf;
  if (per_half > a.tiles) per_half = a.tiles;
  if (per_half < 1) per_half = 1;
  const long long grid = per_half * a.nhalf;
#define PVB_T3_LAUNCH(KSV)                                                                                                  \
  do {                                                                                                                      \
    PVB_CUDA(cudaFuncSetAttribute(conv3d_igemm_tf32x3_kernel<KSV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    conv3d_igemm_tf32x3_kernel<KSV><<<static_cast<unsigned>(grid), kT3Threads, smem, stream>>>(a);                         \
  } while (0)
  switch (KS) {
    case 1: PVB_T3_LAUNCH(1); break;
    case 2: PVB_T3_LAUNCH(2); break;
    case 3: PVB_T3_LAUNCH(3); break;
    default: PVB_T3_LAUNCH(4); break;
  }
#undef PVB_T3_LAUNCH
  PVB_LAUNCHED("conv3d_igemm_tf32x3");
  return PVB200_OK;
}

}  // namespace pvb

extern "C" {

/* tools only (not declared in pvb200.h): 0 = single-CTA kernel for every layer, 1 = CTA-pair kernel where it applies */
void pvb200_debug_set_tf32x3_pair(int on) { pvb::g_t3_pair = on; }

int pvb200_blocked4_channel_groups(int C) { return pvb::t3_groups(C); }

size_t pvb200_conv3d_tf32x3_workspace_bytes(int Cin, int Cout) {
  const size_t f = pvb::t3_ws_bytes(Cin, Cout), d = pvb::t3_ws_bytes(Cout, Cin);
  return f > d ? f : d;
}

int pvb200_nc_to_blocked_f32(const float* x, float* y, int B, int C, int T, int H, int W, int pad, float* amax_out,
                             pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(x && y && B > 0 && C > 0 && T > 0 && H > 0 && W > 0 && pad >= 0, "nc_to_blocked_f32: bad argument");
  const int G = t3_groups(C);
  PVB_REQUIRE(static_cast<long long>(H) * W < (1LL << 31) - 256 && static_cast<long long>(B) * G <= 65535, "nc_to_blocked_f32: input too large");
  const dim3 grid(static_cast<unsigned>(ceil_div(H * W, 256)), static_cast<unsigned>(B * G));
  nc_to_blocked_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, reinterpret_cast<float4*>(y), C, G, T, H, W, pad,
                                                                reinterpret_cast<unsigned*>(amax_out));
  PVB_LAUNCHED("nc_to_blocked_f32");
  return PVB200_OK;
}

int pvb200_blocked_f32_to_nc(const float* x, float* y, int B, int C, int T, int H, int W, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(x && y && B > 0 && C > 0 && T > 0 && H > 0 && W > 0, "blocked_f32_to_nc: bad argument");
  const int G = t3_groups(C);
  const long long thw = static_cast<long long>(T) * H * W;
  const long long total = static_cast<long long>(B) * G * thw;
  long long grid = ceil_div(total, 256LL);
  if (grid > 148 * 32) grid = 148 * 32;
  blocked_f32_to_nc_kernel<<<static_cast<unsigned>(grid), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x), y, C, G, thw,
                                                                                        total);
  PVB_LAUNCHED("blocked_f32_to_nc");
  return PVB200_OK;
}

int pvb200_sat_normalise_blocked_f32(const int16_t* x, float* y, const float* mean, const float* stdv, int B, int C, int T, int H,
                                     int W, float* amax_out, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(x && y && mean && stdv && B > 0 && C > 0 && T > 0 && H > 0 && W > 0, "sat_normalise_blocked_f32: bad argument");
  const int G = t3_groups(C);
  const long long thw = static_cast<long long>(T) * H * W;
  PVB_REQUIRE(thw < (1LL << 31) - 1024 && static_cast<long long>(B) * G <= 65535, "sat_normalise_blocked_f32: input too large");
  const dim3 grid(static_cast<unsigned>(ceil_div(thw, 1024LL)), static_cast<unsigned>(B * G));
  sat_normalise_blocked_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, reinterpret_cast<float4*>(y), mean, stdv, C, G,
                                                                       static_cast<int>(thw), reinterpret_cast<unsigned*>(amax_out));
  PVB_LAUNCHED("sat_normalise_blocked_f32");
  return PVB200_OK;
}

int pvb200_conv3d_fwd_tf32x3(const float* xb, const float* w, const float* bias, float* y_blk, float* y_nc, void* workspace,
                             size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout, int relu, int out_pad,
                             int pad_t, const float* amax_in, float* amax_out, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(xb && w, "conv3d_fwd_tf32x3: null pointer");
  PVB_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && out_pad >= 0 && (pad_t == 0 || pad_t == 1), "conv3d_fwd_tf32x3: bad shape");
  return launch_t3(xb, w, static_cast<long long>(Cin) * 27, 27, 0, bias, nullptr, y_blk, y_nc, workspace, workspace_bytes, B, Cin, Ti,
                   Hi, Wi, Cout, out_pad, relu, /*zero_planes=*/0, /*To=*/Ti + 2 * pad_t - 2, /*plane_off=*/-pad_t, amax_in, amax_out, as_stream(stream));
}

int pvb200_conv3d_dgrad_tf32x3(const float* gz_padded, const float* w, const float* mask_blk, float* gx_blk, float* gx_nc,
                               void* workspace, size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout, int out_pad,
                               int pad_t, const float* amax_in, float* amax_out, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(gz_padded && w, "conv3d_dgrad_tf32x3: null pointer");
  PVB_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && Ti + 2 * pad_t > 2 && Hi > 2 && Wi > 2 && out_pad >= 0 && (pad_t == 0 || pad_t == 1),
              "conv3d_dgrad_tf32x3: bad shape");
  // kernel input = gz (Ti + 2 pad_t - 2 planes) zero-padded by 2: [B][G(Cout)][Ti + 2 pad_t + 2][Hi+2][Wi+2][4];
  // kernel output = gx [B][G(Cin)][Ti][Hi][Wi]: gx[t] reads the padded planes t + pad_t + {0,1,2}
  return launch_t3(gz_padded, w, /*s_co (out role = ci)*/ 27, /*s_ci (in role = co)*/ static_cast<long long>(Cin) * 27, 1, nullptr,
                   mask_blk, gx_blk, gx_nc, workspace, workspace_bytes, B, /*Ci role*/ Cout, Ti + 2 * pad_t + 2, Hi + 2, Wi + 2,
                   /*Co role*/ Cin, out_pad, 0, /*zero_planes=*/2, /*To=*/Ti, /*plane_off=*/pad_t, amax_in, amax_out, as_stream(stream));
}

}  // extern "C"
