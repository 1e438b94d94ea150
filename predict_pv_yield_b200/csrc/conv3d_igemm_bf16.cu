// conv3d_igemm_bf16.cu -- a3/a4/a11 in bf16: Conv3d 3x3x3 forward and data-gradient as an implicit GEMM on the
// 5th-generation tensor cores (tcgen05.mma, fp32 accumulators in TMEM), operands staged by the TMA engine.
//
// Reference call sites: predict_pv_yield/models/conv3d/model.py:80-90,117-120 (and their autograd).
//
// Data layout ("blocked", NC8DHW8c): activations are [B][Cg][T][H][W][8] bf16 -- channel groups of 8 (16 bytes)
// innermost.  With it the implicit GEMM needs NO im2col and no swizzle:
//   * GEMM M = output positions, K = Cin per tap, N = output channels.
//   * positions are flattened with the INPUT pitch: q = ho*Wi + wo, so the A operand of tap (kh,kw) for rows
//     q0..q0+127 is the contiguous run of 16-byte elements starting at q0 + kh*Wi + kw of an input plane: exactly the
//     SWIZZLE_NONE K-major canonical layout (8 rows x 16 B core matrices, SBO = 128 B, LBO = the stride between
//     channel-group planes).  A tap is just a different descriptor start address (measured with
//     tools/probe/mma_probe.cu: start addresses that are not 128-byte aligned cost nothing).
//   * Cout = 32 alone would make N = 32 MMAs, which one thread cannot issue faster than one per ~54 clk (16 clk of
//     math).  So the three TIME taps are merged into N = 96 and the loop runs over INPUT planes ("scatter" form):
//     for input plane p,  D[r, kt*32+co] += sum_{kh,kw,ci} x[p, r + kh*Wi + kw, ci] * w[co,ci,kt,kh,kw]  is the
//     contribution of plane p to the outputs t = p - kt.  The accumulators of consecutive output planes sit side by
//     side in TMEM in DESCENDING order (output j of a run in columns (15-j)*32 ...), so the three blocks of one MMA
//     are one contiguous 96-column window that slides down by 32 columns per input plane: the kt sum happens inside
//     the tensor core, an output plane is complete after three input planes, and the epilogue is a plain
//     TMEM -> bias/ReLU -> bf16 -> store with no cross-row traffic.  (An earlier version merged the kw taps instead;
//     its shift-add epilogue -- 3x the TMEM reads, 2 shuffles per output, a cross-warp exchange -- cost twice the
//     MMA time.)  The first and last two planes of a run use a narrower window (N = 32 / 64).
//   * the first MMA that touches an accumulator must overwrite it: the first MMA of a plane is split into the
//     blocks that are fresh (accumulate = 0) and the rest.
//   * one input plane segment (all channel groups, 128 + 2*Wi + 2 positions) is ONE bulk copy per channel group
//     (cp.async.bulk, completion on an mbarrier) into a ring; each plane is consumed by 18 MMAs and released.
//   * all 27 x Cin x Cout weights (55 KB bf16) stay resident in shared memory for the whole persistent CTA.
//   * wrap columns (wo >= Wo) are computed and dropped in the epilogue (2/Wi ~ 3 % waste).
// Warp roles (576 threads): warp 0 = copy producer, warp 1 = MMA issuer (one elected thread) + TMEM owner,
// warps 2-17 = epilogue: four groups of four warps (one per TMEM lane quadrant), accumulator slot s is read by group
// s % 4; up to 16 output planes of a run live in TMEM (512 columns), so the epilogue has a whole run of slack.
// The data gradient is the same kernel on a zero-padded gz (padding 2) with flipped / transposed weights; input
// planes that lie in the zero padding are skipped.
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pvb {

constexpr int kIgThreads = 576;  // producer warp, MMA warp, 16 epilogue warps
constexpr int kIgMaxSlots = 8;   // input-plane ring
constexpr int kIgTileM = 128;    // MMA rows = output positions per tile
constexpr int kIgAccSlots = 16;  // output-plane accumulators resident in TMEM = longest run segment

struct IgemmArgs {
  const uint4* x;     // [B][Cg][Ti][Hi][Wi] 16-byte elements (8 bf16 channels)
  const uint4* wq;    // [9 (kh,kw)][Cg][3 (kt)][CoP] 16-byte elements: 8 input channels of one (kt, output channel)
  const float* bias;  // [Co] or null
  const uint4* mask;  // [B][CogOut][To][Ho][Wo] or null
  uint4* y;           // [B][CogOut][To+2p][Ho+2p][Wo+2p]
  uint4* y2;          // optional second copy in the wgrad operand layout [B][CogOut][To][QP2], pitch Wo + 2 (or null)
  int QP2;
  int B, Cg, Ti, Hi, Wi;
  int CoP, Co, CogOut, To, Ho, Wo;  // CoP = Cout padded to 16/32; MMA N = 3*CoP (the three kt taps side by side)
  int out_pad, relu;
  int zero_planes;  // the first / last `zero_planes` input time planes are all zero (padded gz): their MMAs are skipped
  int plane_off;    // input plane of output t, tap kt: t + kt + plane_off (-pad_t forward, +pad_t data gradient); planes
                    // outside [zero_planes, Ti - zero_planes) are zero and skipped (this is also the layer's time padding)
  int NP;       // staged positions per (plane, channel group)
  int tiles_q;  // q tiles per output plane
  int nslot;    // ring slots, as many as fit in shared memory (<= 8)
  long long* dbg;  // optional [grid][8] cycle counters (profiling builds of the tools; null in production)
  int dbg_flags;   // profiling builds only: 1 = skip the plane copies, 2 = skip the output stores, 16 = no epilogue
  long long tiles;  // B * tiles_q * To
};

// weights fp32 [Co][Ci][27] -> bf16 [(kh,kw)][Cg][kt*CoP + co][8 ci]; flipped / transposed roles for the data gradient
__global__ void igemm_weight_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wq, int Ci_role,
                                         int Co_role, int Cg, int CoP, long long s_co, long long s_ci, int flip) {
  const int N = 3 * CoP;
  const int total = 9 * Cg * N * 8;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c8 = idx & 7;
    const int n = (idx >> 3) % N;
    const int cg = (idx / (8 * N)) % Cg;
    const int hw = idx / (8 * N * Cg);  // kh*3 + kw
    const int kt = n / CoP, co = n - kt * CoP;
    const int ci = cg * 8 + c8;
    const int tap = kt * 9 + hw;
    float v = 0.f;
    if (co < Co_role && ci < Ci_role) v = w[co * s_co + ci * s_ci + (flip ? 26 - tap : tap)];
    wq[idx] = __float2bfloat16_rn(v);
  }
}

// D[tmem] (+)= A * B with descriptors given as (lo, hi) halves: only `lo` changes between MMAs of a tile
__device__ __forceinline__ void igemm_mma(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t accumulate) {
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

// A "run" = consecutive output time steps of one (sample, q-tile) column handled by one CTA (at most 16: they all
// live in TMEM).  Output tiles are numbered g = (b * tiles_q + qt) * To + t and split evenly (tile-granular) over the
// persistent CTAs; every role walks the same sequence of runs.  A run of n outputs consumes n + 2 input planes.
struct IgRun {
  int b, qt, t0, ntiles;
};
__device__ __forceinline__ IgRun ig_run(long long g, long long g_end, int To, int tiles_q) {
  IgRun r;
  const long long col = g / To;
  r.t0 = static_cast<int>(g - col * To);
  r.qt = static_cast<int>(col % tiles_q);
  r.b = static_cast<int>(col / tiles_q);
  const long long left = g_end - g;
  int n = static_cast<int>(left < (To - r.t0) ? left : (To - r.t0));
  r.ntiles = n < kIgAccSlots ? n : kIgAccSlots;
  return r;
}

template <int CG, bool DBG>
__global__ void __launch_bounds__(kIgThreads, 1) conv3d_igemm_bf16_kernel(const IgemmArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);       // [8]  input plane landed
  uint64_t* empty = full + kIgMaxSlots;                     // [8]  input plane consumed
  uint64_t* wfull = empty + kIgMaxSlots;                    // [1]  weights landed
  uint64_t* tfull = wfull + 1;                              // [16] output accumulator complete
  uint64_t* tempty = tfull + kIgAccSlots;                   // [16] output accumulator read out
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + kIgAccSlots);
  float* bias_s = reinterpret_cast<float*>(smem + 512);     // [32]
  uint8_t* w_s = smem + 640;
  const int CoP = a.CoP;
  const int N = 3 * CoP;
  const uint32_t w_bytes = 9u * CG * N * 16u;
  const uint32_t slot_bytes = static_cast<uint32_t>(CG) * a.NP * 16u;
  uint8_t* slot_s = w_s + ((w_bytes + 127u) & ~127u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = static_cast<uint32_t>(kIgAccSlots * CoP);  // 512 or 256

  if (threadIdx.x == 0) {
    for (int i = 0; i < kIgMaxSlots; ++i) { tc::mbar_init(full + i, 1); tc::mbar_init(empty + i, 1); }
    tc::mbar_init(wfull, 1);
    for (int i = 0; i < kIgAccSlots; ++i) { tc::mbar_init(tfull + i, 1); tc::mbar_init(tempty + i, 4); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, tmem_cols);
  if (threadIdx.x >= 64 && threadIdx.x < 96) bias_s[lane] = (a.bias && lane < a.Co) ? __ldg(a.bias + lane) : 0.f;
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const long long g_begin = a.tiles * blockIdx.x / gridDim.x;
  const long long g_end = a.tiles * (blockIdx.x + 1) / gridDim.x;
  const long long in_plane = static_cast<long long>(a.Hi) * a.Wi;
  const uint32_t nslot = static_cast<uint32_t>(a.nslot);

  if (warp == 0) {
    // =============================== producer ===============================
    if (lane == 0) {
      tc::mbar_arrive_expect_tx(wfull, w_bytes);
      for (uint32_t off = 0; off < w_bytes; off += 32768u) {
        const uint32_t n = (w_bytes - off < 32768u) ? (w_bytes - off) : 32768u;
        tc::bulk_g2s(w_s + off, reinterpret_cast<const uint8_t*>(a.wq) + off, n, wfull);
      }
      uint32_t seq = 0;
      for (long long g = g_begin; g < g_end;) {
        const IgRun r = ig_run(g, g_end, a.To, a.tiles_q);
        const int q0 = r.qt * kIgTileM;
        const long long avail = in_plane - q0;
        const uint32_t npos = static_cast<uint32_t>(avail < a.NP ? avail : a.NP);
        for (int p = 0; p < r.ntiles + 2; ++p) {
          const int pa = r.t0 + p + a.plane_off;  // absolute input plane
          if (pa < a.zero_planes || pa >= a.Ti - a.zero_planes) continue;  // all-zero plane: never staged
          const uint32_t slot = seq % nslot;
          tc::mbar_wait(empty + slot, ((seq / nslot) & 1u) ^ 1u);
          ++seq;
          if (DBG && (a.dbg_flags & 1)) { tc::mbar_arrive(full + slot); continue; }
          tc::mbar_arrive_expect_tx(full + slot, npos * 16u * CG);
#pragma unroll
          for (int cg = 0; cg < CG; ++cg) {
            const uint4* src = a.x + ((static_cast<long long>(r.b) * CG + cg) * a.Ti + pa) * in_plane + q0;
            tc::bulk_g2s(slot_s + slot * slot_bytes + static_cast<uint32_t>(cg) * a.NP * 16u, src, npos * 16u, full + slot);
          }
        }
        g += r.ntiles;
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // The whole warp runs the (warp-uniform) control flow so that descriptor arithmetic stays in uniform registers;
    // only the tcgen05.mma / tcgen05.commit instructions themselves are issued by one elected lane.
    const bool leader = tc::elect_one();
    const uint32_t a_lbo = static_cast<uint32_t>(a.NP) * 16u;
    const uint32_t b_lbo = static_cast<uint32_t>(N) * 16u;
    // descriptor halves: hi = SBO (128 B) | version; lo = start address | LBO
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo_base = ((a_lbo >> 4) << 16);
    const uint32_t b_lo_base = ((b_lbo >> 4) << 16) | ((tc::smem_u32(w_s) >> 4) & 0x3fffu);
    const uint32_t slot_addr16 = tc::smem_u32(slot_s) >> 4;  // in 16-byte units
    const uint32_t slot_16 = slot_bytes >> 4;
    const uint32_t a_ks16 = 2u * (a_lbo >> 4), b_ks16 = 2u * (b_lbo >> 4), b_tap16 = static_cast<uint32_t>(CG) * (b_lbo >> 4);
    uint32_t tap16[9];  // A start offset of tap (kh,kw) in 16-byte units
#pragma unroll
    for (int hw = 0; hw < 9; ++hw) tap16[hw] = static_cast<uint32_t>((hw / 3) * a.Wi + (hw % 3));
    // instruction descriptor for N = nb blocks of CoP columns: idesc0 + nb * idesc_blk
    const uint32_t idesc0 = tc::umma_idesc(128, 0, /*bf16*/ 1, /*K-major*/ 0, 0);
    const uint32_t idesc_blk = static_cast<uint32_t>(CoP >> 3) << 17;
    tc::mbar_wait(wfull, 0);
    uint32_t seq = 0;            // staged input planes consumed so far
    uint32_t acc_phase = 0;      // bit s = number of uses of accumulator slot s so far, mod 2
    long long dbg_full = 0, dbg_tempty = 0, dbg_issue = 0, dbg_planes = 0;
    const long long dbg_t0 = DBG ? clock64() : 0;
    for (long long g = g_begin; g < g_end;) {
      const IgRun r = ig_run(g, g_end, a.To, a.tiles_q);
      const int nt = r.ntiles;
      int skipped = 0;  // number of consecutive skipped planes immediately before plane i
      for (int i = 0; i < nt + 2; ++i) {
        const long long c0 = DBG ? clock64() : 0;
        long long c1 = c0, c2 = c0;
        // the accumulator of output i is first written by this plane: it must have been read out by the epilogue
        if (i < nt) {
          const uint32_t s = static_cast<uint32_t>(kIgAccSlots - 1 - i);
          tc::mbar_wait(tempty + s, ((acc_phase >> s) & 1u) ^ 1u);
          tc::tc_fence_after();
        }
        c1 = DBG ? clock64() : 0;
        const int pa = r.t0 + i + a.plane_off;
        const bool skip = pa < a.zero_planes || pa >= a.Ti - a.zero_planes;
        if (!skip) {
          const uint32_t slot = seq % nslot;
          tc::mbar_wait(full + slot, (seq / nslot) & 1u);
          tc::tc_fence_after();
          c2 = DBG ? clock64() : 0;
          // blocks kt_lo..kt_hi of this plane exist (output i - kt inside the run); blocks kt < nfresh_end are fresh
          const int kt_lo = (i - (nt - 1)) > 0 ? (i - (nt - 1)) : 0;
          const int kt_hi = i < 2 ? i : 2;
          const int fresh_hi = skipped < kt_hi ? skipped : kt_hi;  // blocks kt <= fresh_hi have not been written yet
          const uint32_t a_base = a_lo_base | ((slot_addr16 + slot * slot_16) & 0x3fffu);
          // accumulator column of block kt: output i - kt lives in slot 15 - (i - kt)
          const uint32_t d_lo = tmem_base + static_cast<uint32_t>((kIgAccSlots - 1 - i + kt_lo) * CoP);
          const uint32_t b_base = b_lo_base + static_cast<uint32_t>(kt_lo * CoP);
          const int nb_all = kt_hi - kt_lo + 1;
          const int nb_fresh = fresh_hi >= kt_lo ? (fresh_hi - kt_lo + 1) : 0;
          if (leader) {
            // first MMA (tap 0, channel groups 0,1): overwrite the fresh blocks, accumulate into the others
            if (nb_fresh > 0) igemm_mma(d_lo, a_base, desc_hi, b_base, desc_hi, idesc0 + nb_fresh * idesc_blk, 0u);
            if (nb_all > nb_fresh)
              igemm_mma(d_lo + static_cast<uint32_t>(nb_fresh * CoP), a_base, desc_hi, b_base + static_cast<uint32_t>(nb_fresh * CoP),
                        desc_hi, idesc0 + (nb_all - nb_fresh) * idesc_blk, 1u);
            const uint32_t idesc = idesc0 + nb_all * idesc_blk;
#pragma unroll
            for (int hw = 0; hw < 9; ++hw) {
#pragma unroll
              for (int ks = 0; ks < CG / 2; ++ks) {
                if (hw == 0 && ks == 0) continue;
                igemm_mma(d_lo, a_base + tap16[hw] + ks * a_ks16, desc_hi, b_base + hw * b_tap16 + ks * b_ks16, desc_hi, idesc, 1u);
              }
            }
            tc::umma_commit(empty + slot);  // the input plane is consumed
          }
          ++seq;
          skipped = 0;
        } else {
          ++skipped;
        }
        // output i - 2 has received its last contribution
        if (i >= 2) {
          const uint32_t s = static_cast<uint32_t>(kIgAccSlots - 1 - (i - 2));
          if (leader) tc::umma_commit(tfull + s);
          acc_phase ^= 1u << s;
        }
        __syncwarp();
        if (DBG) { dbg_tempty += c1 - c0; dbg_full += c2 - c1; dbg_issue += clock64() - c2; ++dbg_planes; }
      }
      g += nt;
    }
    if (DBG && a.dbg && lane == 0) {
      long long* d = a.dbg + blockIdx.x * 8;
      d[0] = clock64() - dbg_t0; d[1] = dbg_full; d[2] = dbg_tempty; d[3] = dbg_issue; d[4] = dbg_planes;
    }
  } else {
    // =============================== epilogue (warps 2..17) ===============================
    // four groups x four TMEM lane quadrants; group g reads the accumulator slots s with s % 4 == g
    const int e = warp - 2;
    const int qd = warp & 3;               // the TMEM lane quadrant this warp may access
    const uint32_t grp = e >> 2;
    const int Cog = a.CogOut;
    const int Top = a.To + 2 * a.out_pad, Hop = a.Ho + 2 * a.out_pad, Wop = a.Wo + 2 * a.out_pad;
    const long long oplane = static_cast<long long>(Hop) * Wop;
    const long long mplane = static_cast<long long>(a.Ho) * a.Wo;
    const int row = qd * 32 + lane;
    uint32_t acc_phase = 0;
    long long dbg_tfull = 0, dbg_ld = 0, dbg_rest = 0;
    for (long long g = g_begin; g < g_end;) {
      const IgRun r = ig_run(g, g_end, a.To, a.tiles_q);
      const int q = r.qt * kIgTileM + row;
      const int ho = q / a.Wi, wo = q - ho * a.Wi;
      const bool valid = (ho < a.Ho) && (wo < a.Wo);
      // element offsets of this thread's position in the output / mask tensors at t = t0, channel group 0
      long long o_off = (static_cast<long long>(r.b) * Cog * Top + (r.t0 + a.out_pad)) * oplane +
                        static_cast<long long>(ho + a.out_pad) * Wop + (wo + a.out_pad);
      long long m_off = (static_cast<long long>(r.b) * Cog * a.To + r.t0) * mplane + static_cast<long long>(ho) * a.Wo + wo;
      long long o2_off = (static_cast<long long>(r.b) * Cog * a.To + r.t0) * a.QP2 + static_cast<long long>(ho) * (a.Wo + 2) + wo;
      const long long o2_cg = static_cast<long long>(a.To) * a.QP2;
      const long long o_cg = static_cast<long long>(Top) * oplane, m_cg = static_cast<long long>(a.To) * mplane;
      for (int j = 0; j < r.ntiles; ++j, o_off += oplane, m_off += mplane, o2_off += a.QP2) {
        const uint32_t s = static_cast<uint32_t>(kIgAccSlots - 1 - j);
        const uint32_t ph = (acc_phase >> s) & 1u;
        acc_phase ^= 1u << s;
        // an accumulator slot always belongs to the same group: its warps then wait on consecutive phases of
        // tfull[s] in order (a parity wait issued a whole phase early would pass immediately)
        if ((s & 3u) != grp) continue;
        // ReLU-mask source of the data gradient: issue the loads before waiting for the accumulators
        uint4 mk[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        if (a.mask && valid) {
#pragma unroll
          for (int cgi = 0; cgi < 4; ++cgi)
            if (cgi < Cog) mk[cgi] = __ldg(a.mask + m_off + cgi * m_cg);
        }
        const long long e0 = DBG ? clock64() : 0;
        tc::mbar_wait(tfull + s, ph);
        tc::tc_fence_after();
        const long long l0 = DBG ? clock64() : 0;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + s * CoP;
        uint32_t v[32];
        if (DBG && (a.dbg_flags & 16)) {
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = 0;
        } else if (CoP == 32) {
          tc::tmem_ld_32x32(taddr, v);
        } else {
          uint32_t h[16];
          tc::tmem_ld_32x16(taddr, h);
#pragma unroll
          for (int c = 0; c < 16; ++c) { v[c] = h[c]; v[16 + c] = 0; }
        }
        tc::tmem_ld_wait();
        // all TMEM reads of this warp are done: release the accumulator
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tempty + s);
        const long long l1 = DBG ? clock64() : 0;
        if (valid && !(DBG && (a.dbg_flags & 2))) {
#pragma unroll
          for (int cgi = 0; cgi < 4; ++cgi) {
            if (cgi >= Cog) continue;
            const float4 b0 = reinterpret_cast<const float4*>(bias_s)[2 * cgi], b1 = reinterpret_cast<const float4*>(bias_s)[2 * cgi + 1];
            const float eb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float f[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              float xv = __uint_as_float(v[cgi * 8 + jj]) + eb[jj];
              if (a.relu) xv = fmaxf(xv, 0.f);
              f[jj] = xv;
            }
            if (a.mask) {
              const uint32_t mw[4] = {mk[cgi].x, mk[cgi].y, mk[cgi].z, mk[cgi].w};
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                const uint32_t bits = (jj & 1) ? (mw[jj >> 1] >> 16) : (mw[jj >> 1] & 0xffffu);
                const float mv = __uint_as_float(bits << 16);
                f[jj] = (mv > 0.f) ? f[jj] : 0.f;
              }
            }
            const uint4 ov = make_uint4(tc::pack_bf16(f[0], f[1]), tc::pack_bf16(f[2], f[3]), tc::pack_bf16(f[4], f[5]),
                                        tc::pack_bf16(f[6], f[7]));
            a.y[o_off + cgi * o_cg] = ov;
            if (a.y2) a.y2[o2_off + cgi * o2_cg] = ov;
          }
        }
        if (DBG) { dbg_tfull += l0 - e0; dbg_ld += l1 - l0; dbg_rest += clock64() - l1; }
      }
      g += r.ntiles;
    }
    if (DBG && a.dbg && threadIdx.x == 64) {
      long long* d = a.dbg + blockIdx.x * 8;
      d[5] = dbg_tfull; d[6] = 0; d[7] = dbg_ld; d[4] = -dbg_rest;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, tmem_cols);
}

// [B][C][T][H][W] fp32 -> blocked bf16 [B][Cg][T+2p][H+2p][W+2p][8] interior (channels >= C are zero)
__global__ void nc_to_blocked_bf16_kernel(const float* __restrict__ x, uint4* __restrict__ y, int C, int Cg, int T, int H,
                                          int W, int pad, long long total) {
  const long long thw = static_cast<long long>(T) * H * W;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad, Tp = T + 2 * pad;
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
      f[j] = (c < C) ? x[(b * C + c) * thw + pos] : 0.f;
    }
    const int w = static_cast<int>(pos % W);
    const int h = static_cast<int>((pos / W) % H);
    const int t = static_cast<int>(pos / (static_cast<long long>(W) * H));
    y[((b * Cg + cg) * Tp + (t + pad)) * Hp * Wp + static_cast<long long>(h + pad) * Wp + (w + pad)] =
        make_uint4(tc::pack_bf16(f[0], f[1]), tc::pack_bf16(f[2], f[3]), tc::pack_bf16(f[4], f[5]), tc::pack_bf16(f[6], f[7]));
  }
}

// blocked bf16 [B][Cg][T][H][W][8] -> [B][C][T][H][W] fp32
__global__ void blocked_to_nc_f32_kernel(const uint4* __restrict__ x, float* __restrict__ y, int C, int Cg, long long thw,
                                         long long total) {
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pos = idx % thw;
    const long long r = idx / thw;
    const int cg = static_cast<int>(r % Cg);
    const long long b = r / Cg;
    const uint4 v = x[idx];
    const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cg * 8 + j;
      const uint32_t bits = (j & 1) ? (wv[j >> 1] >> 16) : (wv[j >> 1] & 0xffffu);
      if (c < C) y[(b * C + c) * thw + pos] = __uint_as_float(bits << 16);
    }
  }
}

long long* g_igemm_dbg = nullptr;  // set through pvb200_debug_set_igemm_counters (tools only)
int g_igemm_dbg_flags = 0;

static int igemm_cg(int C) { return 2 * ceil_div(C, 16); }  // channel groups of 8, padded to an even count (UMMA K = 16)

static int igemm_cop(int Co) { return Co <= 16 ? 16 : 32; }

static size_t igemm_ws_bytes(int Ci_role, int Co_role) {
  return static_cast<size_t>(27) * igemm_cg(Ci_role) * igemm_cop(Co_role) * 16;
}

static int launch_igemm(const void* xb, const float* w, long long s_co, long long s_ci, int flip, const float* bias,
                        const void* mask, void* yb, void* yb2, void* ws, size_t ws_bytes, int B, int Ci, int Ti, int Hi, int Wi, int Co,
                        int out_pad, int relu, int zero_planes, int To, int plane_off, cudaStream_t stream) {
  IgemmArgs a;
  a.x = static_cast<const uint4*>(xb);
  a.bias = bias;
  a.mask = static_cast<const uint4*>(mask);
  a.y = static_cast<uint4*>(yb);
  a.y2 = static_cast<uint4*>(yb2);
  a.B = B; a.Cg = igemm_cg(Ci); a.Ti = Ti; a.Hi = Hi; a.Wi = Wi;
  PVB_REQUIRE(Co <= 32, "conv3d_bf16: Cout=%d > 32 is not supported by the tensor-core path (use fp32 mode)", Co);
  a.CoP = igemm_cop(Co); a.Co = Co; a.CogOut = igemm_cg(Co);
  a.To = To; a.Ho = Hi - 2; a.Wo = Wi - 2;
  a.plane_off = plane_off;
  PVB_REQUIRE(a.To > 0 && a.Ho > 0 && a.Wo > 0, "conv3d_bf16: input %dx%dx%d too small", Ti, Hi, Wi);
  PVB_REQUIRE(a.Cg == 2 || a.Cg == 4, "conv3d_bf16: Cin=%d > 32 is not supported by the tensor-core path (use fp32 mode)", Ci);
  a.out_pad = out_pad; a.relu = relu;
  a.zero_planes = zero_planes;
  a.QP2 = static_cast<int>(round_up(static_cast<long long>(a.Ho) * (a.Wo + 2), 128LL));
  a.NP = round_up(kIgTileM + 2 * Wi + 2, 8);
  const int Qtot = (a.Ho - 1) * Wi + a.Wo;
  a.tiles_q = ceil_div(Qtot, kIgTileM);
  const int sms = sm_count();
  PVB_REQUIRE(sms > 0, "conv3d_bf16: no CUDA device");
  a.tiles = static_cast<long long>(B) * a.tiles_q * a.To;
  const size_t need = igemm_ws_bytes(Ci, Co);
  if (!ws || ws_bytes < need) {
    set_error("conv3d_bf16: workspace too small (%zu < %zu bytes)", ws_bytes, need);
    return PVB200_ERR_WORKSPACE;
  }
  PVB_REQUIRE(reinterpret_cast<uintptr_t>(xb) % 16 == 0 && reinterpret_cast<uintptr_t>(yb) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(ws) % 16 == 0, "conv3d_bf16: pointers must be 16-byte aligned");
  a.wq = static_cast<const uint4*>(ws);
  {
    const int total = 27 * a.Cg * a.CoP * 8;
    igemm_weight_prep_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(w, static_cast<__nv_bfloat16*>(ws), Ci, Co, a.Cg, a.CoP,
                                                                      s_co, s_ci, flip);
    PVB_LAUNCHED("igemm_weight_prep");
  }
  const size_t w_bytes = static_cast<size_t>(27) * a.Cg * a.CoP * 16;
  const size_t fixed = 640 + round_up(w_bytes, static_cast<size_t>(128));
  const size_t slot_bytes = static_cast<size_t>(a.Cg) * a.NP * 16;
  long long nslot = (227 * 1024 - static_cast<long long>(fixed)) / static_cast<long long>(slot_bytes);
  if (nslot > kIgMaxSlots) nslot = kIgMaxSlots;
  PVB_REQUIRE(nslot >= 2, "conv3d_bf16: Cin=%d Cout=%d width=%d does not fit in shared memory", Ci, Co, Wi);
  a.nslot = static_cast<int>(nslot);
  a.dbg = g_igemm_dbg;
  a.dbg_flags = g_igemm_dbg_flags;
  const size_t smem = fixed + nslot * slot_bytes;
  long long grid = a.tiles < sms ? a.tiles : sms;
#define PVB_IG_LAUNCH(CG, DBG)                                                                                           \
  do {                                                                                                                   \
    PVB_CUDA(cudaFuncSetAttribute(conv3d_igemm_bf16_kernel<CG, DBG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    conv3d_igemm_bf16_kernel<CG, DBG><<<static_cast<unsigned>(grid), kIgThreads, smem, stream>>>(a);                    \
  } while (0)
  if (a.dbg) {
    if (a.Cg == 2) PVB_IG_LAUNCH(2, true); else PVB_IG_LAUNCH(4, true);
  } else {
    if (a.Cg == 2) PVB_IG_LAUNCH(2, false); else PVB_IG_LAUNCH(4, false);
  }
#undef PVB_IG_LAUNCH
  PVB_LAUNCHED("conv3d_igemm_bf16");
  return PVB200_OK;
}

}  // namespace pvb

extern "C" {

/* tools only (not declared in pvb200.h): per-CTA cycle counters of the igemm kernel, [grid][8] long long */
void pvb200_debug_set_igemm_counters(long long* p) { pvb::g_igemm_dbg = p; }
void pvb200_debug_set_igemm_flags(int f) { pvb::g_igemm_dbg_flags = f; }

int pvb200_blocked_channel_groups(int C) { return pvb::igemm_cg(C); }

size_t pvb200_conv3d_bf16_workspace_bytes(int Cin, int Cout) {
  const size_t f = pvb::igemm_ws_bytes(Cin, Cout), d = pvb::igemm_ws_bytes(Cout, Cin);
  return f > d ? f : d;
}

int pvb200_nc_to_blocked_bf16(const float* x, uint16_t* y, int B, int C, int T, int H, int W, int pad,
                              pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(x && y && B > 0 && C > 0 && T > 0 && H > 0 && W > 0 && pad >= 0, "nc_to_blocked_bf16: bad argument");
  const int Cg = igemm_cg(C);
  const long long total = static_cast<long long>(B) * Cg * T * H * W;
  long long grid = ceil_div(total, 256LL);
  if (grid > 148 * 32) grid = 148 * 32;
  nc_to_blocked_bf16_kernel<<<static_cast<unsigned>(grid), 256, 0, as_stream(stream)>>>(x, reinterpret_cast<uint4*>(y), C, Cg,
                                                                                         T, H, W, pad, total);
  PVB_LAUNCHED("nc_to_blocked_bf16");
  return PVB200_OK;
}

int pvb200_blocked_to_nc_f32(const uint16_t* x, float* y, int B, int C, int T, int H, int W, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(x && y && B > 0 && C > 0 && T > 0 && H > 0 && W > 0, "blocked_to_nc_f32: bad argument");
  const int Cg = igemm_cg(C);
  const long long thw = static_cast<long long>(T) * H * W;
  const long long total = static_cast<long long>(B) * Cg * thw;
  long long grid = ceil_div(total, 256LL);
  if (grid > 148 * 32) grid = 148 * 32;
  blocked_to_nc_f32_kernel<<<static_cast<unsigned>(grid), 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(x), y, C,
                                                                                        Cg, thw, total);
  PVB_LAUNCHED("blocked_to_nc_f32");
  return PVB200_OK;
}

int pvb200_conv3d_fwd_bf16_tpad(const uint16_t* xb, const float* w, const float* bias, uint16_t* yb, void* workspace,
                                size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout, int relu, int out_pad,
                                int pad_t, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(xb && w && yb, "conv3d_fwd_bf16: null pointer");
  PVB_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && out_pad >= 0 && (pad_t == 0 || pad_t == 1), "conv3d_fwd_bf16: bad shape");
  return launch_igemm(xb, w, static_cast<long long>(Cin) * 27, 27, 0, bias, nullptr, yb, nullptr, workspace, workspace_bytes, B, Cin,
                      Ti, Hi, Wi, Cout, out_pad, relu, /*zero_planes=*/0, /*To=*/Ti + 2 * pad_t - 2, /*plane_off=*/-pad_t,
                      as_stream(stream));
}

int pvb200_conv3d_fwd_bf16(const uint16_t* xb, const float* w, const float* bias, uint16_t* yb, void* workspace,
                           size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout, int relu, int out_pad,
                           pvb200_stream_t stream) {
  return pvb200_conv3d_fwd_bf16_tpad(xb, w, bias, yb, workspace, workspace_bytes, B, Cin, Ti, Hi, Wi, Cout, relu, out_pad, 0, stream);
}

int pvb200_conv3d_dgrad_bf16_tpad(const uint16_t* gz_padded, const float* w, const uint16_t* mask_src, uint16_t* gx,
                                  uint16_t* gx_gzw, void* workspace, size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi,
                                  int Cout, int out_pad, int pad_t, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(gz_padded && w && gx, "conv3d_dgrad_bf16: null pointer");
  PVB_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && Ti + 2 * pad_t > 2 && Hi > 2 && Wi > 2 && out_pad >= 0 && (pad_t == 0 || pad_t == 1),
              "conv3d_dgrad_bf16: bad shape");
  // kernel input = gz (Ti + 2 pad_t - 2 planes) zero-padded by 2: [B][Cg(Cout)][Ti + 2 pad_t + 2][Hi+2][Wi+2];
  // kernel output = gx [B][Cg(Cin)][Ti][Hi][Wi]: gx[t] reads the padded planes t + pad_t + {0,1,2}
  return launch_igemm(gz_padded, w, /*s_co (out role = ci)*/ 27, /*s_ci (in role = co)*/ static_cast<long long>(Cin) * 27, 1,
                      nullptr, mask_src, gx, gx_gzw, workspace, workspace_bytes, B, /*Ci role*/ Cout, Ti + 2 * pad_t + 2, Hi + 2, Wi + 2,
                      /*Co role*/ Cin, out_pad, 0, /*zero_planes=*/2, /*To=*/Ti, /*plane_off=*/pad_t, as_stream(stream));
}

int pvb200_conv3d_dgrad_bf16(const uint16_t* gz_padded, const float* w, const uint16_t* mask_src, uint16_t* gx,
                             uint16_t* gx_gzw, void* workspace, size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout,
                             int out_pad, pvb200_stream_t stream) {
  return pvb200_conv3d_dgrad_bf16_tpad(gz_padded, w, mask_src, gx, gx_gzw, workspace, workspace_bytes, B, Cin, Ti, Hi, Wi, Cout, out_pad,
                                       0, stream);
}

}  // extern "C"
