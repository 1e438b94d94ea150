// conv3d_wgrad_bf16_rows.cu -- a11 in bf16 mode, round 2: the Conv3d 3x3x3 weight (and bias) gradient in the row-step
// formulation of conv3d_wgrad_bf16x3.cu, with the blocked bf16 tensors going STRAIGHT from the TMA engine into the
// tensor cores' operand layout.
//
// Reference: autograd of nn.Conv3d, predict_pv_yield/models/conv3d/model.py:80-90,117-120:
//   dW[co][ci][kt][kh][kw] = sum_{b,t,h,w} gz[b][co][t][h][w] * x[b][ci][t+kt][h+kh][w+kw],   db[co] = sum gz.
//
// One step = one input plane p of one sample and one output row h (positions are the K dimension, both operands MN-major):
//   * A (M = 128): box [Cg][3 rows][Wi] of x (one tiled TMA load through a tensor map, SASS UTMALDG): the rows h..h+2 of a
//     channel group are contiguous in memory, so M-group m = 3 cg + kh sits at the uniform stride Wi * 16 B the descriptor
//     needs -- the kh taps cost no copies; M-group 3 Cg is a constant row of ones (bias gradient); kw = descriptor start.
//   * B (N = 96): box [Cg][3 planes][WP] of the gradient (planes p-2 .. p, row h): the tensor map declares the row extent
//     Wo, so the K padding up to WP (a multiple of 16) and the planes outside the tensor arrive as ZEROS from the TMA
//     engine.  N-group n = 3 cg + tt (time tap kt = 2 - tt) at the uniform stride WP * 16 B: all three time taps in one MMA.
//   => 3 (kw) x WP/16 MMAs of 128 x 96 x 16 per step and 28 KB of operands; nothing touches the operands between TMA and MMA.
// The round-1 kernel (conv3d_wgrad_bf16.cu) staged 128-position tiles with a 2 Wi + 2 halo and the three time planes
// separately: 56 KB of L2 -> shared-memory traffic per 128 positions, 1.94x the algorithmic DRAM bytes, 0.35-0.41 of the
// tensor peak.  Accumulation: bf16 mode tolerates the tensor core's toward-zero fp32 accumulator (2e-2 bound), so the nine
// accumulators (3 kw x 96 columns) live in TMEM for the whole kernel and are drained once into a per-CTA partial; a second
// kernel reduces the partials in fixed order (deterministic).
// Warp roles (192 threads): warp 0 producer, warp 1 MMA issuer + TMEM owner, warps 2-5 final drain.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pvb {

constexpr int kWrThreads = 192;
constexpr int kWrMaxStages = 6;
constexpr int kWrAccCols = 96;
extern int g_dynamic_tiles;  // runtime.cu: pvb200_set_dynamic_tiles

struct WrArgs {
  float* partial;  // [grid][3 kw][96 columns][128 rows]
  int B, Cg, CgO, Ti, Hi, Wi, To, Ho, Wo, WP;
  int plane_off, gz_pad, nstage;
  long long steps;  // B * Ti * Ho
  int* sched;       // dynamic step ranges: global counter of claimed chunks (zeroed before the launch), or null = static split
};

constexpr int kWrRowBlock = 6;  // step order inside a sample: blocks of 6 output rows x all input planes x the rows of the block --
                                // a gradient plane is re-read by the steps of the next two input planes 12 steps later (L2 hits),
                                // not a whole plane of rows later (see conv3d_wgrad_bf16x3.cu, w3_step)
constexpr int kWrChunk = 16;    // steps per chunk of the dynamic split (pvb200_set_dynamic_tiles)
constexpr int kWrRing = 4;

__device__ __forceinline__ void wr_tma_5d(void* dst_smem, const CUtensorMap* tm, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
          tc::smem_u32(dst_smem)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// smem: [0,256) barriers and the ring of step ranges | stage s: A (a_bytes: box + ones rows + slack) then B (b_bytes)
__global__ void __launch_bounds__(kWrThreads, 1) conv3d_wgrad_bf16_rows_kernel(const WrArgs a, const __grid_constant__ CUtensorMap tm_x,
                                                                               const __grid_constant__ CUtensorMap tm_gz) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [6] operands landed
  uint64_t* empty = full + kWrMaxStages;                // [6] operands consumed
  uint64_t* done = empty + kWrMaxStages;                // [1] all MMAs complete
  uint64_t* rfull = done + 1;                           // [4] step range published by the producer
  uint64_t* rempty = rfull + kWrRing;                   // [4] step range read by the MMA warp
  uint64_t* flagbar = rempty + kWrRing;                 // [1] `any` flag written
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(flagbar + 1);
  volatile int* any_s = reinterpret_cast<volatile int*>(tmem_ptr + 1);  // did this CTA process a step?
  long long* r_lo = reinterpret_cast<long long*>(smem + 192);            // [4]
  long long* r_hi = r_lo + kWrRing;                                      // [4]  (lo >= hi: no more work)
  const int Cg = a.Cg, Wi = a.Wi, WP = a.WP;
  const uint32_t a_box = static_cast<uint32_t>(3 * Cg * Wi) * 16u;
  const uint32_t a_bytes = ((static_cast<uint32_t>(3 * Cg + 4) * Wi * 16u) + 127u) & ~127u;
  const uint32_t b_bytes = static_cast<uint32_t>(3 * a.CgO * WP) * 16u;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 127u) & ~127u);
  uint8_t* st_s = smem + 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t nstage = static_cast<uint32_t>(a.nstage);

  // zero everything once (slack rows read by the M = 128 instruction stay finite), then the ones rows of every stage
  {
    const uint32_t total16 = (nstage * stage_bytes) >> 4;
    uint4* z = reinterpret_cast<uint4*>(st_s);
    for (uint32_t i = threadIdx.x; i < total16; i += kWrThreads) z[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  for (uint32_t s = 0; s < nstage; ++s) {
    uint4* ones = reinterpret_cast<uint4*>(st_s + s * stage_bytes) + static_cast<uint32_t>(3 * Cg) * Wi;
    for (int i = threadIdx.x; i < 3 * Wi; i += kWrThreads) ones[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < kWrMaxStages; ++i) { tc::mbar_init(full + i, 1); tc::mbar_init(empty + i, 1); }
    tc::mbar_init(done, 1);
    tc::mbar_init(flagbar, 1);
    for (int i = 0; i < kWrRing; ++i) { tc::mbar_init(rfull + i, 1); tc::mbar_init(rempty + i, 1); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_ptr, 512);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // =============================== producer: claims the step ranges and fills the stages ===============================
    if (lane == 0) {
      uint32_t seq = 0;
      for (uint32_t ci = 0;; ++ci) {
        // static split: one contiguous range per CTA; dynamic: chunks of kWrChunk steps from an atomic counter, so that CTAs
        // displaced by NCCL's kernels under data parallelism cost a chunk, not a grid tail
        const uint32_t slot = ci % kWrRing;
        tc::mbar_wait(rempty + slot, ((ci / kWrRing) & 1u) ^ 1u);
        long long s_begin, s_end;
        if (a.sched) {
          s_begin = static_cast<long long>(atomicAdd(a.sched, 1)) * kWrChunk;
          s_end = s_begin + kWrChunk < a.steps ? s_begin + kWrChunk : a.steps;
        } else if (ci == 0) {
          s_begin = a.steps * blockIdx.x / gridDim.x;
          s_end = a.steps * (blockIdx.x + 1) / gridDim.x;
        } else {
          s_begin = s_end = a.steps;
        }
        r_lo[slot] = s_begin;
        r_hi[slot] = s_end;
        tc::mbar_arrive(rfull + slot);
        if (s_begin >= s_end) break;
        // decode (b, row block, p, row) once per range, then advance incrementally
        const long long per_b = static_cast<long long>(a.Ti) * a.Ho;
        int b = static_cast<int>(s_begin / per_b);
        const int rs = static_cast<int>(s_begin - b * per_b);
        const int k = rs / (kWrRowBlock * a.Ti);  // every block before the last one is full
        int h0 = k * kWrRowBlock;
        int hb = a.Ho - h0 < kWrRowBlock ? a.Ho - h0 : kWrRowBlock;
        const int q = rs - k * kWrRowBlock * a.Ti;
        int p = q / hb, h = h0 + q % hb;
        for (long long s = s_begin; s < s_end; ++s, ++seq) {
          const uint32_t stage = seq % nstage;
          tc::mbar_wait(empty + stage, ((seq / nstage) & 1u) ^ 1u);
          uint8_t* dst = st_s + stage * stage_bytes;
          tc::mbar_arrive_expect_tx(full + stage, a_box + b_bytes);
          wr_tma_5d(dst, &tm_x, 0, h, p, 0, b, full + stage);
          wr_tma_5d(dst + a_bytes, &tm_gz, 0, h, p - a.plane_off - 2, 0, b, full + stage);
          if (++h == h0 + hb) {
            h = h0;
            if (++p == a.Ti) {
              p = 0;
              h0 += hb;
              if (h0 >= a.Ho) { h0 = 0; ++b; }
              hb = a.Ho - h0 < kWrRowBlock ? a.Ho - h0 : kWrRowBlock;
              h = h0;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    const bool leader = tc::elect_one();
    const uint32_t a_hi_word = ((static_cast<uint32_t>(Wi) * 16u) >> 4) | (1u << 14);  // SBO = stride between channel groups
    const uint32_t b_hi_word = ((static_cast<uint32_t>(WP) * 16u) >> 4) | (1u << 14);
    const uint32_t lbo_word = (128u >> 4) << 16;  // the two 8-position K groups of a K = 16 step are contiguous
    const uint32_t st16 = tc::smem_u32(st_s) >> 4, stage16 = stage_bytes >> 4, a16 = a_bytes >> 4;
    const uint32_t idesc = tc::umma_idesc(128, 3 * a.CgO * 8, /*BF16*/ 1, /*A MN-major*/ 1, /*B MN-major*/ 1);
    const int k16n = WP >> 4;
    uint32_t seq = 0;
    for (uint32_t ci = 0;; ++ci) {
    const uint32_t slot = ci % kWrRing;
    tc::mbar_wait(rfull + slot, (ci / kWrRing) & 1u);
    const long long s_begin = r_lo[slot], s_end = r_hi[slot];
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(rempty + slot);
    if (s_begin >= s_end) break;
    for (long long s = s_begin; s < s_end; ++s, ++seq) {
      const uint32_t stage = seq % nstage;
      tc::mbar_wait(full + stage, (seq / nstage) & 1u);
      tc::tc_fence_after();
      const uint32_t a0 = lbo_word | ((st16 + stage * stage16) & 0x3fffu);
      const uint32_t b0 = lbo_word | ((st16 + stage * stage16 + a16) & 0x3fffu);
      if (leader) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const uint32_t d = tmem_base + static_cast<uint32_t>(kw * kWrAccCols);
#pragma unroll
          for (int k16 = 0; k16 < 4; ++k16) {
            if (k16 >= k16n) break;
            tc::umma_bf16_lohi(d, a0 + static_cast<uint32_t>(k16 * 16 + kw), a_hi_word, b0 + static_cast<uint32_t>(k16 * 16), b_hi_word, idesc,
                               (seq | static_cast<uint32_t>(k16)) ? 1u : 0u);
          }
        }
        tc::umma_commit(empty + stage);
      }
      __syncwarp();
    }
    }
    if (lane == 0) {
      *any_s = seq > 0 ? 1 : 0;
      __threadfence_block();
      tc::mbar_arrive(flagbar);
    }
    if (leader) tc::umma_commit(done);
    __syncwarp();
  } else {
    // =============================== final drain (warps 2..5) ===============================
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    tc::mbar_wait(done, 0);
    tc::mbar_wait(flagbar, 0);
    tc::tc_fence_after();
    float* mine = a.partial + static_cast<size_t>(blockIdx.x) * 3 * 128 * kWrAccCols + row;  // [kw][column][row]
    const bool any = *any_s != 0;
    for (int kw = 0; kw < 3; ++kw) {
#pragma unroll 1
      for (int c0 = 0; c0 < kWrAccCols; c0 += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + static_cast<uint32_t>(kw * kWrAccCols + c0), v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          mine[(static_cast<size_t>(kw) * kWrAccCols + c0 + j) * 128] = any ? __uint_as_float(v[j]) : 0.f;
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem_base, 512);
}

// column n of an accumulator = (cgo * 3 + tt) * 8 + c8 with kt = 2 - tt, co = cgo * 8 + c8; row = (3 * (ci / 8) + kh) * 8 + ci % 8;
// the ones row (bias) is row 3 * Cg * 8; bias tap kt_bias = pad_t (the tap for which every output plane meets an existing
// input plane exactly once)
__global__ void wgrad_bf16_rows_reduce_kernel(const float* __restrict__ partial, int ncta, float* __restrict__ dw, float* __restrict__ db,
                                              int Ci, int Co, int Cg, int kt_bias) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = Co * Ci * 27;
  const size_t per_cta = static_cast<size_t>(3) * 128 * kWrAccCols;
  if (idx < total) {
    const int tap = idx % 27;
    const int ci = (idx / 27) % Ci;
    const int co = idx / (27 * Ci);
    const int kt = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
    const int col = ((co >> 3) * 3 + (2 - kt)) * 8 + (co & 7);
    const size_t off = (static_cast<size_t>(kw) * kWrAccCols + col) * 128 + ((3 * (ci >> 3) + kh) * 8 + (ci & 7));
    float s = 0.f;
    for (int c = 0; c < ncta; ++c) s += partial[c * per_cta + off];
    dw[idx] = s;
  } else if (idx < total + Co && db) {
    const int co = idx - total;
    const int col = ((co >> 3) * 3 + (2 - kt_bias)) * 8 + (co & 7);
    const size_t off = static_cast<size_t>(col) * 128 + (3 * Cg) * 8;
    float s = 0.f;
    for (int c = 0; c < ncta; ++c) s += partial[c * per_cta + off];
    db[co] = s;
  }
}

// 5-D tensor map over 32-bit elements (a 16-byte blocked element = 4 of them) with explicit extents and byte strides; no
// swizzle, zero fill outside the extents
static int wr_make_tensor_map(CUtensorMap* tm, const void* base, const unsigned long long ext[5], const unsigned long long stride_bytes[4],
                              const unsigned box[5]) {
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
  for (int i = 0; i < 5; ++i) { gdim[i] = ext[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i < 4; ++i) gstr[i] = stride_bytes[i];
  return static_cast<int>(encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 5, const_cast<void*>(base), gdim, gstr, bx, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
}

static int wr_cg(int C) { return 2 * ceil_div(C, 16); }  // channel groups of 8, even (the blocked bf16 layout of the igemm kernels)

static size_t wr_stage_bytes(int Cg, int CgO, int Wi, int Wo) {
  const size_t a_bytes = round_up(static_cast<size_t>(3 * Cg + 4) * Wi * 16, static_cast<size_t>(128));
  const size_t b_bytes = round_up(static_cast<size_t>(3) * CgO * round_up(Wo, 16) * 16, static_cast<size_t>(128));
  return a_bytes + b_bytes;
}

}  // namespace pvb

extern "C" {

/* 1 when the row-step bf16 weight gradient takes this layer: Cin, Cout <= 32, rows of at most 64 positions (a TMA box
 * dimension holds 256 32-bit elements) */
int pvb200_conv3d_wgrad_bf16_rows_supported(int Cin, int Cout, int Hi, int Wi) {
  using namespace pvb;
  if (Cin <= 0 || Cout <= 0 || Cin > 32 || Cout > 32 || Hi < 3 || Wi < 3 || Wi > 64) return 0;
  const int Cg = wr_cg(Cin);
  // the M = 128 instruction reads 16 row groups at stride Wi * 16 B from the start of a stage: stay inside the allocation
  const size_t stage = wr_stage_bytes(Cg, wr_cg(Cout), Wi, Wi - 2);
  return (2 * stage + 256 <= 227 * 1024 && static_cast<size_t>(16) * Wi * 16 + (round_up(Wi, 16) + 16) * 16 <= 2 * stage) ? 1 : 0;
}

size_t pvb200_conv3d_wgrad_bf16_rows_workspace_bytes(void) {
  int sms = pvb::sm_count();
  if (sms <= 0) sms = 148;
  return static_cast<size_t>(sms) * 3 * 128 * pvb::kWrAccCols * sizeof(float) + 64;
}

/* dw [Cout][Cin][3][3][3], db [Cout] (or null), fp32, from x blocked bf16 [B][Cg(Cin)][Ti][Hi][Wi][8] and the pre-activation
 * gradient blocked bf16 zero-padded by gz_pad on T, H, W ([B][Cg(Cout)][To+2p][Ho+2p][Wo+2p][8], To = Ti + 2 pad_t - 2) */
int pvb200_conv3d_wgrad_bf16_rows(const uint16_t* xb, const uint16_t* gzb, int gz_pad, float* dw, float* db, void* workspace,
                                  size_t workspace_bytes, int B, int Cin, int Ti, int Hi, int Wi, int Cout, int pad_t,
                                  pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(xb && gzb && dw, "conv3d_wgrad_bf16_rows: null pointer");
  PVB_REQUIRE(B > 0 && gz_pad >= 0 && (pad_t == 0 || pad_t == 1), "conv3d_wgrad_bf16_rows: bad argument");
  PVB_REQUIRE(pvb200_conv3d_wgrad_bf16_rows_supported(Cin, Cout, Hi, Wi), "conv3d_wgrad_bf16_rows: Cin=%d Cout=%d plane %dx%d is not "
              "supported (use pvb200_conv3d_wgrad_bf16)", Cin, Cout, Hi, Wi);
  WrArgs a;
  a.B = B; a.Cg = wr_cg(Cin); a.CgO = wr_cg(Cout); a.Ti = Ti; a.Hi = Hi; a.Wi = Wi;
  a.To = Ti + 2 * pad_t - 2; a.Ho = Hi - 2; a.Wo = Wi - 2; a.WP = round_up(a.Wo, 16);
  PVB_REQUIRE(a.To > 0, "conv3d_wgrad_bf16_rows: input too short");
  a.plane_off = -pad_t; a.gz_pad = gz_pad;
  a.steps = static_cast<long long>(B) * Ti * a.Ho;
  const int sms = sm_count();
  PVB_REQUIRE(sms > 0, "conv3d_wgrad_bf16_rows: no CUDA device");
  long long grid = a.steps < sms ? a.steps : sms;
  const size_t need = static_cast<size_t>(grid) * 3 * 128 * kWrAccCols * sizeof(float) + 64;  // partials + the chunk counter
  if (!workspace || workspace_bytes < need) {
    set_error("conv3d_wgrad_bf16_rows: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return PVB200_ERR_WORKSPACE;
  }
  a.sched = nullptr;
  if (g_dynamic_tiles) {
    a.sched = reinterpret_cast<int*>(static_cast<uint8_t*>(workspace) + need - 64);
    PVB_CUDA(cudaMemsetAsync(a.sched, 0, sizeof(int), as_stream(stream)));
  }
  PVB_REQUIRE(reinterpret_cast<uintptr_t>(xb) % 16 == 0 && reinterpret_cast<uintptr_t>(gzb) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(workspace) % 16 == 0, "conv3d_wgrad_bf16_rows: pointers must be 16-byte aligned");
  a.partial = static_cast<float*>(workspace);
  const size_t stage = wr_stage_bytes(a.Cg, a.CgO, Wi, a.Wo);
  long long nstage = (227 * 1024 - 256) / static_cast<long long>(stage);
  if (nstage > kWrMaxStages) nstage = kWrMaxStages;
  a.nstage = static_cast<int>(nstage);
  const size_t smem = 256 + nstage * stage;
  CUtensorMap tm_x, tm_gz;
  {
    // x: [B][Cg][Ti][Hi][Wi] 16-byte elements = Wi * 4 words per row; box [1][Cg][1][3][Wi * 4]
    const unsigned long long xe[5] = {static_cast<unsigned long long>(Wi) * 4, static_cast<unsigned long long>(Hi), static_cast<unsigned long long>(Ti),
                                      static_cast<unsigned long long>(a.Cg), static_cast<unsigned long long>(B)};
    const unsigned long long xs[4] = {static_cast<unsigned long long>(Wi) * 16, static_cast<unsigned long long>(Hi) * Wi * 16,
                                      static_cast<unsigned long long>(Ti) * Hi * Wi * 16, static_cast<unsigned long long>(a.Cg) * Ti * Hi * Wi * 16};
    const unsigned xbx[5] = {static_cast<unsigned>(Wi) * 4, 3, 1, static_cast<unsigned>(a.Cg), 1};
    // gz: the VALID region [B][CgO][To][Ho][Wo] of the padded tensor (base shifted by the padding, strides of the padded
    // tensor): everything outside -- K padding up to WP, planes before / after the tensor -- is zero-filled by the TMA engine
    const unsigned long long Tz = a.To + 2 * gz_pad, Hz = a.Ho + 2 * gz_pad, Wz = a.Wo + 2 * gz_pad;
    const unsigned long long ge[5] = {static_cast<unsigned long long>(a.Wo) * 4, static_cast<unsigned long long>(a.Ho), static_cast<unsigned long long>(a.To),
                                      static_cast<unsigned long long>(a.CgO), static_cast<unsigned long long>(B)};
    const unsigned long long gs[4] = {Wz * 16, Hz * Wz * 16, Tz * Hz * Wz * 16, static_cast<unsigned long long>(a.CgO) * Tz * Hz * Wz * 16};
    const unsigned gbx[5] = {static_cast<unsigned>(a.WP) * 4, 1, 3, static_cast<unsigned>(a.CgO), 1};
    const uint8_t* gbase = reinterpret_cast<const uint8_t*>(gzb) + ((static_cast<size_t>(gz_pad) * Hz + gz_pad) * Wz + gz_pad) * 16;
    const int r1 = wr_make_tensor_map(&tm_x, xb, xe, xs, xbx);
    const int r2 = wr_make_tensor_map(&tm_gz, gbase, ge, gs, gbx);
    PVB_REQUIRE(r1 == 0 && r2 == 0, "conv3d_wgrad_bf16_rows: cuTensorMapEncodeTiled failed (%d, %d)", r1, r2);
  }
  PVB_CUDA(cudaFuncSetAttribute(conv3d_wgrad_bf16_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  conv3d_wgrad_bf16_rows_kernel<<<static_cast<unsigned>(grid), kWrThreads, smem, as_stream(stream)>>>(a, tm_x, tm_gz);
  PVB_LAUNCHED("conv3d_wgrad_bf16_rows");
  const int total = Cout * Cin * 27 + Cout;
  wgrad_bf16_rows_reduce_kernel<<<ceil_div(total, 128), 128, 0, as_stream(stream)>>>(a.partial, static_cast<int>(grid), dw, db, Cin, Cout,
                                                                                     a.Cg, pad_t);
  PVB_LAUNCHED("wgrad_bf16_rows_reduce");
  return PVB200_OK;
}

}  // extern "C"
