// conv3d_wgrad_f32.cu -- a11: Conv3d weight + bias gradient in fp32 (autograd of model.py:117-120).
//
//   dw[co,ci,kt,kh,kw] = sum_{b,t,h,w} gz[b,co,t,h,w] * x[b,ci,t+kt,h+kh,w+kw]      db[co] = sum gz
//
// A GEMM with M = Cout, N = Cin*27 and a reduction over B*To*Ho*Wo (1.1-2.1 M positions at B=32), so
// the whole [Cout x Cin*27] result lives in REGISTERS of one CTA (each thread: 4 output channels x
// 27 (or 9) taps for one input channel) while persistent CTAs stream disjoint position ranges through
// shared memory ("split-K"); a second tiny kernel sums the per-CTA partials in a fixed order
// (deterministic, no atomics).  FP32 FMA pipe bound: 432 FMAs (216 packed FFMA2) per 22 LDS per thread iteration.
// Positions use the same flattened-pitch trick as the forward kernel (q = ho*Wps + wo): taps are plain
// offsets kh*Wps + kw into a contiguous staged window; gz is zero-filled at the wrap columns.
//
// Two kernels share the FMA step: conv3d_wgrad_f32_ws_* (warp-specialised: producer warps issue every copy, consumer warps
// only run FMAs; fp32 input with even widths, i.e. every layer of the BASELINE models) and conv3d_wgrad_f32_kernel (every
// warp stages and computes; int16 input fused with the normalisation, odd widths, spatial padding).  Environment switches
// for A/B measurements (tools/bench_kernels.py): PVB200_WGRAD_WS=0 forces the second kernel, PVB200_WGRAD_NARROW=0|2|3
// selects the narrow-layer configuration (default 2).
#include "common.cuh"
#include "tc_common.cuh"

namespace pvb {

constexpr int kWgQC = 64;         // output positions staged per step
constexpr int kWgMaxCtas = 448;   // upper bound on persistent CTAs (workspace sizing): 3 per SM on 148 SMs
constexpr int kWgMaxCi = 64;

struct WgradArgs {
  const void* x;      // [B,Ci,Ti,Hi,Wi] fp32 or int16
  const float* mean;
  const float* stdv;
  const float* gz;    // [B,Co,To,Ho,Wo]
  float* partial;     // [gridDim.x][Co*Ci*27 + Co]
  int B, Ci, Ti, Hi, Wi, Co, To, Ho, Wo;
  int Wps, NP, NPs;   // pitch, staged positions per plane, padded plane stride (NPs % 8 == 4)
  int tiles_per_plane;
  long long total_steps;  // B * tiles_per_plane * To
  int ncog;           // ceil(Co / 4)
  int items;          // ncog * Ci * KTS
  int pairs;          // Wi even and 8-byte aligned tensors: stage two positions per copy (never straddles a row)
  int pad_t;          // time padding of the layer (0 or 1): input plane to + kt - pad_t, zero outside [0, Ti)
  int pad_hw;         // spatial padding (0 or 1): staged position (hp, wp) is input (hp - pad_hw, wp - pad_hw), zero outside
  int nprod;          // warp-specialised kernel: producer warps of the CTA (the last nprod warps)
};

// FMA work of one step (kWgQC positions) of one thread: 4 output channels (two PAIRS {2j, 2j+1}) x NKT*9 taps of one input
// channel.  gp: this thread's two channel-pair rows of the staged gz ([pair][q][2], see stage_gz); xk[k]: its input-channel
// row of time plane k.  The loop issues the packed fma.rn.f32x2 of sm_100 (SASS FFMA2: two IEEE fp32 FMAs per issue
// slot, bit-identical to two fmaf) with the input value broadcast to both halves.
template <int NKT>
__device__ __forceinline__ void wgrad_fma_step(const float* __restrict__ gp, const float* const (&xk)[NKT], int Wps,
                                               float2 (&acc2)[NKT * 9][2], float (&bacc)[4]) {
#pragma unroll 1
  for (int q = 0; q < kWgQC; q += 4) {
    float2 gv2[2][4];  // [co pair j][pos i] = {gz[2j][i], gz[2j+1][i]}
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float4 u = *reinterpret_cast<const float4*>(gp + (2 * j) * kWgQC + 2 * q);      // positions q, q+1
      const float4 v = *reinterpret_cast<const float4*>(gp + (2 * j) * kWgQC + 2 * q + 4);  // positions q+2, q+3
      gv2[j][0] = make_float2(u.x, u.y); gv2[j][1] = make_float2(u.z, u.w);
      gv2[j][2] = make_float2(v.x, v.y); gv2[j][3] = make_float2(v.z, v.w);
      bacc[2 * j] += (u.x + u.z) + (v.x + v.z);
      bacc[2 * j + 1] += (u.y + u.w) + (v.y + v.w);
    }
#pragma unroll
    for (int k = 0; k < NKT; ++k) {
      const float* xp = xk[k] + q;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const float4 v0 = *reinterpret_cast<const float4*>(xp + kh * Wps);
        const float2 v1 = *reinterpret_cast<const float2*>(xp + kh * Wps + 4);
        const float xv[6] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y};
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              acc2[k * 9 + kh * 3 + kw][j] =
                  __ffma2_rn(gv2[j][i], make_float2(xv[i + kw], xv[i + kw]), acc2[k * 9 + kh * 3 + kw][j]);
          }
        }
      }
    }
  }
}

// The same step with the input window CARRIED across iterations: positions q+4 .. q+7 loaded for the tail of iteration q are
// the head of iteration q+4, so each of the 9 (kt, kh) rows costs one LDS.128 per iteration instead of an LDS.128 + a
// (2-way bank-conflicted) LDS.64 -- half the shared-memory wavefronts of the loop.  Needs 36 more live registers: only the
// warp-specialised kernel, whose consumer warps take registers from the producer warps (setmaxnreg), can afford it.
__device__ __forceinline__ void wgrad_fma_step_carry(const float* __restrict__ gp, const float* const (&xk)[3], int Wps,
                                                     float2 (&acc2)[27][2], float (&bacc)[4]) {
  float4 cur[9];
#pragma unroll
  for (int r = 0; r < 9; ++r) cur[r] = *reinterpret_cast<const float4*>(xk[r / 3] + (r % 3) * Wps);
#pragma unroll 2
  for (int q = 0; q < kWgQC; q += 4) {
    float2 gv2[2][4];  // [co pair j][pos i] = {gz[2j][i], gz[2j+1][i]}
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const float4 u = *reinterpret_cast<const float4*>(gp + (2 * j) * kWgQC + 2 * q);      // positions q, q+1
      const float4 v = *reinterpret_cast<const float4*>(gp + (2 * j) * kWgQC + 2 * q + 4);  // positions q+2, q+3
      gv2[j][0] = make_float2(u.x, u.y); gv2[j][1] = make_float2(u.z, u.w);
      gv2[j][2] = make_float2(v.x, v.y); gv2[j][3] = make_float2(v.z, v.w);
      bacc[2 * j] += (u.x + u.z) + (v.x + v.z);
      bacc[2 * j + 1] += (u.y + u.w) + (v.y + v.w);
    }
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      // the row is staged NP = kWgQC + 2 Wps + 8 floats long: q + 4 .. q + 7 stays inside for every q < kWgQC
      const float4 nxt = *reinterpret_cast<const float4*>(xk[r / 3] + (r % 3) * Wps + q + 4);
      const float xv[6] = {cur[r].x, cur[r].y, cur[r].z, cur[r].w, nxt.x, nxt.y};
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            acc2[r * 3 + kw][j] = __ffma2_rn(gv2[j][i], make_float2(xv[i + kw], xv[i + kw]), acc2[r * 3 + kw][j]);
        }
      }
      cur[r] = nxt;
    }
  }
}

// this thread's partial: dw[4 cog + j][ci][kt][kh][kw] (+ db from the ci == 0, first-kt thread)
template <int NKT>
__device__ __forceinline__ void wgrad_write_partial(const WgradArgs& a, int cog, int ci, int kt0, bool writes_bias,
                                                    const float2 (&acc2)[NKT * 9][2], const float (&bacc)[4]) {
  float* part = a.partial + static_cast<long long>(blockIdx.x) * (static_cast<long long>(a.Co) * a.Ci * 27 + a.Co);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int co = 4 * cog + j;
    if (co >= a.Co) continue;
#pragma unroll
    for (int k = 0; k < NKT; ++k) {
#pragma unroll
      for (int t = 0; t < 9; ++t)
        part[(static_cast<long long>(co) * a.Ci + ci) * 27 + (kt0 + k) * 9 + t] =
            (j & 1) ? acc2[k * 9 + t][j >> 1].y : acc2[k * 9 + t][j >> 1].x;
    }
    if (writes_bias) part[static_cast<long long>(a.Co) * a.Ci * 27 + co] = bacc[j];
  }
}

// Work is ordered (b, tile, to) with `to` fastest: a CTA walks DOWN the time axis of one (sample, position tile)
// column, so consecutive steps share two of their three input planes.  Planes live in a 4-slot ring filled by
// cp.async (LDGSTS, zero-fill for out-of-range positions) one step ahead of the FMA loop.
template <bool kI16, int KTS>
__global__ void __launch_bounds__(KTS == 1 ? 256 : 384, KTS == 1 ? 1 : 2) conv3d_wgrad_f32_kernel(const WgradArgs a) {
  constexpr int NKT = 3 / KTS;  // kt values handled by one thread
  extern __shared__ __align__(16) float smem[];
  float* x_s = smem;                                        // [4 slots][Ci][NPs]
  float* gz_s = x_s + 4 * a.Ci * a.NPs;                     // [2][4*ncog][kWgQC]
  int* off_s = reinterpret_cast<int*>(gz_s + 2 * 4 * a.ncog * kWgQC);  // [NP] input-plane offsets
  int* goff_s = off_s + a.NP;                               // [kWgQC] gz-plane offsets

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  const int item = blockIdx.y * blockDim.x + tid;
  const bool active = item < a.items;
  int ci = 0, cog = 0, ktg = 0;
  if (active) {
    ci = item % a.Ci;
    const int r = item / a.Ci;
    cog = r % a.ncog;
    ktg = r / a.ncog;
  }

  // accumulators as output-channel PAIRS {2j, 2j+1} (wgrad_fma_step)
  float2 acc2[NKT * 9][2];
#pragma unroll
  for (int t = 0; t < NKT * 9; ++t)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc2[t][j] = make_float2(0.f, 0.f);
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};

  const long long g_begin = a.total_steps * blockIdx.x / gridDim.x;
  const long long g_end = a.total_steps * (blockIdx.x + 1) / gridDim.x;
  const long long xplane = static_cast<long long>(a.Hi) * a.Wi;
  const long long gplane = static_cast<long long>(a.Ho) * a.Wo;

  // stage input time plane `ti` of sample b (all Ci channels) into ring slot `slot`
  auto stage_x = [&](int b, int ti, int slot) {
    ti -= a.pad_t;
    if (ti < 0 || ti >= a.Ti) {  // a plane of the time padding: zeros (block-uniform)
      for (int i = tid; i < a.Ci * a.NPs; i += blockDim.x) x_s[slot * a.Ci * a.NPs + i] = 0.f;
      return;
    }
    for (int c = warp; c < a.Ci; c += nwarp) {
      float* dst = x_s + (slot * a.Ci + c) * a.NPs;
      const long long base = ((static_cast<long long>(b) * a.Ci + c) * a.Ti + ti) * xplane;
      if (kI16) {
        const int16_t* src = static_cast<const int16_t*>(a.x) + base;
        const float m = __ldg(a.mean + c), s = __ldg(a.stdv + c);
        for (int i = lane; i < a.NP; i += 32) {
          const int o = off_s[i];
          dst[i] = (o >= 0) ? sat_norm(__ldg(src + o), m, s) : 0.f;
        }
      } else if (a.pairs) {
        const float* src = static_cast<const float*>(a.x) + base;
        const uint32_t d0 = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
        for (int i = 2 * lane; i < a.NP; i += 64) {
          const int o = off_s[i];
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d0 + 4u * i), "l"(src + (o >= 0 ? o : 0)),
                       "r"(o >= 0 ? 8 : 0)
                       : "memory");
        }
      } else {
        const float* src = static_cast<const float*>(a.x) + base;
        const uint32_t d0 = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
        for (int i = lane; i < a.NP; i += 32) {
          const int o = off_s[i];
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d0 + 4u * i), "l"(src + (o >= 0 ? o : 0)),
                       "r"(o >= 0 ? 4 : 0)
                       : "memory");
        }
      }
    }
  };
  // stage gz of output time `to` into buffer `buf`: channel PAIRS interleaved, [co / 2][q][2], so that one LDS.128 of the
  // FMA loop delivers the {co, co + 1} operand pairs of two positions already aligned for the packed FMA; zero at the wrap
  // columns / beyond the plane
  auto stage_gz = [&](int b, int to, int buf) {
    for (int p = warp; p < 4 * a.ncog; p += nwarp) {
      float* dst = gz_s + (buf * 4 * a.ncog + (p & ~1)) * kWgQC + (p & 1);
      const bool ok_p = p < a.Co;
      const float* src = a.gz + ((static_cast<long long>(b) * a.Co + (ok_p ? p : 0)) * a.To + to) * gplane;
      const uint32_t d0 = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
      for (int i = lane; i < kWgQC; i += 32) {
        const int o = goff_s[i];
        const bool ok = ok_p && (o >= 0);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d0 + 8u * i), "l"(src + (ok ? o : 0)),
                     "r"(ok ? 4 : 0)
                     : "memory");
      }
    }
  };

  for (long long g = g_begin; g < g_end;) {
    // ---- a run: consecutive output times of one (b, tile) column ----
    const long long col = g / a.To;
    const int t0 = static_cast<int>(g - col * a.To);
    const int tile = static_cast<int>(col % a.tiles_per_plane);
    const int b = static_cast<int>(col / a.tiles_per_plane);
    const long long left = g_end - g;
    const int nstep = static_cast<int>(left < (a.To - t0) ? left : (a.To - t0));
    const int q0 = tile * kWgQC;

    __syncthreads();  // previous run fully consumed
    for (int i = tid; i < a.NP; i += blockDim.x) {
      const int pos = q0 + i;
      const int hp = pos / a.Wps, wp = pos - hp * a.Wps;
      const int hi = hp - a.pad_hw, wi = wp - a.pad_hw;
      off_s[i] = (hi >= 0 && hi < a.Hi && wi >= 0 && wi < a.Wi) ? hi * a.Wi + wi : -1;
    }
    for (int i = tid; i < kWgQC; i += blockDim.x) {
      const int pos = q0 + i;
      const int ho = pos / a.Wps, wo = pos - ho * a.Wps;
      goff_s[i] = (ho < a.Ho && wo < a.Wo) ? ho * a.Wo + wo : -1;
    }
    __syncthreads();
    // prologue: the three planes of the first step + its gz
    stage_x(b, t0, 0);
    stage_x(b, t0 + 1, 1);
    stage_x(b, t0 + 2, 2);
    stage_gz(b, t0, 0);
    asm volatile("cp.async.commit_group;" ::: "memory");

    for (int s = 0; s < nstep; ++s) {
      if (s + 1 < nstep) {  // prefetch the one new plane of the next step, and its gz
        stage_x(b, t0 + s + 3, (s + 3) & 3);
        stage_gz(b, t0 + s + 1, (s + 1) & 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
      __syncthreads();

      if (active) {
        const float* gp = gz_s + ((s & 1) * 4 * a.ncog + 4 * cog) * kWgQC;
        const float* xk[NKT];
#pragma unroll
        for (int k = 0; k < NKT; ++k) {
          const int kt = (KTS == 1) ? k : ktg;
          xk[k] = x_s + (((s + kt) & 3) * a.Ci + ci) * a.NPs;
        }
wgrad_fma_step<NKT>(gp, xk, a.Wps, acc2, bacc);
      }
      __syncthreads();  // slot (s & 3) and gz buffer (s & 1) may be overwritten by the next prefetch
    }
    g += nstep;
  }

  // ---- write this CTA's partial ----
  if (active) wgrad_write_partial<NKT>(a, cog, ci, (KTS == 1) ? 0 : ktg, ci == 0 && ktg == 0, acc2, bacc);
}

// ---- warp-specialised variant (wide layers, fp32 input, even widths) -----------------------------------------------
// ncu of the kernel above (profiles/ncu_fp32_r01c.txt): with the packed FMA the loop is no longer issue bound, but every
// warp also runs the staging code between two barriers of every step, and the FMA pipe idles meanwhile (23 % of the stall
// samples outside the FMA loop, pipe 65 % busy).  Here the 8 consumer warps ONLY run FMAs; four extra PRODUCER warps (one
// per scheduler, so that every scheduler carries the same load: with a single producer warp its scheduler was 13 % slower
// and the other three waited for it at every step) issue every cp.async of the CTA and run up to kWsD steps ahead:
//   * per step a "package" = the step's gz tile + its NEW input planes (3 at the start of a run, else 1);
//   * full[d] / empty[d] mbarriers per package slot d = n % kWsD (n = global step counter of the CTA): the 128 producer
//     lanes arrive on full[d] through cp.async.mbarrier.arrive.noinc (fires when the lane's copies have landed), each
//     consumer warp arrives on empty[d] after its FMA loop;
//   * input planes live in a ring of kWsR = 3 * kWsD slots addressed by a running position: the planes of the (at most
//     kWsD - 1) steps still in flight plus the package being loaded span at most 3 * kWsD positions, so a slot is never
//     overwritten while a consumer may still read it -- also across run boundaries, which removes the exposed
//     three-plane prologue of every run;
//   * copy offsets of a lane are the same for every channel and plane of a run: kept in registers, no tables.  Input
//     positions outside the plane (wrap columns, rows past the end) are only ever multiplied by zero-filled gz entries,
//     so they need no zero fill, only finite contents: the ring is zeroed once at kernel start and their copies are
//     redirected to a dummy pair in the padding of the row (branch-free producer loop).
constexpr int kWsD = 2;         // packages in flight
constexpr int kWsR = 3 * kWsD;  // input plane ring slots
constexpr int kWsK = 6;         // 8-byte copies per lane and channel row: NP <= 64 * kWsK
constexpr int kWsMaxProd = 4;   // producer warps of a wide layer (256 consumer threads); narrow layers use one
constexpr int kWsMaxThreads = 256 + 32 * kWsMaxProd;

__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// kMode 1: the 8 consumer + 4 producer warp configuration of the wide layers (register reallocation + carried input
// window; ptxas only honours setmaxnreg when it is unconditional, hence a template parameter and not a run-time flag).
// kMode 2: narrow layers, <= 96 consumer threads + 1 producer warp (4 warps: two CTAs per SM put two warps on every
// scheduler, whose 16 K registers then allow 255 per thread; a fifth warp would cap it at 168) without any
// reallocation, so the carried window fits as well.  kMode 0: generic (168 registers, no carried window).
template <int kMode>
__device__ __forceinline__ void wgrad_ws_body(const WgradArgs& a) {
  extern __shared__ __align__(16) float smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);  // [kWsD]
  uint64_t* empty = full + kWsD;                       // [kWsD]
  float* x_s = smem + 16;                              // [kWsR][Ci][NPs]   (64 bytes reserved for the barriers)
  float* gz_s = x_s + kWsR * a.Ci * a.NPs;             // [kWsD][4*ncog / 2][kWgQC][2]

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int nprod = a.nprod;
  const int ncons = blockDim.x - 32 * nprod;  // consumer threads (a multiple of 32); the last nprod warps produce
  if (tid == 0) {
    for (int i = 0; i < kWsD; ++i) {
      tc::mbar_init(full + i, 32 * nprod);
      tc::mbar_init(empty + i, ncons / 32);
    }
    tc::fence_barrier_init();
  }
  for (int i = tid; i < kWsR * a.Ci * a.NPs; i += blockDim.x) x_s[i] = 0.f;  // finite contents everywhere (see above)
  __syncthreads();

  const long long g_begin = a.total_steps * blockIdx.x / gridDim.x;
  const long long g_end = a.total_steps * (blockIdx.x + 1) / gridDim.x;
  uint32_t n = 0;  // global step counter: package slot n % kWsD, barrier phase (n / kWsD) & 1
  int pos = 0;     // ring slot of the kt = 0 plane of the current step
  int next = 0;    // ring slot of the first plane of the next run

  // register reallocation between the roles (wide layers: 8 consumer + 4 producer warps = three warpgroups of a CTA that
  // was allocated 168 registers per thread): producers keep 56, consumers grow to 224 (8*224 + 4*56 = 12*168)
  if (tid >= ncons) {
    if (kMode == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    // =========================== producer warps ===========================
    const int pw = (tid - ncons) >> 5;  // this warp stages input channels / gz rows pw, pw + nprod, ...
    const long long xplane = static_cast<long long>(a.Hi) * a.Wi;
    const long long gplane = static_cast<long long>(a.Ho) * a.Wo;
    const uint32_t x_u32 = tc::smem_u32(x_s), gz_u32 = tc::smem_u32(gz_s);
    const uint32_t row_bytes = static_cast<uint32_t>(a.NPs) * 4u;
    const uint32_t slot_bytes = static_cast<uint32_t>(a.Ci) * row_bytes;
    const long long x_cstride = static_cast<long long>(a.Ti) * xplane * 4;  // bytes between channels of x
    const long long g_cstride = static_cast<long long>(a.To) * gplane * 4;  // bytes between channels of gz
    const bool quad = (a.Wi % 4 == 0) && (reinterpret_cast<uintptr_t>(a.x) % 16 == 0);
    const int epc = quad ? 4 : 2;                     // elements per copy
    const int nk = (a.NP + 32 * epc - 1) / (32 * epc);  // copies per lane and row
    uint32_t xo[kWsK];  // byte offset inside an input plane of staged positions epc * (lane + 32k) ..; 0 for the dummy copy
    uint32_t xd[kWsK];  // byte offset inside the shared-memory row; the row padding (float NP) for the dummy copy
    uint32_t go[2];     // byte offset inside a gz plane of positions lane + 32k
    uint32_t gs[2];     // 4 or 0 (zero fill at the wrap columns / beyond the plane: these MUST be zero)

    auto stage_x = [&](int b, int ti, int slot) {
      ti -= a.pad_t;
      uint32_t dst = x_u32 + static_cast<uint32_t>(slot) * slot_bytes;
      if (ti < 0 || ti >= a.Ti) {  // a plane of the time padding: zero-fill copies over the whole slot
        for (uint32_t j = 8u * (pw * 32 + lane); j < slot_bytes; j += 256u * nprod)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8, 0;" ::"r"(dst + j), "l"(a.x) : "memory");
        return;
      }
      dst += pw * row_bytes;
      const char* src = reinterpret_cast<const char*>(static_cast<const float*>(a.x) +
                                                      (static_cast<long long>(b) * a.Ci * a.Ti + ti) * xplane) +
                        pw * x_cstride;
      if (quad) {  // 16-byte copies (Wi a multiple of 4): a quarter of the copy elements the LSU has to process
        for (int c = pw; c < a.Ci; c += nprod) {
#pragma unroll
          for (int k = 0; k < kWsK / 2; ++k) {
            if (k < nk)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + xd[k]), "l"(src + xo[k]) : "memory");
          }
          dst += nprod * row_bytes;
          src += nprod * x_cstride;
        }
        return;
      }
      for (int c = pw; c < a.Ci; c += nprod) {
#pragma unroll
        for (int k = 0; k < kWsK; ++k) {
          if (k < nk)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + xd[k]), "l"(src + xo[k]) : "memory");
        }
        dst += nprod * row_bytes;
        src += nprod * x_cstride;
      }
    };
    // gz of output time `to`: channel PAIRS interleaved, [co / 2][q][2], so that one LDS.128 of the FMA loop delivers the
    // {co, co + 1} operand pairs of two positions already aligned for the packed FMA
    auto stage_gz = [&](int b, int to, int d) {
      const uint32_t dst = gz_u32 + static_cast<uint32_t>(d * 4 * a.ncog * kWgQC) * 4u + 8u * lane;
      const char* src = reinterpret_cast<const char*>(a.gz + (static_cast<long long>(b) * a.Co * a.To + to) * gplane);
      for (int p = pw; p < 4 * a.ncog; p += nprod) {
        const bool ok_p = p < a.Co;
        const uint32_t dp = dst + static_cast<uint32_t>((p & ~1) * kWgQC + (p & 1)) * 4u;
        const char* sp = src + (ok_p ? p : 0) * g_cstride;
#pragma unroll
        for (int k = 0; k < 2; ++k)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dp + 256u * k), "l"(sp + go[k]),
                       "r"(ok_p ? gs[k] : 0u)
                       : "memory");
      }
    };

    for (long long g = g_begin; g < g_end;) {
      const long long col = g / a.To;
      const int t0 = static_cast<int>(g - col * a.To);
      const int tile = static_cast<int>(col % a.tiles_per_plane);
      const int b = static_cast<int>(col / a.tiles_per_plane);
      const long long left = g_end - g;
      const int nstep = static_cast<int>(left < (a.To - t0) ? left : (a.To - t0));
      const int q0 = tile * kWgQC;
      // this lane's copy offsets for the run (pad_hw == 0 and Wi a multiple of epc: a copy never straddles a row)
#pragma unroll
      for (int k = 0; k < kWsK; ++k) {
        const int i = epc * (lane + 32 * k);
        const int p = q0 + i;
        const int hi = p / a.Wps, wi = p - hi * a.Wps;
        const bool ok = i < a.NP && hi < a.Hi && wi < a.Wi;
        xo[k] = ok ? static_cast<uint32_t>(hi * a.Wi + wi) * 4u : 0u;
        xd[k] = ok ? 4u * i : 4u * a.NP;
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int p = q0 + lane + 32 * k;
        const int ho = p / a.Wps, wo = p - ho * a.Wps;
        const bool ok = ho < a.Ho && wo < a.Wo;
        go[k] = ok ? static_cast<uint32_t>(ho * a.Wo + wo) * 4u : 0u;
        gs[k] = ok ? 4u : 0u;
      }
      for (int s = 0; s < nstep; ++s, ++n) {
        const int d = n % kWsD;
        tc::mbar_wait(empty + d, ((n / kWsD) & 1u) ^ 1u);  // step n - kWsD consumed (passes at once for the first kWsD)
        if (s == 0) {
          pos = next;
          stage_x(b, t0, pos);
          stage_x(b, t0 + 1, (pos + 1) % kWsR);
          stage_x(b, t0 + 2, (pos + 2) % kWsR);
        } else {
          pos = (pos + 1) % kWsR;
          stage_x(b, t0 + s + 2, (pos + 2) % kWsR);
        }
        next = (pos + 3) % kWsR;
        stage_gz(b, t0 + s, d);
        cp_async_arrive_noinc(full + d);
      }
      g += nstep;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    return;
  }

  // =========================== consumer warps ===========================
  if (kMode == 1) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
  const int item = blockIdx.y * ncons + tid;
  const bool active = item < a.items;
  int ci = 0, cog = 0;
  if (active) {
    ci = item % a.Ci;
    cog = item / a.Ci;
  }
  float2 acc2[27][2];
#pragma unroll
  for (int t = 0; t < 27; ++t)
#pragma unroll
    for (int j = 0; j < 2; ++j) acc2[t][j] = make_float2(0.f, 0.f);
  float bacc[4] = {0.f, 0.f, 0.f, 0.f};

  for (long long g = g_begin; g < g_end;) {
    const long long col = g / a.To;
    const int t0 = static_cast<int>(g - col * a.To);
    const long long left = g_end - g;
    const int nstep = static_cast<int>(left < (a.To - t0) ? left : (a.To - t0));
    for (int s = 0; s < nstep; ++s, ++n) {
      const int d = n % kWsD;
      pos = (s == 0) ? next : (pos + 1) % kWsR;
      next = (pos + 3) % kWsR;
      tc::mbar_wait(full + d, (n / kWsD) & 1u);
      if (active) {
        const float* gp = gz_s + (d * 4 * a.ncog + 4 * cog) * kWgQC;
        const float* xk[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) xk[k] = x_s + (((pos + k) % kWsR) * a.Ci + ci) * a.NPs;
        if (kMode != 0)
          wgrad_fma_step_carry(gp, xk, a.Wps, acc2, bacc);
        else
          wgrad_fma_step<3>(gp, xk, a.Wps, acc2, bacc);
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(empty + d);  // this warp is done with package slot d (and the plane it retires)
    }
    g += nstep;
  }
  if (active) wgrad_write_partial<3>(a, cog, ci, 0, ci == 0, acc2, bacc);
}

// one entry point per mode: the launch bounds (and with them the register budget ptxas compiles for) differ
__global__ void __launch_bounds__(kWsMaxThreads, 1) conv3d_wgrad_f32_ws_kernel(const WgradArgs a) { wgrad_ws_body<0>(a); }
__global__ void __launch_bounds__(kWsMaxThreads, 1) conv3d_wgrad_f32_ws_wide_kernel(const WgradArgs a) { wgrad_ws_body<1>(a); }
__global__ void __launch_bounds__(128, 2) conv3d_wgrad_f32_ws_narrow_kernel(const WgradArgs a) { wgrad_ws_body<2>(a); }

// dw[i] = sum over CTAs (fixed order) ; the last Co entries are db
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, float* __restrict__ db,
                                    int n_w, int n_b, int n_part) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = n_w + n_b;
  if (i >= n) return;
  float s = 0.f;
  for (int p = 0; p < n_part; ++p) s += partial[static_cast<long long>(p) * n + i];
  if (i < n_w) dw[i] = s;
  else if (db) db[i - n_w] = s;
}

// PVB200_WGRAD_WS=0 in the environment selects the non-specialised kernel (A/B measurements, tools/bench_kernels.py)
static const bool g_wgrad_ws_enabled = [] {
  const char* e = getenv("PVB200_WGRAD_WS");
  return !(e && e[0] == '0');
}();

// narrow layers: 2 = two CTAs per SM with the carried input window (default), 3 = three CTAs per SM without it, 0 = two without
static const int g_wgrad_narrow_mode = [] {
  const char* e = getenv("PVB200_WGRAD_NARROW");
  return (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 2;
}();

template <bool kI16, int KTS>
static int launch_wgrad(WgradArgs a, int grid_x, int grid_y, int threads, size_t smem, cudaStream_t stream) {
  PVB_CUDA(cudaFuncSetAttribute(conv3d_wgrad_f32_kernel<kI16, KTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  conv3d_wgrad_f32_kernel<kI16, KTS><<<dim3(grid_x, grid_y), threads, smem, stream>>>(a);
  PVB_LAUNCHED("conv3d_wgrad_f32");
  return PVB200_OK;
}

}  // namespace pvb

extern "C" {

size_t pvb200_conv3d_wgrad_workspace_bytes(int Cin, int Cout) {
  return static_cast<size_t>(pvb::kWgMaxCtas) * (static_cast<size_t>(Cout) * Cin * 27 + Cout) * sizeof(float);
}

int pvb200_conv3d_wgrad_f32(const void* x, int x_is_i16, const float* mean, const float* std, const float* gz,
                            float* dw, float* db, void* workspace, size_t workspace_bytes, int B, int Cin, int Ti,
                            int Hi, int Wi, int Cout, pvb200_stream_t stream) {
  return pvb200_conv3d_wgrad_f32_tpad(x, x_is_i16, mean, std, gz, dw, db, workspace, workspace_bytes, B, Cin, Ti, Hi, Wi, Cout, 0,
                                      stream);
}

int pvb200_conv3d_wgrad_f32_tpad(const void* x, int x_is_i16, const float* mean, const float* std, const float* gz,
                                 float* dw, float* db, void* workspace, size_t workspace_bytes, int B, int Cin, int Ti,
                                 int Hi, int Wi, int Cout, int pad_t, pvb200_stream_t stream) {
  return pvb200_conv3d_wgrad_f32_pad(x, x_is_i16, mean, std, gz, dw, db, workspace, workspace_bytes, B, Cin, Ti, Hi, Wi, Cout, pad_t,
                                     0, stream);
}

int pvb200_conv3d_wgrad_f32_pad(const void* x, int x_is_i16, const float* mean, const float* std, const float* gz,
                                float* dw, float* db, void* workspace, size_t workspace_bytes, int B, int Cin, int Ti,
                                int Hi, int Wi, int Cout, int pad_t, int pad_hw, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(x && gz && dw, "conv3d_wgrad: null pointer");
  PVB_REQUIRE((pad_t == 0 || pad_t == 1) && (pad_hw == 0 || pad_hw == 1), "conv3d_wgrad: padding (%d, %d) not in {0, 1}", pad_t, pad_hw);
  PVB_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && Ti + 2 * pad_t > 2 && Hi + 2 * pad_hw > 2 && Wi + 2 * pad_hw > 2, "conv3d_wgrad: bad shape");
  PVB_REQUIRE(Cin <= kWgMaxCi, "conv3d_wgrad: Cin=%d > %d not supported", Cin, kWgMaxCi);
  PVB_REQUIRE(!x_is_i16 || (mean && std), "conv3d_wgrad: int16 input needs mean/std");
  const size_t need = pvb200_conv3d_wgrad_workspace_bytes(Cin, Cout);
  if (!workspace || workspace_bytes < need) {
    set_error("conv3d_wgrad: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return PVB200_ERR_WORKSPACE;
  }
  WgradArgs a;
  a.x = x; a.mean = mean; a.stdv = std; a.gz = gz; a.partial = static_cast<float*>(workspace);
  a.B = B; a.Ci = Cin; a.Ti = Ti; a.Hi = Hi; a.Wi = Wi; a.Co = Cout;
  a.To = Ti + 2 * pad_t - 2; a.Ho = Hi + 2 * pad_hw - 2; a.Wo = Wi + 2 * pad_hw - 2;
  a.pad_t = pad_t; a.pad_hw = pad_hw;
  a.Wps = round_up(Wi + 2 * pad_hw, 4);
  a.NP = kWgQC + 2 * a.Wps + 8;
  a.NPs = round_up(a.NP, 8) + 4;
  const int Qtot = (a.Ho - 1) * a.Wps + a.Wo;
  a.tiles_per_plane = ceil_div(Qtot, kWgQC);
  a.total_steps = static_cast<long long>(B) * a.tiles_per_plane * a.To;
  a.ncog = ceil_div(Cout, 4);
  // narrow layers (conv0: Cin = 12) split the 27 taps over 3 threads to keep the CTA full
  const int kts = (a.ncog * Cin <= 128) ? 3 : 1;
  a.items = a.ncog * Cin * kts;
  a.pairs = (!x_is_i16 && pad_hw == 0 && Wi % 2 == 0 && reinterpret_cast<uintptr_t>(x) % 8 == 0 &&
             reinterpret_cast<uintptr_t>(gz) % 8 == 0) ? 1 : 0;
  const int cap = (kts == 1) ? 256 : 384;
  const int grid_y = ceil_div(a.items, cap);
  const int threads = round_up(ceil_div(a.items, grid_y), 32);
  const int sms = sm_count();
  PVB_REQUIRE(sms > 0, "conv3d_wgrad: no CUDA device");
  const size_t smem = (static_cast<size_t>(4) * Cin * a.NPs + 2 * 4 * a.ncog * kWgQC) * sizeof(float) +
                      (a.NP + kWgQC) * sizeof(int);
  // persistent CTAs: one per SM for the wide layers (254 registers, ~120 KB of shared memory), two per SM for the
  // narrow ones (kts == 3: 36 accumulators per thread), which doubles the warps that hide the shared-memory latency
  long long gx = (kts == 3 && 2 * smem <= 220 * 1024) ? 2LL * sms : sms;
  if (gx > kWgMaxCtas) gx = kWgMaxCtas;
  if (gx > a.total_steps) gx = a.total_steps;
  PVB_REQUIRE(smem <= 227 * 1024, "conv3d_wgrad: Cin=%d, width %d needs %zu B of shared memory (> 227 KB)", Cin, Wi, smem);
  cudaStream_t st = as_stream(stream);
  int rc;
  // fp32 input and even widths: warp-specialised kernel (producer warps + FMA-only consumer warps), every consumer thread
  // owning all 27 taps of (4 output channels, 1 input channel).  Wide layers: 256 consumer threads + 4 producer warps, one
  // CTA per SM.  Narrow layers (conv0: 8 x 12 = 96 items): 3 consumer warps + 1 producer warp, two CTAs per SM.
  const int ws_items = a.ncog * Cin;
  const int ws_grid_y = ceil_div(ws_items, 256);
  const int ws_cons = round_up(ceil_div(ws_items, ws_grid_y), 32);
  const int ws_prod = ws_cons > 128 ? kWsMaxProd : 1;
  const size_t smem_ws = 64 + (static_cast<size_t>(kWsR) * Cin * a.NPs + kWsD * 4 * a.ncog * kWgQC) * sizeof(float);
  const bool use_ws = !x_is_i16 && pad_hw == 0 && Wi % 2 == 0 && reinterpret_cast<uintptr_t>(x) % 8 == 0 &&
                      a.NP <= 64 * kWsK && smem_ws <= 227 * 1024 && g_wgrad_ws_enabled;
  if (use_ws) {
    a.items = ws_items;
    a.nprod = ws_prod;
    // CTAs per SM: as many as registers (168 x threads; 200 in the carried-window narrow mode, at most two) and shared memory allow
    int per_sm = 1;
    if (2 * (ws_cons + 32 * ws_prod) <= 384 && 2 * (smem_ws + 1024) <= 228 * 1024) per_sm = 2;
    if (g_wgrad_narrow_mode == 3 && 3 * (ws_cons + 32 * ws_prod) <= 384 && 3 * (smem_ws + 1024) <= 228 * 1024) per_sm = 3;
    long long wgx = static_cast<long long>(per_sm) * sms;
    if (wgx > kWgMaxCtas) wgx = kWgMaxCtas;
    if (wgx > a.total_steps) wgx = a.total_steps;
    gx = wgx;
    const dim3 wgrid((unsigned)gx, ws_grid_y);
    const int wthreads = ws_cons + 32 * ws_prod;
    if (ws_prod == kWsMaxProd && ws_cons == 256) {
      PVB_CUDA(cudaFuncSetAttribute(conv3d_wgrad_f32_ws_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws));
      conv3d_wgrad_f32_ws_wide_kernel<<<wgrid, wthreads, smem_ws, st>>>(a);
    } else if (wthreads <= 128 && g_wgrad_narrow_mode == 2) {
      PVB_CUDA(cudaFuncSetAttribute(conv3d_wgrad_f32_ws_narrow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws));
      conv3d_wgrad_f32_ws_narrow_kernel<<<wgrid, wthreads, smem_ws, st>>>(a);
    } else {
      PVB_CUDA(cudaFuncSetAttribute(conv3d_wgrad_f32_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ws));
      conv3d_wgrad_f32_ws_kernel<<<wgrid, wthreads, smem_ws, st>>>(a);
    }
    PVB_LAUNCHED("conv3d_wgrad_f32_ws");
    rc = PVB200_OK;
  } else if (x_is_i16)
    rc = (kts == 1) ? launch_wgrad<true, 1>(a, (int)gx, grid_y, threads, smem, st)
                    : launch_wgrad<true, 3>(a, (int)gx, grid_y, threads, smem, st);
  else
    rc = (kts == 1) ? launch_wgrad<false, 1>(a, (int)gx, grid_y, threads, smem, st)
                    : launch_wgrad<false, 3>(a, (int)gx, grid_y, threads, smem, st);
  if (rc != PVB200_OK) return rc;
  // when the item space is split over grid_y, every y-slice wrote disjoint entries of the same partial rows
  const int n_w = Cout * Cin * 27;
  wgrad_reduce_kernel<<<ceil_div(n_w + Cout, 256), 256, 0, st>>>(a.partial, dw, db, n_w, Cout, (int)gx);
  PVB_LAUNCHED("wgrad_reduce");
  return PVB200_OK;
}

}  // extern "C"
