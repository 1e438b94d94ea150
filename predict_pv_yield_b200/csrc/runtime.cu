// runtime.cu -- library state of libpvb200: last error, launch counter, device queries.
#include <atomic>
#include <string.h>

#include "common.cuh"

namespace pvb {

static thread_local char g_err[1024] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static std::atomic<int> g_reserved_sms{0};
int g_dynamic_tiles = 0;  // pvb200_set_dynamic_tiles: persistent kernels that support it claim their work in chunks

// SMs the persistent kernels size their grids for: the device's SM count minus the SMs reserved for concurrently
// running collectives (pvb200_reserve_sms).  A persistent grid of one CTA per SM with a static work split loses up to
// half its speed when a communication kernel already occupies some SMs: the displaced CTAs only start when another
// CTA of the same grid has finished.  Sizing the grid for the SMs that are actually free keeps the split balanced.
int sm_count() {
  static int cached = -1;
  if (cached <= 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached = n;
  }
  const int r = g_reserved_sms.load(std::memory_order_relaxed);
  const int n = cached - r;
  return n < 8 ? (cached < 8 ? cached : 8) : n;
}

}  // namespace pvb

extern "C" {

int pvb200_abi_version(void) { return PVB200_ABI_VERSION; }

const char* pvb200_last_error(void) { return pvb::g_err; }

unsigned long long pvb200_launch_count(void) { return pvb::g_launches.load(); }

void pvb200_reset_launch_count(void) { pvb::g_launches.store(0); }

/* reserve `n` SMs for kernels of other libraries that run concurrently (NCCL collectives under data parallelism):
 * every persistent kernel launched afterwards uses (SM count - n) CTAs.  n = 0 restores the default.  Returns the old value. */
int pvb200_reserve_sms(int n) { return pvb::g_reserved_sms.exchange(n < 0 ? 0 : n); }

int pvb200_set_dynamic_tiles(int on) {
  const int old = pvb::g_dynamic_tiles;
  pvb::g_dynamic_tiles = on ? 1 : 0;
  return old;
}

int pvb200_sm_count(void) {
  int n = pvb::sm_count();
  if (n <= 0) pvb::set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
  return n;
}

}  // extern "C"

// ---- FP32 FMA peak probe -------------------------------------------------------------------------------
// MEASURED_PEAKS.json holds HBM and bf16 tensor peaks only; the fp32-mode convolutions are bound by the
// FP32 FMA pipe, so bench.py measures that ceiling live with this kernel (16 independent FMA chains per
// thread, 8 CTAs x 256 threads per SM) and reports the conv kernels against it.
namespace pvb {
__global__ void __launch_bounds__(256) fma_probe_kernel(float* sink, int iters, float a, float b) {
  float c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = static_cast<float>(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fmaf(c[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 123.456f) sink[0] = s;  // never true in practice; keeps the chains alive
}

// the same chains issued as the packed fma.rn.f32x2 of sm_100 (SASS FFMA2): 16 independent pair chains per thread
__global__ void __launch_bounds__(256) fma2_probe_kernel(float* sink, int iters, float a, float b) {
  float2 c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = make_float2(static_cast<float>(threadIdx.x + i), static_cast<float>(threadIdx.x - i));
  const float2 a2 = make_float2(a, a * 0.5f), b2 = make_float2(b, b * 0.5f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = __ffma2_rn(c[i], a2, b2);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i].x + c[i].y;
  if (s == 123.456f) sink[0] = s;
}
}  // namespace pvb

extern "C" int pvb200_probe_fp32_fma(float* sink, int iters, double* flops_out, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(sink && flops_out && iters > 0, "probe_fp32_fma: bad argument");
  const int sms = sm_count();
  PVB_REQUIRE(sms > 0, "probe_fp32_fma: no CUDA device");
  const int grid = sms * 8;
  fma_probe_kernel<<<grid, 256, 0, as_stream(stream)>>>(sink, iters, 0.999f, 0.001f);
  PVB_LAUNCHED("fma_probe");
  *flops_out = 2.0 * 16.0 * static_cast<double>(iters) * 256.0 * grid;
  return PVB200_OK;
}

extern "C" int pvb200_probe_fp32_fma2(float* sink, int iters, double* flops_out, pvb200_stream_t stream) {
  using namespace pvb;
  PVB_REQUIRE(sink && flops_out && iters > 0, "probe_fp32_fma2: bad argument");
  const int sms = sm_count();
  PVB_REQUIRE(sms > 0, "probe_fp32_fma2: no CUDA device");
  const int grid = sms * 8;
  fma2_probe_kernel<<<grid, 256, 0, as_stream(stream)>>>(sink, iters, 0.999f, 0.001f);
  PVB_LAUNCHED("fma2_probe");
  *flops_out = 2.0 * 32.0 * static_cast<double>(iters) * 256.0 * grid;
  return PVB200_OK;
}
