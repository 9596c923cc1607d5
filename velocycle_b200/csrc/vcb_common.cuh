// Shared device helpers for libvcb (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libvcb is written for sm_100a (B200) only"
#endif

namespace vcb {

// Host side: run an entry point on the device that OWNS the caller's buffers.  Kernels launch in the calling thread's current
// device context; a caller that holds tensors on cuda:1 while device 0 is current would otherwise get
// cudaErrorInvalidResourceHandle (the stream belongs to another device).  The previous device is restored on return.
struct DeviceGuard {
  int prev = -1, dev = -1;
  explicit DeviceGuard(const void* device_ptr) {
    cudaPointerAttributes at;
    if (device_ptr != nullptr && cudaPointerGetAttributes(&at, device_ptr) == cudaSuccess &&
        (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged)) {
      dev = at.device;
      if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
        cudaSetDevice(dev);
      } else {
        prev = -1;
      }
    } else {
      cudaGetLastError();  // a host pointer is an argument error for the entry point to report, not a sticky CUDA error
    }
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr double kLn2d = 0.693147180559945309417232121458;

// ---- MUFU wrappers (one SFU op each) -------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS: UBLKCP / SYNCS); barriers and destinations are passed
// as shared-window addresses (smem_u32) so that hot loops do not redo the generic -> shared conversion ---------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Potentially-suspending probe; the suspend time is capped at ~1 us because the waiting warps also have
// refill duties that depend on OTHER barriers (a long park would delay them).
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(1000u)
      : "memory");
  return ok != 0;
}
// Non-suspending probe (try_wait may park the thread for a HW-defined time; a poller must not).
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a lost transaction must trap instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
// global -> shared::cta bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst),
               "l"(gmem_src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- special functions used off the streaming path ------------------------------------------------
// digamma for x > 0 (float): recurrence up to x >= 6, then the asymptotic series.
__device__ __forceinline__ float digamma_f(float x) {
  float acc = 0.f;
  while (x < 6.f) {
    acc -= 1.f / x;
    x += 1.f;
  }
  const float xi = 1.f / x, x2 = xi * xi;
  return acc + logf(x) - 0.5f * xi - x2 * (1.f / 12.f - x2 * (1.f / 120.f - x2 * (1.f / 252.f)));
}
__device__ __forceinline__ double digamma_d(double x) {
  double acc = 0.0;
  while (x < 10.0) {
    acc -= 1.0 / x;
    x += 1.0;
  }
  const double xi = 1.0 / x, x2 = xi * xi;
  // Bernoulli series: 1/12, 1/120, 1/252, 1/240, 1/132
  return acc + log(x) - 0.5 * xi -
         x2 * (1.0 / 12.0 - x2 * (1.0 / 120.0 - x2 * (1.0 / 252.0 - x2 * (1.0 / 240.0 - x2 * (1.0 / 132.0)))));
}

// lgamma(x) and digamma(x) for x > 0 in one go (double): shift x up to y = x + n >= 10 with the recurrences
// lgamma(x) = lgamma(y) - ln(x (x+1) ... (x+n-1)), psi(x) = psi(y) - sum 1/(x+j), then the Stirling / asymptotic series at y,
// which share ln y.  Truncation error < 2e-14 at y = 10; one fp64 log (two when n > 0) instead of the library lgamma plus a
// separate digamma -- this is the inner loop of the count-spectrum sums (~10^5 evaluations per step).
__device__ __forceinline__ void lgamma_digamma_d(double x, double& lg, double& psi) {
  double prod = 1.0, rec = 0.0;
  while (x < 10.0) {
    prod *= x;
    rec += 1.0 / x;
    x += 1.0;
  }
  const double ly = log(x), xi = 1.0 / x, x2 = xi * xi;
  // Stirling: (y - 1/2) ln y - y + ln(2 pi)/2 + 1/(12 y) - 1/(360 y^3) + 1/(1260 y^5) - 1/(1680 y^7) + 1/(1188 y^9) - 691/(360360 y^11)
  lg = (x - 0.5) * ly - x + 0.918938533204672741780329736406 +
       xi * (1.0 / 12.0 - x2 * (1.0 / 360.0 - x2 * (1.0 / 1260.0 - x2 * (1.0 / 1680.0 - x2 * (1.0 / 1188.0 - x2 * (691.0 / 360360.0))))));
  if (prod != 1.0) lg -= log(prod);
  // psi(y) = ln y - 1/(2y) - 1/(12 y^2) + 1/(120 y^4) - 1/(252 y^6) + 1/(240 y^8) - 1/(132 y^10)
  psi = ly - 0.5 * xi - x2 * (1.0 / 12.0 - x2 * (1.0 / 120.0 - x2 * (1.0 / 252.0 - x2 * (1.0 / 240.0 - x2 * (1.0 / 132.0))))) - rec;
}

// lgamma(r+k) - lgamma(r) - lgamma(k+1) (natural log) and digamma(r+k) - digamma(r), per element, fp32.
// Integer k <= 16 uses the exact finite sums; anything else the library functions.
__device__ __forceinline__ float lgamma_terms_inline(float r, float k, float& psi_diff) {
  if (k == 0.f) {
    psi_diff = 0.f;
    return 0.f;
  }
  if (k <= 16.f && k == floorf(k)) {
    float s = 0.f, p = 0.f;
    for (float j = 0.f; j < k; j += 1.f) {
      const float t = r + j;
      s += lg2_approx(__fdividef(t, j + 1.f));
      p += rcp_approx(t);
    }
    psi_diff = p;
    return s * kLn2;
  }
  psi_diff = digamma_f(r + k) - digamma_f(r);
  return lgammaf(r + k) - lgammaf(r) - lgammaf(k + 1.f);
}

}  // namespace vcb
