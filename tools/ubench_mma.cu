// Micro-benchmark (scratch): throughput of the legacy warp-level mma.sync.m16n8k8 TF32 path, MUFU and FFMA2
// on sm_100a, alone and mixed, with 16 warps/SM -- the numbers that bound vcb_stream_kernel's pipes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_mma tools/ubench_mma.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NCHAIN, int MUFU_PER, int FMA_PER>
__global__ void __launch_bounds__(512, 1) k_mix(float* out, int iters, float seed) {
  float c[NCHAIN ? NCHAIN : 1][4];
  uint32_t a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(seed + threadIdx.x * 0.001f + i);
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(seed * 0.5f + i);
  for (int n = 0; n < NCHAIN; ++n)
    for (int i = 0; i < 4; ++i) c[n][i] = 0.f;
  float m[8], f[8];
  for (int i = 0; i < 8; ++i) {
    m[i] = seed + i;
    f[i] = seed * 0.25f + i;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int n = 0; n < NCHAIN; ++n) mma_tf32(c[n], a, b);
#pragma unroll
    for (int j = 0; j < MUFU_PER; ++j) {
      float y;
      asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(m[j & 7]));
      m[j & 7] = y + 3.f;
    }
#pragma unroll
    for (int j = 0; j < FMA_PER; ++j) f[j & 7] = fmaf(f[j & 7], 1.0001f, 0.5f);
  }
  float s = 0.f;
  for (int n = 0; n < NCHAIN; ++n)
    for (int i = 0; i < 4; ++i) s += c[n][i];
  for (int i = 0; i < 8; ++i) s += m[i] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NCHAIN, int MUFU_PER, int FMA_PER>
static void run(const char* name, float* out, int sms, float ghz) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_mix<NCHAIN, MUFU_PER, FMA_PER><<<sms, 512>>>(out, 100, 1.f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_mix<NCHAIN, MUFU_PER, FMA_PER><<<sms, 512>>>(out, iters, 1.f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double clk = ms * 1e-3 * ghz * 1e9;  // cycles per SM
  const double per_iter = clk / iters;       // cycles per iteration of all 16 warps (4 per SMSP)
  printf("%-28s %8.3f ms  %7.1f clk/iter/SM", name, ms, per_iter);
  if (NCHAIN) printf("  mma: %.2f clk per warp-MMA per SMSP, %.0f TF32 TFLOP/s", per_iter / ((NCHAIN ? NCHAIN : 1) * 4.0),
                     2.0 * 1024 * NCHAIN * 16.0 * iters * sms / (ms * 1e-3) / 1e12);
  if (MUFU_PER) printf("  mufu: %.2f clk per warp-op per SMSP", per_iter / ((MUFU_PER ? MUFU_PER : 1) * 4.0));
  if (FMA_PER) printf("  ffma: %.2f clk per warp-op per SMSP", per_iter / ((FMA_PER ? FMA_PER : 1) * 4.0));
  printf("\n");
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const float ghz = khz * 1e-6f;
  printf("%s, %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
  float* out;
  cudaMalloc(&out, 512 * p.multiProcessorCount * 4);
  const int sms = p.multiProcessorCount;
  run<8, 0, 0>("mma x8 chains", out, sms, ghz);
  run<4, 0, 0>("mma x4 chains", out, sms, ghz);
  run<1, 0, 0>("mma x1 chain (latency)", out, sms, ghz);
  run<0, 8, 0>("mufu x8", out, sms, ghz);
  run<0, 0, 8>("ffma x8", out, sms, ghz);
  run<0, 0, 32>("ffma x32", out, sms, ghz);
  run<8, 8, 0>("mma x8 + mufu x8", out, sms, ghz);
  run<8, 0, 32>("mma x8 + ffma x32", out, sms, ghz);
  run<8, 8, 32>("mma x8 + mufu x8 + ffma x32", out, sms, ghz);
  run<15, 20, 60>("mma15 + mufu20 + ffma60", out, sms, ghz);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
