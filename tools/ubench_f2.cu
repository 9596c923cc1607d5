// Micro-benchmark (scratch): issue / pipe cost of the packed fp32 instructions (FFMA2, FMUL2, FADD2) against scalar FFMA, alone
// and mixed with MUFU and ALU work, 16 warps per SM (4 per scheduler) -- the mix of vcb_stream2's element math.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_f2 tools/ubench_f2.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

template <int F2, int F1, int MU, int AL>
__global__ void __launch_bounds__(512, 1) k_mix(float* out, int iters, float seed) {
  float2 p[8];
  float f[8], m[8];
  uint32_t a[8];
  for (int i = 0; i < 8; ++i) {
    p[i] = make_float2(seed + i, seed - i);
    f[i] = seed * 0.25f + i;
    m[i] = seed + i;
    a[i] = threadIdx.x + i;
  }
  const float2 k1 = make_float2(1.0001f, 0.9999f), k0 = make_float2(0.5f, 0.25f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < F2; ++j) p[j & 7] = __ffma2_rn(p[j & 7], k1, k0);
#pragma unroll
    for (int j = 0; j < F1; ++j) f[j & 7] = fmaf(f[j & 7], 1.0001f, 0.5f);
#pragma unroll
    for (int j = 0; j < MU; ++j) {
      float y;
      asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(m[j & 7]));
      m[j & 7] = y;
    }
#pragma unroll
    for (int j = 0; j < AL; ++j) a[j & 7] = (a[j & 7] & 0xffffe000u) ^ (a[(j + 1) & 7] >> 3);
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += p[i].x + p[i].y + f[i] + m[i] + __uint_as_float(a[i] & 0x3fffffffu);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int F2, int F1, int MU, int AL>
static void run(const char* name, float* out, int sms, float ghz) {
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k_mix<F2, F1, MU, AL><<<sms, 512>>>(out, 100, 1.f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k_mix<F2, F1, MU, AL><<<sms, 512>>>(out, iters, 1.f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double per_iter = ms * 1e-3 * ghz * 1e9 / iters / 4.0;  // cycles per iteration per warp-slot of a scheduler (4 warps each)
  printf("%-44s %7.1f clk per warp-iteration per SMSP  (sum of instructions %d)\n", name, per_iter, F2 + F1 + MU + 2 * AL);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const float ghz = khz * 1e-6f;
  printf("%s, %d SMs, %.3f GHz nominal; 4 warps per scheduler; clk = cycles of one scheduler per iteration of ONE of its warps\n", p.name,
         p.multiProcessorCount, ghz);
  float* out;
  cudaMalloc(&out, 512 * p.multiProcessorCount * 4);
  const int sms = p.multiProcessorCount;
  run<32, 0, 0, 0>("ffma2 x32", out, sms, ghz);
  run<0, 32, 0, 0>("ffma x32", out, sms, ghz);
  run<0, 64, 0, 0>("ffma x64", out, sms, ghz);
  run<0, 0, 8, 0>("mufu x8", out, sms, ghz);
  run<0, 0, 0, 16>("alu x32 (lop3 + shf)", out, sms, ghz);
  run<32, 0, 8, 0>("ffma2 x32 + mufu x8", out, sms, ghz);
  run<0, 64, 8, 0>("ffma x64 + mufu x8", out, sms, ghz);
  run<32, 16, 8, 0>("ffma2 x32 + ffma x16 + mufu x8", out, sms, ghz);
  run<32, 16, 8, 8>("ffma2 x32 + ffma x16 + mufu x8 + alu x16", out, sms, ghz);
  run<26, 13, 10, 6>("stream2 element mix per pair (26 f2, 13 f, 10 mufu, 12 alu)", out, sms, ghz);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
