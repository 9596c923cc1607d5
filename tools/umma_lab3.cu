// Third lab: the thread <-> (lane, column) mapping of tcgen05.ld / tcgen05.st .16x128b (and .16x256b).
#define main lab1_main
#include "umma_lab.cu"
#undef main

struct Lab3Out {
  float a[128][4];   // 16x128b.x2 at lane offset 0 : per thread 4 regs
  float b[128][4];   // 16x128b.x2 at lane offset 16
  float c[128][4];   // 16x256b.x1 at lane offset 0
  float d[128][16];  // 32x32b read-back of what st.16x128b.x2 wrote (values = 1000*tid + reg)
};

__global__ void __launch_bounds__(128, 1) lab3_kernel(Lab3Out* out) {
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
  // value at (lane L, column c) = 100*L + c
  for (int c0 = 0; c0 < 16; c0 += 8) {
    float v[8];
    for (int i = 0; i < 8; ++i) v[i] = 100.f * (32 * warp + lane) + c0 + i;
    tmem_st8(tmem + lane_base + c0, v);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(tmem + lane_base));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 4; ++i) out->a[tid][i] = __uint_as_float(r[i]);
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(tmem + lane_base + (16u << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 4; ++i) out->b[tid][i] = __uint_as_float(r[i]);
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(tmem + lane_base));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 4; ++i) out->c[tid][i] = __uint_as_float(r[i]);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // st.16x128b.x2 at columns 32.. (both lane halves), then read everything back with 32x32b
  for (int h = 0; h < 2; ++h) {
    uint32_t w[4];
    for (int i = 0; i < 4; ++i) w[i] = __float_as_uint(1000.f * lane + 10.f * h + i);
    asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1,%2,%3,%4};" ::"r"(tmem + lane_base + ((uint32_t)(16 * h) << 16) + 32), "r"(w[0]),
                 "r"(w[1]), "r"(w[2]), "r"(w[3])
                 : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  float v[16];
  tmem_ld16(tmem + lane_base + 32, v);
  for (int i = 0; i < 16; ++i) out->d[tid][i] = v[i];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  Lab3Out* dout;
  CK(cudaMalloc(&dout, sizeof(Lab3Out)));
  CK(cudaMemset(dout, 0, sizeof(Lab3Out)));
  lab3_kernel<<<1, 128>>>(dout);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  Lab3Out* o = new Lab3Out;
  CK(cudaMemcpy(o, dout, sizeof(Lab3Out), cudaMemcpyDeviceToHost));
  printf("value = 100*lane + column.  thread: 16x128b.x2 @lanes 0-15 | @lanes 16-31 | 16x256b.x1\n");
  for (int t = 0; t < 40; ++t) {
    if (t >= 12 && t < 32) continue;
    printf("t%3d: %6.0f %6.0f %6.0f %6.0f | %6.0f %6.0f %6.0f %6.0f | %6.0f %6.0f %6.0f %6.0f\n", t, o->a[t][0], o->a[t][1], o->a[t][2], o->a[t][3],
           o->b[t][0], o->b[t][1], o->b[t][2], o->b[t][3], o->c[t][0], o->c[t][1], o->c[t][2], o->c[t][3]);
  }
  // check the hypothesis: reg 2j+i of 16x128b.x2 = (lane base + T/4 + 8i, column T%4 + 4j)
  int bad = 0;
  for (int t = 0; t < 128; ++t) {
    const int w = t / 32, T = t % 32;
    for (int j = 0; j < 2; ++j)
      for (int i = 0; i < 2; ++i) {
        if (o->a[t][2 * j + i] != 100.f * (32 * w + T / 4 + 8 * i) + (T % 4 + 4 * j)) ++bad;
        if (o->b[t][2 * j + i] != 100.f * (32 * w + 16 + T / 4 + 8 * i) + (T % 4 + 4 * j)) ++bad;
      }
  }
  printf("hypothesis reg[2j+i] = (lane T/4 + 8i, col T%%4 + 4j): %d mismatches\n", bad);
  printf("st.16x128b.x2 read back with 32x32b (value = 1000*thread + 10*half + reg), lanes 0,1,8,9,16,17,24 of warp 0, columns 0..7:\n");
  const int ls[7] = {0, 1, 8, 9, 16, 17, 24};
  for (int k = 0; k < 7; ++k) {
    printf("lane %2d:", ls[k]);
    for (int c = 0; c < 8; ++c) printf(" %7.0f", o->d[ls[k]][c]);
    printf("\n");
  }
  return 0;
}
