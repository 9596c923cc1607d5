// Second lab: (a) where M=64 puts D, (b) MN-major A for kind::tf32 (both readings of LBO / SBO), (c) MMA throughput
// with independent accumulators.  Build like umma_lab.cu (shares its helpers by inclusion).
#define main lab1_main
#include "umma_lab.cu"
#undef main

struct Lab2Out {
  float d[6][128 * 16];
  long long cyc[16];
};

// MN-major (no swizzle) byte offset: 4 MN elements contiguous (16 B), 8 K rows at 16 B stride, then the two strides
__host__ __device__ inline uint32_t mnmajor_off(int mn, int k, uint32_t mn_stride, uint32_t k_stride) {
  return (uint32_t)(mn >> 2) * mn_stride + (uint32_t)(k >> 3) * k_stride + (uint32_t)(k & 7) * 16u + (uint32_t)(mn & 3) * 4u;
}

constexpr int S2_A = 1024;             // A as K-major [128 x 16]: LBO 128, SBO 512          (8 KB)
constexpr int S2_AT = S2_A + 8192;     // the same matrix MN-major: mn_stride 256 (2 k groups x 128), k_stride 128  (8 KB)
constexpr int S2_B = S2_AT + 8192;     // B K-major [16 x 16]: LBO 128, SBO 512             (1 KB)
constexpr int S2_TOTAL = 65536;

__global__ void __launch_bounds__(128, 1) lab2_kernel(const float* A, const float* B, Lab2Out* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 64);
  for (int i = tid; i < 128 * 16; i += 128) {
    const int r = i / 16, k = i % 16;
    *reinterpret_cast<float*>(smem + S2_A + kmajor_off(r, k, 128, 512)) = A[i];
    *reinterpret_cast<float*>(smem + S2_AT + mnmajor_off(r, k, 256, 128)) = A[i];
  }
  for (int i = tid; i < 16 * 16; i += 128) {
    const int r = i / 16, k = i % 16;
    *reinterpret_cast<float*>(smem + S2_B + kmajor_off(r, k, 128, 512)) = B[i];
  }
  if (tid == 0) mbar_init(bar, 1);
  fence_async_smem();
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
  uint32_t phase = 0;
  // zero 6 x 16 columns
  {
    float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int c = 0; c < 96; c += 8) tmem_st8(tmem + lane_base + c, z);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (tid == 0) {
    const uint64_t bd0 = make_desc(sbase + S2_B, 128, 512), bd1 = make_desc(sbase + S2_B + 256, 128, 512);
    // 0: M=128 K-major (reference)
    mma_ss(tmem + 0, make_desc(sbase + S2_A, 128, 512), bd0, make_idesc(128, 16, 0, 0), 0);
    mma_ss(tmem + 0, make_desc(sbase + S2_A + 256, 128, 512), bd1, make_idesc(128, 16, 0, 0), 1);
    // 1: M=64 K-major
    mma_ss(tmem + 16, make_desc(sbase + S2_A, 128, 512), bd0, make_idesc(64, 16, 0, 0), 0);
    mma_ss(tmem + 16, make_desc(sbase + S2_A + 256, 128, 512), bd1, make_idesc(64, 16, 0, 0), 1);
    // 2: M=128 MN-major A, LBO = k-group stride (128), SBO = mn-chunk stride (256)
    mma_ss(tmem + 32, make_desc(sbase + S2_AT, 128, 256), bd0, make_idesc(128, 16, 1, 0), 0);
    mma_ss(tmem + 32, make_desc(sbase + S2_AT + 128, 128, 256), bd1, make_idesc(128, 16, 1, 0), 1);
    // 3: M=128 MN-major A, LBO = mn-chunk stride (256), SBO = k-group stride (128)
    mma_ss(tmem + 48, make_desc(sbase + S2_AT, 256, 128), bd0, make_idesc(128, 16, 1, 0), 0);
    mma_ss(tmem + 48, make_desc(sbase + S2_AT + 128, 256, 128), bd1, make_idesc(128, 16, 1, 0), 1);
    // 4 / 5: M=64 MN-major A, both readings
    mma_ss(tmem + 64, make_desc(sbase + S2_AT, 128, 256), bd0, make_idesc(64, 16, 1, 0), 0);
    mma_ss(tmem + 64, make_desc(sbase + S2_AT + 128, 128, 256), bd1, make_idesc(64, 16, 1, 0), 1);
    mma_ss(tmem + 80, make_desc(sbase + S2_AT, 256, 128), bd0, make_idesc(64, 16, 1, 0), 0);
    mma_ss(tmem + 80, make_desc(sbase + S2_AT + 128, 256, 128), bd1, make_idesc(64, 16, 1, 0), 1);
    tc_commit(bar);
  }
  mbar_wait(bar, phase);
  phase ^= 1;
  tc_fence_after();
  for (int t = 0; t < 6; ++t) {
    float v[16];
    tmem_ld16(tmem + lane_base + 16 * t, v);
    for (int i = 0; i < 16; ++i) out->d[t][(32 * warp + lane) * 16 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();

  // ---- throughput: R MMAs over NACC independent accumulators --------------------------------------------------------
  auto timed = [&](int slot, int R, int nacc, int M, int N, bool ts) {
    tc_fence_after();
    long long t0 = 0;
    if (tid == 0) {
      const uint32_t idesc = make_idesc(M, N, 0, 0);
      const uint64_t ad = make_desc(sbase + S2_A, 128, 512), bd = make_desc(sbase + S2_B, 128, 512);
      t0 = clock64();
      int a = 0;
      for (int r = 0; r < R; ++r) {
        if (ts)
          mma_ts(tmem + 128 + a * N, tmem + 0, bd, idesc, 1);
        else
          mma_ss(tmem + 128 + a * N, ad, bd, idesc, 1);
        if (++a == nacc) a = 0;
      }
      tc_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    if (tid == 0) out->cyc[slot] = clock64() - t0;
    tc_fence_before();
    __syncthreads();
  };
  timed(0, 32, 1, 128, 16, false);
  timed(1, 128, 1, 128, 16, false);
  timed(2, 128, 2, 128, 16, false);
  timed(3, 128, 4, 128, 16, false);
  timed(4, 128, 8, 128, 16, false);
  timed(5, 128, 8, 128, 16, true);
  timed(6, 128, 4, 128, 64, false);
  timed(7, 128, 1, 128, 64, false);
  timed(8, 128, 8, 64, 16, false);
  timed(9, 128, 1, 128, 256, false);
  timed(10, 128, 8, 128, 32, false);
  timed(11, 32, 8, 128, 16, false);

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  srand(2);
  std::vector<float> A(128 * 16), B(16 * 16);
  for (auto& x : A) x = frand();
  for (auto& x : B) x = frand();
  float *dA, *dB;
  Lab2Out* dout;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dout, sizeof(Lab2Out)));
  CK(cudaMemset(dout, 0, sizeof(Lab2Out)));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(lab2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_TOTAL));
  lab2_kernel<<<1, 128, S2_TOTAL>>>(dA, dB, dout);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  Lab2Out* o = new Lab2Out;
  CK(cudaMemcpy(o, dout, sizeof(Lab2Out), cudaMemcpyDeviceToHost));
  std::vector<double> X(128 * 16);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 16; ++n) {
      double s = 0;
      for (int k = 0; k < 16; ++k) s += (double)tf32_trunc(A[m * 16 + k]) * tf32_trunc(B[n * 16 + k]);
      X[m * 16 + n] = s;
    }
  const char* nm[6] = {"M128 K-major", "M64 K-major", "M128 MN-major lbo=k sbo=mn", "M128 MN-major lbo=mn sbo=k", "M64 MN-major lbo=k sbo=mn",
                       "M64 MN-major lbo=mn sbo=k"};
  for (int t = 0; t < 6; ++t) {
    const int M = (t == 1 || t >= 4) ? 64 : 128;
    double e_id = 0, e_64 = 0;
    int nz = 0;
    for (int l = 0; l < 128; ++l) {
      bool z = true;
      for (int n = 0; n < 16; ++n) z &= o->d[t][l * 16 + n] == 0.f;
      nz += !z;
    }
    for (int r = 0; r < M; ++r)
      for (int n = 0; n < 16; ++n) {
        e_id = fmax(e_id, fabs(o->d[t][r * 16 + n] - X[r * 16 + n]));
        e_64 = fmax(e_64, fabs(o->d[t][((r / 16) * 32 + r % 16) * 16 + n] - X[r * 16 + n]));
      }
    printf("%-30s: non-zero lanes %3d ; max err rows at lanes r: %.3e ; rows at lanes (r/16)*32+r%%16: %.3e\n", nm[t], nz, e_id, e_64);
    if (t == 1 || t >= 4) {
      printf("    non-zero lanes:");
      for (int l = 0; l < 128; ++l) {
        bool z = true;
        for (int n = 0; n < 16; ++n) z &= o->d[t][l * 16 + n] == 0.f;
        if (!z) printf(" %d", l);
      }
      printf("\n");
    }
  }
  const char* tn[12] = {"SS M128 N16 1 acc R=32", "SS M128 N16 1 acc R=128", "SS M128 N16 2 acc", "SS M128 N16 4 acc", "SS M128 N16 8 acc",
                        "TS M128 N16 8 acc", "SS M128 N64 4 acc", "SS M128 N64 1 acc", "SS M64 N16 8 acc", "SS M128 N256 1 acc",
                        "SS M128 N32 8 acc", "SS M128 N16 8 acc R=32"};
  const int Rs[12] = {32, 128, 128, 128, 128, 128, 128, 128, 128, 128, 128, 32};
  for (int i = 0; i < 12; ++i) printf("%-26s: %7lld cycles total, %6.1f per MMA\n", tn[i], o->cyc[i], (double)o->cyc[i] / Rs[i]);
  return 0;
}
