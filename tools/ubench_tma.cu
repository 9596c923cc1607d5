// Micro-benchmark (scratch): how fast can ONE CTA per SM stream HBM into a shared-memory ring on B200, as a
// function of the mechanism (1-D bulk TMA copies vs per-thread cp.async 16 B vs plain LDG.128), the copy
// size, the ring depth and the number of CTAs per SM?  Decides the ring design of vcb_stream_kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tma tools/ubench_tma.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}

// ---- A: bulk TMA ring driven by one thread -------------------------------------------------------------------
// The CTA streams `n_stage` stages; stage i = ncopy copies of cb bytes; copy j of stage i reads
// src + ((cta*n_stage + i)*ncopy + j)*stride (stride >= cb: emulates 2 KB pieces of 8 KB rows when stride = 4*cb).
__global__ void k_bulk(const char* src, long long stride, int n_stage, int ns, int ncopy, int cb, int* sink) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* full = (uint64_t*)sm;
  char* buf = (char*)sm + 256;
  const int stage_bytes = ncopy * cb;
  if (threadIdx.x == 0) {
    for (int s = 0; s < ns; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const char* base = src + (long long)blockIdx.x * n_stage * ncopy * stride;
    auto issue = [&](int i) {
      const int s = i % ns;
      mbar_expect_tx(&full[s], stage_bytes);
      for (int j = 0; j < ncopy; ++j) bulk_g2s(buf + (size_t)s * stage_bytes + j * cb, base + ((long long)i * ncopy + j) * stride, cb, &full[s]);
    };
    int issued = 0;
    for (; issued < ns && issued < n_stage; ++issued) issue(issued);
    for (int i = 0; i < n_stage; ++i) {
      while (!mbar_try_wait(&full[i % ns], (i / ns) & 1)) {}
      if (issued < n_stage) issue(issued++);
    }
    if (buf[17] == 123) *sink = 1;
  }
}

// ---- B: cp.async 16 B per thread, commit groups, ns groups in flight ------------------------------------------------
template <int NS>
__global__ void k_cpasync(const char* src, int n_stage, int stage_bytes, int* sink) {
  extern __shared__ __align__(128) unsigned char sm[];
  char* buf = (char*)sm;
  const char* base = src + (long long)blockIdx.x * n_stage * stage_bytes;
  const int per_thread = stage_bytes / (blockDim.x * 16);
  auto issue = [&](int i) {
    const int s = i % NS;
    for (int j = 0; j < per_thread; ++j) {
      const int off = (j * blockDim.x + threadIdx.x) * 16;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(buf + (size_t)s * stage_bytes + off)),
                   "l"(base + (long long)i * stage_bytes + off) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int issued = 0;
  for (; issued < NS && issued < n_stage; ++issued) issue(issued);
  for (int i = 0; i < n_stage; ++i) {
    asm volatile("cp.async.wait_group %0;" ::"n"(NS - 1) : "memory");
    __syncthreads();
    if (issued < n_stage) issue(issued++); else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  if (buf[threadIdx.x] == 123) *sink = 1;
}

// ---- B2: the access pattern of vcb_stream_kernel: lane (grp, q) of warp w loads 16 B at row cs+q (+4), column
// tile*512 + w*32 + 4*grp of two [Nc][ld] matrices; per-thread private ring of NS groups, no barrier at all.
template <int NS, bool REMAP>
__global__ void k_rows(const float* S, const float* U, long long ld, long long n_groups, int n_split, int* sink) {
  extern __shared__ __align__(128) unsigned char sm[];
  float4* buf = (float4*)sm + threadIdx.x;
  const int nthr = blockDim.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane >> 2, q = lane & 3;
  const long long G0 = n_groups * blockIdx.y / n_split, G1 = n_groups * (blockIdx.y + 1) / n_split;
  const int n_stage = (int)(G1 - G0);
  // REMAP: 8 consecutive lanes read one full 128-byte line (row lane/8, chunk lane%8) instead of 4 rows x 32 B
  const int lrow = REMAP ? (lane >> 3) : q, lchunk = REMAP ? (lane & 7) : grp;
  const float* s0 = S + blockIdx.x * 512 + warp * 32 + 4 * lchunk;
  const float* u0 = U + blockIdx.x * 512 + warp * 32 + 4 * lchunk;
  auto issue = [&](int i, int d) {
    if (i < n_stage) {
      const long long c = (G0 + i) * 8 + lrow;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(buf + (d * 4 + 0) * nthr)), "l"(s0 + c * ld) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(buf + (d * 4 + 1) * nthr)), "l"(s0 + (c + 4) * ld) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(buf + (d * 4 + 2) * nthr)), "l"(u0 + c * ld) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(buf + (d * 4 + 3) * nthr)), "l"(u0 + (c + 4) * ld) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int i = 0; i < NS; ++i) issue(i, i);
  float acc = 0.f;
  int d = 0;
  for (int i = 0; i < n_stage; ++i) {
    asm volatile("cp.async.wait_group %0;" ::"n"(NS - 1) : "memory");
    acc += buf[(d * 4) * nthr].x + buf[(d * 4 + 3) * nthr].w;
    issue(i + NS, d);
    if (++d == NS) d = 0;
  }
  if (acc == 123.456f) *sink = 1;
}

// ---- C: plain LDG.128, UNR loads in flight per thread ----------------------------------------------------------
template <int UNR>
__global__ void k_ldg(const float4* src, long long n_per_cta, float* sink) {
  const float4* p = src + (long long)blockIdx.x * n_per_cta;
  float acc = 0.f;
  for (long long i = threadIdx.x; i + (long long)(UNR - 1) * blockDim.x < n_per_cta; i += (long long)UNR * blockDim.x) {
    float4 v[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) v[u] = __ldcs(p + i + (long long)u * blockDim.x);
#pragma unroll
    for (int u = 0; u < UNR; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 123.456f) *sink = acc;
}

static float timeit(void (*launch)(void*), void* ctx) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch(ctx);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch(ctx);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("  !! %s\n", cudaGetErrorString(e));
  return ms;
}

struct BulkCfg { const char* src; long long stride; int grid, n_stage, ns, ncopy, cb; int* sink; };
static void launch_bulk(void* c) {
  BulkCfg* b = (BulkCfg*)c;
  const int smem = 256 + b->ns * b->ncopy * b->cb;
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_bulk<<<b->grid, 32, smem>>>(b->src, b->stride, b->n_stage, b->ns, b->ncopy, b->cb, b->sink);
}
struct CpCfg { const char* src; int grid, n_stage, stage_bytes, ns, nthr; int* sink; };
static void launch_cp(void* c) {
  CpCfg* b = (CpCfg*)c;
  const int smem = b->ns * b->stage_bytes;
#define CP(NS) case NS: cudaFuncSetAttribute(k_cpasync<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
                        k_cpasync<NS><<<b->grid, b->nthr, smem>>>(b->src, b->n_stage, b->stage_bytes, b->sink); break;
  switch (b->ns) { CP(2) CP(3) CP(4) CP(6) CP(8) CP(12) }
}
struct LdgCfg { const float4* src; long long n_per_cta; int grid, nthr, unr; float* sink; };
static void launch_ldg(void* c) {
  LdgCfg* b = (LdgCfg*)c;
  switch (b->unr) {
    case 4: k_ldg<4><<<b->grid, b->nthr>>>(b->src, b->n_per_cta, b->sink); break;
    case 8: k_ldg<8><<<b->grid, b->nthr>>>(b->src, b->n_per_cta, b->sink); break;
    case 16: k_ldg<16><<<b->grid, b->nthr>>>(b->src, b->n_per_cta, b->sink); break;
  }
}

int main() {
  const size_t total = 8ull << 30;  // 8 GiB source
  char* src;
  cudaMalloc(&src, total);
  cudaMemset(src, 1, total);
  int* sink;
  cudaMalloc(&sink, 16);
  const int sms = 148;
  printf("mechanism, config -> GB/s (%% of 6518.6)\n");
  // bulk: per-CTA bytes = n_stage * ncopy * cb; keep total streamed ~6.4 GB (well above L2)
  struct { int ctas_per_sm, ns, ncopy, cb; int stride_mult; } bc[] = {
      {1, 6, 16, 2048, 4}, {1, 6, 1, 32768, 1},
  };
  for (auto& c : bc) {
    const int grid = sms * c.ctas_per_sm;
    const long long stage = (long long)c.ncopy * c.cb;
    const long long footprint_per_stage = (long long)c.ncopy * c.cb * c.stride_mult;
    int n_stage = (int)((6ll << 30) / grid / footprint_per_stage);
    BulkCfg b{src, (long long)c.cb * c.stride_mult, grid, n_stage, c.ns, c.ncopy, c.cb, sink};
    const float ms = timeit(launch_bulk, &b);
    const double gbs = (double)grid * n_stage * stage / ms / 1e6;
    printf("bulk  ctas/sm=%d ring=%2d x (%2d copies x %5d B, stride x%d) = %3lld KB in flight/SM: %7.1f GB/s (%4.1f%%)\n",
           c.ctas_per_sm, c.ns, c.ncopy, c.cb, c.stride_mult, c.ctas_per_sm * c.ns * stage / 1024, gbs, gbs / 65.186);
  }
  struct { int ctas_per_sm, ns, stage_bytes, nthr; } cc[] = {
      {1, 6, 32768, 512}, {1, 12, 16384, 512}, {1, 3, 65536, 512}, {2, 6, 16384, 256}, {2, 6, 16384, 512}, {1, 8, 16384, 1024}, {4, 6, 8192, 256},
  };
  for (auto& c : cc) {
    const int grid = sms * c.ctas_per_sm;
    int n_stage = (int)((6ll << 30) / grid / c.stage_bytes);
    CpCfg b{src, grid, n_stage, c.stage_bytes, c.ns, c.nthr, sink};
    const float ms = timeit(launch_cp, &b);
    const double gbs = (double)grid * n_stage * c.stage_bytes / ms / 1e6;
    printf("cp.async ctas/sm=%d thr=%4d ring=%2d x %5d B = %3d KB in flight/SM: %7.1f GB/s (%4.1f%%)\n", c.ctas_per_sm, c.nthr,
           c.ns, c.stage_bytes, c.ctas_per_sm * c.ns * c.stage_bytes / 1024, gbs, gbs / 65.186);
  }
  {
    const long long ld = 2000, Nc = 400000, n_groups = Nc / 8;
    const float* S = (const float*)src;
    const float* U = (const float*)(src + (4ull << 30));
    for (int ns0 : {6, 106}) {
      int ns = ns0;
      const bool remap = ns > 100; if (remap) ns -= 100;
      const int smem = ns * 4 * 512 * 16;
      float ms = 0;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        dim3 grid(4, 37);
        if (remap) { cudaFuncSetAttribute(k_rows<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k_rows<6, true><<<grid, 512, smem>>>(S, U, ld, n_groups, 37, sink); }
        else { cudaFuncSetAttribute(k_rows<6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k_rows<6, false><<<grid, 512, smem>>>(S, U, ld, n_groups, 37, sink); }
        cudaEventRecord(e1); cudaDeviceSynchronize(); cudaEventElapsedTime(&ms, e0, e1);
      }
      const double gbs = 2.0 * Nc * ld * 4 / ms / 1e6;
      printf("rows pattern (vcb_stream) remap=%d depth=%2d = %3d KB in flight/SM: %.3f ms %7.1f GB/s (%4.1f%%) %s\n", (int)remap, ns, smem / 1024, ms, gbs, gbs / 65.186, cudaGetErrorString(cudaGetLastError()));
    }
  }
  struct { int ctas_per_sm, nthr, unr; } lc[] = {{1, 512, 4}, {1, 512, 8}, {1, 512, 16}, {1, 1024, 8}, {2, 512, 8}, {4, 512, 4}, {4, 256, 8}};
  for (auto& c : lc) {
    const int grid = sms * c.ctas_per_sm;
    const long long n_per_cta = (6ll << 30) / grid / 16;
    LdgCfg b{(const float4*)src, n_per_cta, grid, c.nthr, c.unr, (float*)sink};
    const float ms = timeit(launch_ldg, &b);
    const double gbs = (double)grid * n_per_cta * 16 / ms / 1e6;
    printf("ldg.128 ctas/sm=%d thr=%4d unroll=%2d = %3d KB in flight/SM: %7.1f GB/s (%4.1f%%)\n", c.ctas_per_sm, c.nthr, c.unr,
           c.ctas_per_sm * c.nthr * c.unr * 16 / 1024, gbs, gbs / 65.186);
  }
  return 0;
}
