// libvcb: the step's single exchange -- a SUM all-reduce of the flat gene-level buffer (~90 KB at 2k genes) across the GPUs
// of one box -- as ONE kernel over NVLink peer memory.
//
// NCCL needs ~100 us for this payload at 8 ranks (latency, not bandwidth; bench breakdown_ms.allreduce), 6 % of a strong-
// scaling step at 250k cells per GPU.  Every rank owns a receive buffer of 2 sets x world slots that its peers have mapped
// (CUDA IPC); a call
//   1. pushes the rank's payload into slot[rank] of every peer's buffer (128-bit stores over NVLink / NVSwitch),
//   2. publishes it: __threadfence_system + a release store of the call's epoch into the peer's flag word,
//   3. waits until its own flag words show the epoch for every rank (acquire loads, bounded spin),
//   4. adds the slots in rank order 0..world-1 -- the same order on every rank, so the result is reproducible AND bitwise
//      identical on all ranks -- and writes the sum back over the payload.
// The two sets alternate with the epoch (kept in device memory, so CUDA-graph replays advance it): a rank can only run one
// call ahead of the slowest rank, and then it writes the other set.
#include "vcb.h"

#include <cuda_runtime.h>
#include <stdint.h>

#include "vcb_common.cuh"

namespace vcb {

constexpr int kCommThreads = 1024;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kCommThreads) vcb_allreduce_oneshot_kernel(const vcb_comm_t C, float* __restrict__ data,
                                                                               long long n) {
  const int tid = threadIdx.x, world = C.world, rank = C.rank;
  const uint32_t e = *C.epoch;
  const long long set = (long long)(e & 1u) * world;
  const long long n4 = n >> 2;  // the payload is padded to a multiple of 4 floats by the caller
  // 1. push
  const float4* src = reinterpret_cast<const float4*>(data);
  for (int k = 0; k < world; ++k) {
    const int peer = (rank + k) % world;  // staggered start: not everybody hammers rank 0 first
    float4* dst = reinterpret_cast<float4*>(C.slots[peer] + (set + rank) * C.slot_floats);
    for (long long i = tid; i < n4; i += kCommThreads) dst[i] = src[i];
  }
  // 2. publish
  __threadfence_system();
  __syncthreads();
  if (tid < world) st_release_sys(C.flags[tid] + set + rank, e + 1u);
  // 3. wait for everybody's payload in MY buffer
  if (tid < world) {
    const uint32_t* f = C.flags[rank] + set + tid;
    unsigned long long spins = 0;
    while (ld_acquire_sys(f) != e + 1u) {
      if (++spins > (1ull << 31)) __trap();  // a lost peer must not hang the GPU forever
    }
  }
  __syncthreads();
  // 4. reduce in rank order
  const float* mine = C.slots[rank] + set * C.slot_floats;
  float4* out = reinterpret_cast<float4*>(data);
  for (long long i = tid; i < n4; i += kCommThreads) {
    float4 s = reinterpret_cast<const float4*>(mine)[i];
    for (int r = 1; r < world; ++r) {
      const float4 v = reinterpret_cast<const float4*>(mine + (long long)r * C.slot_floats)[i];
      s.x += v.x;
      s.y += v.y;
      s.z += v.z;
      s.w += v.w;
    }
    out[i] = s;
  }
  __syncthreads();
  if (tid == 0) *C.epoch = e + 1u;
}

}  // namespace vcb

extern "C" int vcb_allreduce_sum(const vcb_comm_t* c, float* data, int64_t n, void* stream) {
  if (c == nullptr || data == nullptr || c->epoch == nullptr) return VCB_ERR_NULL;
  if (c->world < 1 || c->world > VCB_MAX_RANKS || c->rank < 0 || c->rank >= c->world) return VCB_ERR_SIZE;
  if (n < 0 || (n & 3) != 0 || n > c->slot_floats || (c->slot_floats & 3) != 0) return VCB_ERR_SIZE;
  if ((((uintptr_t)data) & 15) != 0) return VCB_ERR_ALIGN;
  for (int r = 0; r < c->world; ++r)
    if (c->slots[r] == nullptr || c->flags[r] == nullptr) return VCB_ERR_NULL;
  if (c->world == 1 || n == 0) return VCB_OK;
  vcb::DeviceGuard guard(data);
  vcb::vcb_allreduce_oneshot_kernel<<<1, vcb::kCommThreads, 0, (cudaStream_t)stream>>>(*c, data, (long long)n);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? VCB_OK : (int)e;
}
