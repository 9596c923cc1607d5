// The tcgen05 streaming kernel (velocity model, gradients on): the same single pass over the [cells][genes] count
// tiles as vcb_stream.cuh, with every K-term contraction moved from warp-level mma.sync to the 5th-generation
// tensor core (tcgen05.mma kind::tf32, accumulators and the backward A operand in tensor memory).
//
// Why.  The mma.sync kernel spends 59 issue slots per cell.gene; 40 % of them only exist because the operands of
// a warp-level MMA live in registers (fragment loads, hi/lo packing, accumulator moves, per-group ring
// bookkeeping).  Here one elected thread issues all tensor work; the 16 compute warps only
// read finished products from tensor memory, evaluate the negative-binomial terms and write the backward operand
// back to tensor memory.  tools/umma_lab*.cu pin every hardware assumption made below (profiles/r01_umma_lab.txt):
// K-major no-swizzle descriptors, TS mode (A in TMEM), kind::tf32 TRUNCATES its inputs and the fp32 accumulation
// in TMEM TRUNCATES too, the .16x128b thread mapping.
//
// Geometry.  A CTA owns 256 genes (two 128-row MMA blocks) and a contiguous range of 16-cell chunks.
//   forward   D[128 genes x 48] = A_nu[128 x 8] . B[8 x 48]   per block; columns = {y', d, e} x 16 cells,
//             y' = log2e (cf + nu.zeta), d = nu.zeta', e = nu.(omega zeta'');  3 MMAs: hi.hi + lo.hi + hi.lo
//   backward  D_g[128 x 16] += G[128 x 16 cells] . Zg[16 cells x 16],  D_w likewise with W and Zw;
//             G = g_hi | g_lo planes in TMEM (tcgen05.st), Z tiles = [hi of slots 0..7 | lo of slots 0..7]
// Compute warp (mb, ch, lq): block mb, cells 8ch..8ch+7 of the chunk, TMEM lanes 32lq..32lq+31.  Through the
// .16x128b access shape lane (grp, q) owns the rows grp + 8j (j = 0..3) -- which the A operand maps to the 4 ADJACENT
// genes 4grp + j -- and the cells q, q+4: one LDS.128 per cell and matrix brings its counts, and the register pairs the
// tensor memory load returns are gene pairs, so all fp32 math is packed f32x2 without a single register move.
//
// Roles and synchronisation (mbarriers only; one __syncthreads at setup and one at teardown).
//   16 compute warps   wait fwd_full[b] -> tcgen05.ld {y', d, e} -> cp.async.wait (their own count ring: a warp loads exactly
//                      the 32 genes x 8 cells x {S, U} per chunk it consumes) -> add up cell `warp` of chunk ci-2 from the
//                      parked partial sums -> element math -> tcgen05.st the g / w hi | lo planes into G[b] -> park this
//                      chunk's per-cell sums -> arrive g_full[b] -> refill the count slot.  (fwd(ci) is issued behind
//                      bwd(ci-2), so fwd_full[b] also says that G[b] and the parked sums of chunk ci-2 are free / complete.)
//   producer warp      bulk-copies the forward / backward table tiles of chunk ti into 8-deep rings as soon as the MMAs of
//                      chunk ti-8 have released the slot (tcgen05.commit -> tabf_free / tabb_free).
//   MMA warp           converged, one elected lane issues: fwd(0), fwd(1); then per chunk: wait g_full[b] -> bwd(ci) (16
//                      TS MMAs) -> commit bwd_done[b], tabb_free -> fwd(ci+2) (6 SS MMAs) -> commit fwd_full[b], tabf_free.
//   D_fwd and G are double-buffered in TMEM (b = ci & 1); TMEM is full: 192 + 256 + 64 = 512 columns.
//
// Precision.  Inputs: x = hi + lo with hi = rna_tf32(x) for the tables and nu (once per step) and hi = trunc(x) for
// the per-element gradients; the hardware truncates lo to TF32: 2^-21 relative per product.  Accumulation: the
// forward accumulator starts from zero every chunk; the backward accumulators are drained into fp32 global sums
// every kDrain chunks (truncation bias <= 4 kDrain ulp).
#pragma once
#include "vcb_common.cuh"
#include "vcb_stream.cuh"

#include <type_traits>

namespace vcb {
namespace umma {

constexpr int GT = 256;          // genes per CTA
constexpr int NC = 16;           // cells per chunk
constexpr int NCW = 16;          // compute warps
constexpr int NTHR = (NCW + 2) * 32;  // + producer warp + MMA warp
constexpr int NF = 3 * NC;       // forward columns per block
constexpr int kStages = 4;       // count / partial rings (the chunk loop is unrolled by this)
constexpr int kTabStages = 8;    // table rings: a refill costs a DRAM round trip, so they run 6 chunks ahead
constexpr int kDrain = 32;       // chunks between two drains of the backward accumulators

// ---- shared memory map (bytes) ----------------------------------------------------------------------------------------
// counts: warp-private cp.async rings (like vcb_stream.cuh): a warp loads exactly the 32 genes x 8 cells x {S, U} per chunk
// that it consumes, 8 lanes per 128-byte line; row pitch = 32 mod 128 so that the 4 rows of a quarter-warp LDS.128 hit
// distinct banks.  (A producer warp issuing one bulk copy per row was the bottleneck of the first version: per-lane
// addresses turn UBLKCP into a 32-trip waterfall loop.)
constexpr int CNT_PITCH = 128;               // 16-byte chunk c of row r lives at chunk c ^ 2(r & 3): the 4 rows of a quarter-warp
constexpr int CNT_MAT = 8 * CNT_PITCH;       // LDS.128 hit distinct banks without padding
constexpr int CNT_WARP = 2 * CNT_MAT;        // S rows, U rows of one chunk            (2 KB)
constexpr int CNT_STAGE = NCW * CNT_WARP;    // [warp]                                 (32 KB)
constexpr int TABF_BYTES = 2 * 1536;         // B_hi, B_lo: [48 x 8] K-major, LBO 128, SBO 256
constexpr int TABB_BYTES = 2 * 1024;         // Zg, Zw:     [16 x 16] K-major, LBO 128, SBO 512
constexpr int SM_BAR = 0;
constexpr int SM_TMEM = 512;
constexpr int SM_ANU = 1024;                         // [mb][hi/lo] x [128 x 8] K-major, LBO 128, SBO 256 (4 KB each)
constexpr int SM_GENE = SM_ANU + 4 * 4096;           // [gene pair][2] float4: {-r0,-r1,c0,c1}, {gamma0,gamma1,1/beta0,1/beta1}
constexpr int SM_TABF = SM_GENE + (GT / 2) * 32;
constexpr int SM_TABB = SM_TABF + kTabStages * TABF_BYTES;
constexpr int SM_PART = SM_TABB + kTabStages * TABB_BYTES;  // [stage][warp][3][32] floats
constexpr int SM_CNT = SM_PART + kStages * NCW * 96 * 4;
constexpr int SM_TOTAL = SM_CNT + kStages * CNT_STAGE;

// barriers (8 bytes each)
constexpr int B_FWD_FULL = 0, B_G_FULL = 2, B_BWD_DONE = 4, B_PART_FREE = 6, B_TABF_FULL = 10, B_TABF_FREE = 18, B_TABB_FULL = 26,
              B_TABB_FREE = 34, B_COUNT = 42;

// ---- tensor memory map (columns) --------------------------------------------------------------------------------------
constexpr int TM_FWD = 0;     // [buf][mb][qty][16]          192
constexpr int TM_G = 192;     // [buf][mb][plane][16]        256   planes: g_hi, g_lo, w_hi, w_lo
constexpr int TM_DN = 448;    // [mb][g / w][16]              64
constexpr int TM_COLS = 512;

// ---- global table layout (floats per chunk) ----------------------------------------------------------------------------
constexpr int TABF_FLOATS = TABF_BYTES / 4, TABB_FLOATS = TABB_BYTES / 4;

struct Params {
  const float* S;
  const float* U;
  const float* tabF;   // [n_chunks][TABF_FLOATS]
  const float* tabB;   // [n_chunks][TABB_FLOATS]
  const float* omega;  // [n_chunks * NC]
  const float* zero;   // >= GT floats of zeros
  const float* nu;
  const float* dnu;    // [Ng] offsets of the single batch, or null
  const float* shape_inv;
  const float* logbeta;
  const float* gamma;
  float* genepart;     // [2 n_split][rows][ld]
  float* cellpart;     // [n_tiles][Ncp][3]
  long long Nc, Ng, ld, Ncp;
  int n_split, H, rows;
  long long* trace;  // debug: [chunk][8] clock64 stamps of CTA (0,0), chunks 64..127
  int debug;  // timing experiments only (VCB_UMMA_DEBUG): 1 no lo planes backward, 2 no backward MMAs, 4 no lo terms forward, 8 no element math
};

// Instrumentation (clock64 stamps of the chunk pipeline, switches that drop parts of the work for timing experiments) is
// compiled in only with -DVCB_UMMA_INSTRUMENT; the product build has none of it in the loop.
#ifdef VCB_UMMA_INSTRUMENT
#define VCB_TRACE(slot, cidx)                                                                       \
  do {                                                                                              \
    if (P.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (cidx) >= 64 && (cidx) < 128) \
      P.trace[((cidx)-64) * 8 + (slot)] = clock64();                                                \
  } while (0)
#define VCB_DEBUG_BITS (P.debug)
#else
#define VCB_TRACE(slot, cidx) do {} while (0)
#define VCB_DEBUG_BITS 0
#endif

// Barrier wait on the critical path: try_wait without a suspend-time hint (the hardware parks the warp and wakes it when the
// phase completes: a parked warp leaves its issue slots to the service warps it is waiting for).  Bounded: traps.
__device__ __forceinline__ void mbar_wait_hot(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 26)) __trap();
  }
}

// ---- tcgen05 wrappers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// shared-memory matrix descriptor, no swizzle, K-major: LBO = bytes between the 16-byte K chunks of a row,
// SBO = bytes between 8-row groups (mma_sm100_desc.hpp: SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) |
         (1ull << 46);
}
// instruction descriptor: kind::tf32, fp32 accumulate, both operands K-major (mma_sm100_desc.hpp: InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// The MMA wrappers are called by a CONVERGED warp with warp-uniform operands; one elected lane issues.  (Calling them
// from a single-lane branch makes the compiler wrap every instruction in an ELECT / R2UR waterfall loop: ~150 cycles each.)
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
// .16x128b.x2: register 2j+i <-> (lane base + T/4 + 8i, column base + T%4 + 4j)   (tools/umma_lab3.cu)
// Three loads (the quantities y', d, e of one lane half: column strides of `qstride`) and the wait in ONE asm statement, so
// that no use of the results can be scheduled before the wait.
__device__ __forceinline__ void tmem_ld3_16x128b_x2(uint32_t taddr, uint32_t qstride, float2 (&y)[2], float2 (&d)[2], float2 (&e)[2]) {
  uint32_t r[12];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x128b.x2.b32 {%0,%1,%2,%3}, [%12];\n\t"
      "tcgen05.ld.sync.aligned.16x128b.x2.b32 {%4,%5,%6,%7}, [%13];\n\t"
      "tcgen05.ld.sync.aligned.16x128b.x2.b32 {%8,%9,%10,%11}, [%14];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11])
      : "r"(taddr), "r"(taddr + qstride), "r"(taddr + 2 * qstride)
      : "memory");
  y[0] = make_float2(__uint_as_float(r[0]), __uint_as_float(r[1]));
  y[1] = make_float2(__uint_as_float(r[2]), __uint_as_float(r[3]));
  d[0] = make_float2(__uint_as_float(r[4]), __uint_as_float(r[5]));
  d[1] = make_float2(__uint_as_float(r[6]), __uint_as_float(r[7]));
  e[0] = make_float2(__uint_as_float(r[8]), __uint_as_float(r[9]));
  e[1] = make_float2(__uint_as_float(r[10]), __uint_as_float(r[11]));
}
__device__ __forceinline__ void tmem_st_16x128b_x2(uint32_t taddr, float2 c0, float2 c1) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x2.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(__float_as_uint(c0.x)),
               "r"(__float_as_uint(c0.y)), "r"(__float_as_uint(c1.x)), "r"(__float_as_uint(c1.y))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// canonical K-major (no swizzle) byte offset of element (row, k): core matrix = 8 rows x 16 bytes
__host__ __device__ constexpr uint32_t kmajor_off(int row, int k, uint32_t lbo, uint32_t sbo) {
  return (uint32_t)(row >> 3) * sbo + (uint32_t)(row & 7) * 16u + (uint32_t)(k >> 2) * lbo + (uint32_t)(k & 3) * 4u;
}
// MMA row r of a block <-> gene (tile-local, within the block): rows grp + 8j of a 32-row quarter are the adjacent genes 4grp + j
__host__ __device__ constexpr int row_gene(int r) { return (r & ~31) + 4 * (r & 7) + ((r >> 3) & 3); }

__device__ __forceinline__ float tf32_hi_rna(float x) { return __uint_as_float(tf32_rna(x)); }

// ======================================================================================================================
// Per-cell prologue: the B tiles of every 16-cell chunk, already in the shared-memory layout the MMAs read
// (basis: utils.py:400-437; omega: velocity_inference_model.py:365).  One 64-thread block per chunk.
// ======================================================================================================================
struct TableParams {
  const float* phi;
  const float* cf;
  const int32_t* cond_id;
  const float* nu_omega;
  float* tabF;
  float* tabB;
  float* omega;
  float* zero;
  long long Nc, n_chunks;
  int H, Hw;
};

#ifdef VCB_UMMA_KERNELS  // the kernels are compiled in vcb_umma.cu only; vcb.cu needs the structs and constants above
__global__ void __launch_bounds__(64) vcb_umma_tables_kernel(const TableParams P) {
  __shared__ float v[5][NC][8];  // y', d, e, backward zeta, backward omega zeta'
  const long long chunk = blockIdx.x;
  const int t = threadIdx.x;
  if (chunk == 0)
    for (int i = t; i < GT; i += 64) P.zero[i] = 0.f;
  if (t < NC) {
    const long long c = chunk * NC + t;
    const bool valid = c < P.Nc;
    const float phi = valid ? P.phi[c] : 0.f;
    float omega = 0.f;
    if (valid && P.nu_omega != nullptr) {
      const int x = P.cond_id ? P.cond_id[c] : 0;
      const int Kw = 2 * P.Hw + 1;
      const float* nw = P.nu_omega + (long long)x * Kw;
      omega = nw[0];
      for (int n = 1; n <= P.Hw; ++n) {
        float s, co;
        sincosf((float)n * phi, &s, &co);
        omega = fmaf(nw[2 * n - 1], s, omega);
        omega = fmaf(nw[2 * n], co, omega);
      }
    }
    for (int q = 0; q < 5; ++q)
      for (int k = 0; k < 8; ++k) v[q][t][k] = 0.f;
    // slot 0 of the forward operand: A holds 1 there, B the size factor (a padding cell gets y' = -inf-ish: every term vanishes)
    v[0][t][0] = (valid ? (P.cf ? P.cf[c] : 0.f) : -30000.f) * kLog2e;
    v[3][t][0] = 1.f;  // sum_c g        -> d/dnu_0
    v[4][t][7] = 1.f;  // sum_c w        -> d/dgamma (slot 7 is free: 2H <= 6)
    for (int n = 1; n <= P.H; ++n) {
      float s, co;
      const float fn = (float)n;
      sincosf(fn * phi, &s, &co);
      const int ks = 2 * n - 1, kc = 2 * n;
      v[0][t][ks] = s * kLog2e;
      v[0][t][kc] = co * kLog2e;
      v[1][t][ks] = fn * co;
      v[1][t][kc] = -fn * s;
      v[2][t][ks] = omega * (-fn * fn * s);
      v[2][t][kc] = omega * (-fn * fn * co);
      v[3][t][ks] = s;
      v[3][t][kc] = co;
      v[4][t][ks] = omega * (fn * co);
      v[4][t][kc] = omega * (-fn * s);
    }
    P.omega[c] = omega;
  }
  __syncthreads();
  // forward tiles: word o of a [48 x 8] K-major tile (LBO 128, SBO 256) is element (n, k)
  float* tf = P.tabF + chunk * TABF_FLOATS;
  for (int o = t; o < 384; o += 64) {
    const int byte = 4 * o, grp8 = byte >> 8, rem = byte & 255;
    const int n = 8 * grp8 + ((rem & 127) >> 4), k = 4 * (rem >> 7) + ((rem & 15) >> 2);
    const float x = v[n / NC][n % NC][k];
    const float hi = tf32_hi_rna(x);
    tf[o] = hi;
    tf[384 + o] = x - hi;
  }
  // backward tiles: [16 x 16] K-major (LBO 128, SBO 512): row n = slot (n < 8: hi, else lo of slot n-8), k = cell
  float* tb = P.tabB + chunk * TABB_FLOATS;
  for (int o = t; o < 256; o += 64) {
    const int byte = 4 * o, grp8 = byte >> 9, rem = byte & 511;
    const int n = 8 * grp8 + ((rem & 127) >> 4), cell = 4 * (rem >> 7) + ((rem & 15) >> 2);
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const float x = v[3 + w][cell][n & 7];
      const float hi = tf32_hi_rna(x);
      tb[w * 256 + o] = (n < 8) ? hi : (x - hi);
    }
  }
}

#endif  // VCB_UMMA_KERNELS

// ======================================================================================================================
// The streaming kernel
// ======================================================================================================================
#ifdef VCB_UMMA_KERNELS
__global__ void __launch_bounds__(NTHR, 1) vcb_umma_stream_kernel(const Params P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  auto bar = [&](int i) -> uint32_t { return sbase + SM_BAR + 8u * (uint32_t)i; };
  const int tile = blockIdx.x, split = blockIdx.y;
  const long long g_base = (long long)tile * GT;
  const long long remg = P.ld - g_base;
  const int W = (int)(remg < GT ? remg : GT);  // genes of this tile inside the row pitch (multiple of 4)
  const long long n_chunks_all = P.Ncp / NC;
  const long long C0 = (n_chunks_all * split) / P.n_split, C1 = (n_chunks_all * (split + 1)) / P.n_split;
  const int n = (int)(C1 - C0);
  const int K = 2 * P.H + 1;

  // ---- one-time setup -----------------------------------------------------------------------------------------------
  for (int i = tid; i < 2 * 128 * 8; i += NTHR) {  // forward A operand: [1, nu_1..nu_2H, 0..] per row, hi and lo
    const int mb = i >> 10, r = (i >> 3) & 127, k = i & 7;
    const long long g = g_base + mb * 128 + row_gene(r);
    float x = 0.f;
    if (g < P.Ng) x = (k == 0) ? 1.f : (k < K ? P.nu[g * K + k] : 0.f);
    const float hi = tf32_hi_rna(x);
    const uint32_t off = kmajor_off(r, k, 128, 256);
    *reinterpret_cast<float*>(smem + SM_ANU + (mb * 2 + 0) * 4096 + off) = hi;
    *reinterpret_cast<float*>(smem + SM_ANU + (mb * 2 + 1) * 4096 + off) = x - hi;
  }
  for (int i = tid; i < GT / 2; i += NTHR) {  // per-gene parameters, as gene pairs
    float nr[2], c0[2], gm[2], ib[2];
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      const long long g = g_base + 2 * i + o;
      const bool ok = g < P.Ng;
      nr[o] = ok ? -1.0f / P.shape_inv[g] : -1.f;
      c0[o] = ok ? (P.nu[g * K] + logf(P.shape_inv[g]) + (P.dnu ? P.dnu[g] : 0.f)) * kLog2e : -1e30f;
      gm[o] = ok ? P.gamma[g] : 1.f;
      ib[o] = ok ? expf(-P.logbeta[g]) : 1.f;
    }
    const int quarter = (2 * i) >> 5, grp_i = ((2 * i) & 31) >> 2, h_i = i & 1;  // gene 2i = 32 quarter + 4 grp + 2 h
    reinterpret_cast<float4*>(smem + SM_GENE)[((quarter * 2 + h_i) * 2 + 0) * 8 + grp_i] = make_float4(nr[0], nr[1], c0[0], c0[1]);
    reinterpret_cast<float4*>(smem + SM_GENE)[((quarter * 2 + h_i) * 2 + 1) * 8 + grp_i] = make_float4(gm[0], gm[1], ib[0], ib[1]);
  }
  if (tid == 0) {
    for (int i = 0; i < B_COUNT; ++i) {
      const bool many = (i >= B_G_FULL && i < B_G_FULL + 2);
      mbar_init(bar(i), many ? NCW : 1);
    }
    mbar_fence_init();
  }
  fence_async_smem();  // generic-proxy writes (A operand, zeroed count stages) before async-proxy reads / writes
  if (warp == NCW + 1) tmem_alloc(sbase + SM_TMEM, TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM);

  // Per-cell partial sums: every compute warp parks 3 x 32 floats per chunk ([qty][(cc*4 + q)*4 + g']: the 4 lanes that hold the
  // same cell are adjacent).  Two chunks later compute warp w adds up cell w of that chunk: lane (qty, t) takes the float4 of
  // warp t of the cell's half, then a fixed 3-level tree over t (deterministic); 15 instructions per warp and chunk.
  float* const s_part = reinterpret_cast<float*>(smem + SM_PART);
  float* const cellpart_t = P.cellpart + (long long)tile * P.Ncp * 3;
  auto drain_partials = [&](int cj) {  // called by the compute warps only (warp < NCW = NC)
    const int t = lane & 7, qty = lane >> 3;
    float v = 0.f;
    if (qty < 3) {
      const float4 x = *reinterpret_cast<const float4*>(s_part + ((size_t)(cj & (kStages - 1)) * NCW + ((t >> 2) * 8 + (warp >> 3) * 4 + (t & 3))) * 96 +
                                                        qty * 32 + (warp & 7) * 4);
      v = (x.x + x.y) + (x.z + x.w);
    }
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    if (t == 0 && qty < 3) cellpart_t[(C0 + cj) * (NC * 3) + warp * 3 + qty] = v;
  };

  if (warp == NCW) {
    // ================================ producer: the table tiles (8-deep rings: a refill is a DRAM round trip) =============
    for (int ti = 0; ti < n; ++ti) {
      const int st = ti & (kTabStages - 1);
      if (ti >= kTabStages) {
        const uint32_t par_prev = (uint32_t)(((ti / kTabStages) - 1) & 1);
        mbar_wait(bar(B_TABF_FREE + st), par_prev);  // (the parking wait of vcb_common.cuh: this warp is never urgent)
        mbar_wait(bar(B_TABB_FREE + st), par_prev);
      }
      if (lane == 0) {
        const long long chunk = C0 + ti;
        mbar_expect_tx(bar(B_TABF_FULL + st), TABF_BYTES);
        bulk_g2s(sbase + SM_TABF + st * TABF_BYTES, P.tabF + chunk * TABF_FLOATS, TABF_BYTES, bar(B_TABF_FULL + st));
        mbar_expect_tx(bar(B_TABB_FULL + st), TABB_BYTES);
        bulk_g2s(sbase + SM_TABB + st * TABB_BYTES, P.tabB + chunk * TABB_FLOATS, TABB_BYTES, bar(B_TABB_FULL + st));
      }
      __syncwarp();
    }
  } else if (warp == NCW + 1) {
    // ================================ MMA issuer (whole warp, converged; one elected lane issues) ======================
    {
      constexpr uint32_t idesc_f = make_idesc(128, NF), idesc_b = make_idesc(128, 16);
      const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
      auto fwd = [&](int ci) {
        const int b = ci & 1, st = ci & (kTabStages - 1);
        mbar_wait_hot(bar(B_TABF_FULL + st), (uint32_t)((ci / kTabStages) & 1));
        tc_fence_after();
        const uint32_t tb = sbase + SM_TABF + st * TABF_BYTES;
        const uint64_t b_hi = make_desc(tb, 128, 256), b_lo = make_desc(tb + 1536, 128, 256);
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          const uint32_t d = tm + TM_FWD + (b * 2 + mb) * NF;
          const uint64_t a_hi = make_desc(sbase + SM_ANU + (mb * 2 + 0) * 4096, 128, 256);
          const uint64_t a_lo = make_desc(sbase + SM_ANU + (mb * 2 + 1) * 4096, 128, 256);
          mma_ss(d, a_hi, b_hi, idesc_f, 0);
          if (!(VCB_DEBUG_BITS & 4)) {
            mma_ss(d, a_lo, b_hi, idesc_f, 1);
            mma_ss(d, a_hi, b_lo, idesc_f, 1);
          }
        }
        tc_commit_elect(bar(B_FWD_FULL + b));
        tc_commit_elect(bar(B_TABF_FREE + st));
      };
      if (n > 0) fwd(0);
      if (n > 1) fwd(1);
      for (int ci = 0; ci < n; ++ci) {
        const int b = ci & 1, st = ci & (kTabStages - 1);
        mbar_wait_hot(bar(B_TABB_FULL + st), (uint32_t)((ci / kTabStages) & 1));
        mbar_wait_hot(bar(B_G_FULL + b), (uint32_t)((ci >> 1) & 1));
        VCB_TRACE(4, ci);
        tc_fence_after();
        const uint32_t fresh = (ci % kDrain) == 0 ? 0u : 1u;  // the compute warps drained the accumulators before arriving
        const uint32_t tb = sbase + SM_TABB + st * TABB_BYTES;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb)
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            const uint32_t d = tm + TM_DN + (mb * 2 + w) * 16;
#pragma unroll
            for (int pl = 0; pl < 2; ++pl)
#pragma unroll
              for (int ks = 0; ks < 2; ++ks)
                if (!(VCB_DEBUG_BITS & 2) && !(pl == 1 && (VCB_DEBUG_BITS & 1)))
                  mma_ts(d, tm + TM_G + ((b * 2 + mb) * 4 + 2 * w + pl) * 16 + 8 * ks, make_desc(tb + w * 1024 + ks * 256, 128, 512),
                       idesc_b, (pl | ks) ? 1u : fresh);
          }
        tc_commit_elect(bar(B_BWD_DONE + b));
        tc_commit_elect(bar(B_TABB_FREE + st));
        VCB_TRACE(5, ci);
        if (ci + 2 < n) fwd(ci + 2);
        VCB_TRACE(6, ci);
        if ((VCB_DEBUG_BITS & 16) && ci + 2 < n) {  // measurement only: issue -> completion latency of the forward MMAs
          mbar_wait_hot(bar(B_FWD_FULL + b), (uint32_t)(((ci + 2) >> 1) & 1));
          VCB_TRACE(7, ci);
        }
      }
    }
    __syncwarp();
  } else {
    // ================================ compute warps ====================================================================
    const int mb = warp >> 3, ch = (warp >> 2) & 1, lq = warp & 3;
    const int grp = lane >> 2, q = lane & 3;
    const uint32_t tm_lane = tmem + ((uint32_t)(32 * lq) << 16);
    const int gl = mb * 128 + lq * 32 + 4 * grp;  // tile-local index of this lane's 4 adjacent genes
    const float4* s_gene = reinterpret_cast<const float4*>(smem + SM_GENE) + (mb * 4 + lq) * 32 + grp;  // [h][which][grp]
    const float2 zero2 = f2s(0.f), one2 = f2s(1.f), neg1 = f2s(-1.f), eps2 = f2s(1e-5f);
    float2 accAS[2], accLS[2], accAU[2], accLU[2], accGU[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) accAS[h] = accLS[h] = accAU[h] = accLU[h] = accGU[h] = zero2;
    const int rows = P.rows;
    float* gp_slot = P.genepart + ((long long)(split * 2 + ch) * rows) * P.ld;

    // backward accumulators of this warp's rows -> fp32 global sums (thread = TMEM lane = one gene; ch 0 takes D_g, ch 1 D_w)
    auto drain_dnu = [&](const bool first) {
      float v[16];
      tmem_ld_32x32b_x16(tm_lane + TM_DN + (mb * 2 + ch) * 16, v);
      const int gene_l = mb * 128 + lq * 32 + 4 * (lane & 7) + (lane >> 3);
      if (gene_l < W) {
        float* dst = gp_slot + g_base + gene_l;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int row = (k == 7) ? ROW_W : ROW_DNU + k;
          if (k == 7 || k < K) {
            const float x = v[k] + v[8 + k];
            float* p = dst + (long long)row * P.ld;
            *p = first ? x : (*p + x);
          }
        }
      }
    };

    // ---- count ring: this warp's rows of S and U, loaded 8 lanes per 128-byte line (row lane/8 + 4j, 16-byte chunk lane%8);
    // cells past Nc and genes past the row pitch are zero-filled by the copy itself (src-size 0 from a valid address)
    const int l_row = lane >> 3, l_chunk = lane & 7;
    const int l_gene = mb * 128 + lq * 32 + 4 * l_chunk;
    const uint32_t l_sz = l_gene < W ? 16u : 0u;
    const uint32_t s_cnt_w = sbase + SM_CNT + warp * CNT_WARP;
    const uint32_t l_dst = s_cnt_w + l_row * CNT_PITCH + ((l_chunk ^ (2 * l_row)) * 16);
    const float* l_S = P.S + ((C0 * NC + 8 * ch + l_row) * P.ld + g_base + (l_sz ? l_gene : 0));
    const float* l_U = P.U + ((C0 * NC + 8 * ch + l_row) * P.ld + g_base + (l_sz ? l_gene : 0));
    const long long l_step = (long long)NC * P.ld, l_half = 4 * P.ld;
    const long long n_whole = P.Nc / NC - C0;  // chunks [0, n_whole) of this CTA lie entirely inside the matrix
    auto load_counts = [&](int cj, int slot) {  // chunks are loaded in order; one commit group per chunk, empty past the end
      if (cj < n) {
        const uint32_t dst = l_dst + slot * CNT_STAGE;
        if (cj < n_whole) {
          cp_async16(dst, l_S, l_sz);
          cp_async16(dst + 4 * CNT_PITCH, l_S + l_half, l_sz);
          cp_async16(dst + CNT_MAT, l_U, l_sz);
          cp_async16(dst + CNT_MAT + 4 * CNT_PITCH, l_U + l_half, l_sz);
        } else {
          const long long c_row = (C0 + cj) * NC + 8 * ch + l_row;
          const bool ok0 = c_row < P.Nc, ok1 = c_row + 4 < P.Nc;
          cp_async16(dst, ok0 ? l_S : P.S, ok0 ? l_sz : 0u);
          cp_async16(dst + 4 * CNT_PITCH, ok1 ? l_S + l_half : P.S, ok1 ? l_sz : 0u);
          cp_async16(dst + CNT_MAT, ok0 ? l_U : P.S, ok0 ? l_sz : 0u);
          cp_async16(dst + CNT_MAT + 4 * CNT_PITCH, ok1 ? l_U + l_half : P.S, ok1 ? l_sz : 0u);
        }
        l_S += l_step;
        l_U += l_step;
      }
      cp_async_commit();
    };
#pragma unroll
    for (int d = 0; d < kStages; ++d) load_counts(d, d);
    const float* om_ptr = P.omega + C0 * NC + 8 * ch + q;
    float om_next[2] = {0.f, 0.f};
    if (n > 0) {
      om_next[0] = om_ptr[0];
      om_next[1] = om_ptr[4];
    }

    // One chunk.  U = ci mod 4 is a compile-time constant (the loop below is unrolled by the ring depth), so every
    // ring slot, TMEM column and barrier address is an immediate.
    auto chunk_body = [&](auto u_tag, const int ci) {
      constexpr int U = decltype(u_tag)::value;
      constexpr int b = U & 1, st = U;
      const uint32_t par2 = (uint32_t)((ci >> 1) & 1);
      // ---- forward products of this chunk (the commit behind them also covers bwd(ci-2): G[b] is free) -----------
      mbar_wait_hot(bar(B_FWD_FULL + b), par2);
      if (warp == 0) VCB_TRACE(0, ci);
      tc_fence_after();
      float2 Y[2][2], Dd[2][2], E[2][2];  // [gene pair h][cell cc]
      {
        const uint32_t base = tm_lane + TM_FWD + (b * 2) * NF + mb * NF + 8 * ch;
#pragma unroll
        for (int h = 0; h < 2; ++h) tmem_ld3_16x128b_x2(base + ((uint32_t)(16 * h) << 16), NC, Y[h], Dd[h], E[h]);
      }
      // ---- counts ---------------------------------------------------------------------------------------------------
      if (warp == 0) VCB_TRACE(1, ci);
      cp_async_wait<kStages - 1>();  // this lane's loads of chunk ci have landed ...
      __syncwarp();
      if (warp == 0) VCB_TRACE(2, ci);                  // ... and so have the other lanes' (counts never cross a warp)
      const unsigned char* cs = smem + SM_CNT + st * CNT_STAGE + warp * CNT_WARP + ((grp ^ (2 * q)) * 16);
      float4 kS4[2], kU4[2];
      const float om[2] = {om_next[0], om_next[1]};
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        kS4[cc] = *reinterpret_cast<const float4*>(cs + (q + 4 * cc) * CNT_PITCH);
        kU4[cc] = *reinterpret_cast<const float4*>(cs + CNT_MAT + (q + 4 * cc) * CNT_PITCH);
      }
      if (ci + 1 < n) {  // omega of the next chunk (global, L1/L2 resident: written by the table kernel)
        om_next[0] = om_ptr[(ci + 1) * NC];
        om_next[1] = om_ptr[(ci + 1) * NC + 4];
      }
      // ---- the parked partial sums of chunk ci-2 are complete (fwd(ci) was issued behind bwd(ci-2)): this warp adds up its cell
      if (ci >= 2) {
        mbar_wait_hot(bar(B_G_FULL + b), par2 ^ 1u);  // (already complete; acquires the other warps' parked sums)
        drain_partials(ci - 2);
      }
      float2 pcf2[2] = {zero2, zero2}, pphi2[2] = {zero2, zero2}, pom2[2] = {zero2, zero2};
      if (!(VCB_DEBUG_BITS & 8))
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float2 Gh[2], Gl[2], Wh[2], Wl[2];
        // per-gene parameters of gene pair h: two LDS.128 (the 8 grp values of a warp are contiguous: no bank conflicts)
        const float4 ga = s_gene[(2 * h + 0) * 8], gb = s_gene[(2 * h + 1) * 8];
        const float2 nr2_h = f2(ga.x, ga.y), c02_h = f2(ga.z, ga.w), gam2_h = f2(gb.x, gb.y), invb2_h = f2(gb.z, gb.w);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const float2 kS = h ? f2(kS4[cc].z, kS4[cc].w) : f2(kS4[cc].x, kS4[cc].y);
          const float2 kU = h ? f2(kU4[cc].z, kU4[cc].w) : f2(kU4[cc].x, kU4[cc].y);
          const float2 d = Dd[h][cc];
          const float2 y = add2(Y[h][cc], c02_h);
          const float2 u = ex2_2(y);
          const float2 s = add2(u, one2);
          const float2 LS = lg2_2(s);
          accAS[h] = fma2(kS, fma2(LS, neg1, y), accAS[h]);
          accLS[h] = add2(accLS[h], LS);
          const float2 a = fma2(d, f2s(om[cc]), gam2_h);
          const float2 m = add2(f2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f)), eps2);
          const float2 mbeta = mul2(m, invb2_h);
          const float2 uU = mul2(u, mbeta);
          const float2 sU = add2(uU, one2);
          const float2 LU = lg2_2(sU);
          const float2 lmb = lg2_2(mbeta);
          accAU[h] = fma2(kU, fma2(LU, neg1, add2(y, lmb)), accAU[h]);
          accLU[h] = add2(accLU[h], LU);
          const float2 sUm = mul2(sU, m);
          const float2 rc = rcp_2(mul2(s, sUm));  // one MUFU for 1/s and 1/(sU m)
          const float2 inv_s = mul2(rc, sUm), inv_sUm = mul2(rc, s);
          const float2 gS = mul2(fma2(nr2_h, u, kS), inv_s);
          const float2 w0 = mul2(fma2(nr2_h, uU, kU), inv_sUm);
          const float2 gU = mul2(w0, m);
          const float2 w = f2(a.x > 0.f ? w0.x : 0.f, a.y > 0.f ? w0.y : 0.f);
          const float2 g = add2(gS, gU);
          accGU[h] = add2(accGU[h], gU);
          pom2[cc] = fma2(w, d, pom2[cc]);
          pphi2[cc] = fma2(w, E[h][cc], pphi2[cc]);
          pphi2[cc] = fma2(g, d, pphi2[cc]);
          pcf2[cc] = add2(pcf2[cc], g);
          // backward A operand: hi = truncation (what the tensor core keeps anyway), lo = the exact remainder
          Gh[cc] = f2(__uint_as_float(__float_as_uint(g.x) & 0xffffe000u), __uint_as_float(__float_as_uint(g.y) & 0xffffe000u));
          Wh[cc] = f2(__uint_as_float(__float_as_uint(w.x) & 0xffffe000u), __uint_as_float(__float_as_uint(w.y) & 0xffffe000u));
          Gl[cc] = add2(g, f2(-Gh[cc].x, -Gh[cc].y));
          Wl[cc] = add2(w, f2(-Wh[cc].x, -Wh[cc].y));
        }
        const uint32_t ga_t = tm_lane + ((uint32_t)(16 * h) << 16) + TM_G + (b * 2) * 64 + mb * 64 + 8 * ch;
        tmem_st_16x128b_x2(ga_t, Gh[0], Gh[1]);
        tmem_st_16x128b_x2(ga_t + 16, Gl[0], Gl[1]);
        tmem_st_16x128b_x2(ga_t + 32, Wh[0], Wh[1]);
        tmem_st_16x128b_x2(ga_t + 48, Wl[0], Wl[1]);
      }
      // ---- per-cell partial sums: own 4 genes, one shuffle level, park (the rest is summed by drain_partials) -------
      {
        float pcf[2], pphi[2], pom[2];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          pcf[cc] = pcf2[cc].x + pcf2[cc].y;
          pphi[cc] = pphi2[cc].x + pphi2[cc].y;
          pom[cc] = pom2[cc].x + pom2[cc].y;
        }
        const bool up = (lane & 16) != 0;  // after the swap lanes 0-15 hold cell q, lanes 16-31 cell q+4
        float* dst = s_part + ((size_t)st * NCW + warp) * 96 + ((lane >> 4) * 4 + q) * 4 + ((lane >> 2) & 3);
        dst[0] = (up ? pcf[1] : pcf[0]) + __shfl_xor_sync(0xffffffffu, up ? pcf[0] : pcf[1], 16);
        dst[32] = (up ? pphi[1] : pphi[0]) + __shfl_xor_sync(0xffffffffu, up ? pphi[0] : pphi[1], 16);
        dst[64] = (up ? pom[1] : pom[0]) + __shfl_xor_sync(0xffffffffu, up ? pom[0] : pom[1], 16);
      }
      // ---- accumulator drain (every kDrain chunks, before the MMA issuer may restart them) ------------------------------
      if (U == 0 && ci > 0 && (ci % kDrain) == 0) {
        mbar_wait_hot(bar(B_BWD_DONE + 1), (uint32_t)(((ci - 1) >> 1) & 1));  // bwd(ci-1): ci-1 is odd
        tc_fence_after();
        drain_dnu(ci == kDrain);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (warp == 0) VCB_TRACE(3, ci);
      if (warp == 7 && !(VCB_DEBUG_BITS & 16)) VCB_TRACE(7, ci);
      if (lane == 0) mbar_arrive(bar(B_G_FULL + b));
      load_counts(ci + kStages, st);  // refill the slot this warp has just consumed (every lane is past its reads: __syncwarp)
    };
    static_assert(kStages == 4 && kDrain % 4 == 0, "the chunk loop is unrolled by the ring depth");
    for (int c0 = 0; c0 < n; c0 += 4) {
      chunk_body(std::integral_constant<int, 0>{}, c0);
      if (c0 + 1 < n) chunk_body(std::integral_constant<int, 1>{}, c0 + 1);
      if (c0 + 2 < n) chunk_body(std::integral_constant<int, 2>{}, c0 + 2);
      if (c0 + 3 < n) chunk_body(std::integral_constant<int, 3>{}, c0 + 3);
    }
    // ---- tail: everything still in flight ------------------------------------------------------------------------------
    cp_async_wait<0>();
    if (n > 0) {
      if (n > 1) mbar_wait_hot(bar(B_BWD_DONE + ((n - 2) & 1)), (uint32_t)(((n - 2) >> 1) & 1));
      mbar_wait_hot(bar(B_BWD_DONE + ((n - 1) & 1)), (uint32_t)(((n - 1) >> 1) & 1));
      tc_fence_after();
      drain_dnu(n <= kDrain);
      for (int cj = (n > 2 ? n - 2 : 0); cj < n; ++cj) {
        mbar_wait_hot(bar(B_G_FULL + (cj & 1)), (uint32_t)((cj >> 1) & 1));
        drain_partials(cj);
      }
    }
    // per-gene scalar sums: the 4 lanes of a grp hold the same genes (different cells)
    auto lane_sum4 = [&](float2 v) -> float2 {
      v.x += __shfl_xor_sync(0xffffffffu, v.x, 1);
      v.y += __shfl_xor_sync(0xffffffffu, v.y, 1);
      v.x += __shfl_xor_sync(0xffffffffu, v.x, 2);
      v.y += __shfl_xor_sync(0xffffffffu, v.y, 2);
      return v;
    };
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const bool ok = gl + 2 * h < W;
      auto put = [&](int row, float2 v) {
        v = lane_sum4(v);
        if (ok && q == 0) *reinterpret_cast<float2*>(gp_slot + (long long)row * P.ld + g_base + gl + 2 * h) = v;
      };
      put(ROW_AS, accAS[h]);
      put(ROW_LS, accLS[h]);
      put(ROW_AU, accAU[h]);
      put(ROW_LU, accLU[h]);
      put(ROW_GU, accGU[h]);
    }
    if (n == 0) {  // (never with the plan of vcb.cu: n_split <= chunks) the rows the drains would have written
      const int gene_l = mb * 128 + lq * 32 + lane;
      if (gene_l < W)
        for (int k = 0; k < 8; ++k)
          if (k == 7 || k < K) gp_slot[(long long)((k == 7) ? ROW_W : ROW_DNU + k) * P.ld + g_base + gene_l] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NCW + 1) tmem_dealloc(tmem, TM_COLS);
}

#endif  // VCB_UMMA_KERNELS

}  // namespace umma
}  // namespace vcb
