// The streaming kernel: one pass over the [cells][genes] count tiles of S (and U) that produces the
// per-gene log-prob partial sums AND every gradient partial sum.
//
// Where the arithmetic runs.  Per (cell, gene) the path needs three K-term contractions forward
// (eta = nu.zeta, d = nu.zeta', nu.zeta'') and two backward (d/dnu += gE zeta + w omega zeta'): 31 of the
// 61 fp32 operations of the element.  The fp32 pipe, not HBM, bounded the first version of this kernel
// (DESIGN.md section 4), so those contractions now run on the tensor pipe as warp-level m16n8k8 TF32 MMAs
// with the split x = hi + lo that keeps fp32-level accuracy: A.B ~ Ahi.Bhi (one m16n8k8 TF32 MMA, SASS
// HMMA.1688.F32.TF32) + [Alo | A].[B ; Blo] (both cross terms in ONE m16n8k16 FP16 / BF16 MMA, HMMA.16816.F32:
// they are 2^-11 of the product, so 8-11 mantissa bits suffice); lo*lo is dropped.  The fp32 pipe keeps the
// negative-binomial terms, the MUFU pipe the 5 transcendentals.  The harmonics 1..2H (<= 6) fit one 8-wide
// k-step together with one spare slot that carries the per-cell size factor forward and sum_c w (= d/dgamma)
// backward; the constant term nu_0 - ln r (+ batch offset) is added in fp32 (it is the largest addend).  (tcgen05 is the wrong tool here: its operands live in shared memory, and
// the backward operand gE is produced in registers by the same threads that consume the forward result.)
//
// Mapping (genes are the contiguous axis of the counts, preprocessing.py:193-194):
//   * a CTA owns a gene tile of 32*NPAIR*warps genes and a contiguous range of 8-cell groups;
//   * a warp owns 32*NPAIR genes = 2*NPAIR MMA row tiles; lane (grp = lane/4, q = lane%4) owns 4 adjacent
//     genes per pair (rows grp and grp+8 of two row tiles) x the cells q and q+4 of the group.  Its nu
//     fragments (hi/lo), dispersion, kinetics and every per-gene accumulator stay in registers for the whole
//     kernel (no atomics on the hot path).  The forward accumulator fragment of a row tile IS the backward
//     A fragment ({c0,c2,c1,c3} -> {a0,a1,a2,a3}) once the forward B columns are permuted (column 2j -> cell j,
//     2j+1 -> cell j+4), so nothing is transposed or shuffled between the two GEMMs;
//   * counts stream through a shared-memory ring of kCountDepth 8-cell groups filled with 128-bit cp.async
//     (LDGSTS.128, zero-filled outside the matrix): every warp loads exactly the 32 genes x 8 cells x {S, U} per
//     group it consumes itself (a warp instruction covers 4 rows x one full 128-byte line), so the count stream
//     needs no barrier beyond cp.async.wait_group + __syncwarp and a slow warp never delays another warp's loads;
//   * the operand table of a group (built by vcb_cell_tables_kernel: the B fragments of every MMA, already split
//     and stored in lane order, 2.6 KB) is shared by all warps and staged by the TMA engine (cp.async.bulk +
//     mbarrier complete_tx) in a second, deeper ring; its issue rotates over the warps;
//   * per-cell sums (d/dphi, d/dcf, d/domega run over genes, i.e. over grp lanes and warps) are reduced
//     once per group with 9 shuffles, parked per warp in shared memory and summed in a fixed order by the
//     warp that re-issues the slot's table copy (deterministic, no atomics);
//   * no CTA-wide barrier and no warp waits for another warp in steady state: full[]/done[] mbarriers only.
//
// Arithmetic per (cell, gene), SURVEY.md Appendix A, in units of mu/r and base-2 logs so that every
// transcendental is a single MUFU op and the n r log r terms cancel analytically:
//   y = (etaS - ln r) log2e; u = 2^y = muS/r; s = 1+u; LS = lg2 s; gS = (kS - r u)/s
//   a = d*omega+gamma; m = relu(a)+1e-5; mb = m/beta; uU = u mb; sU = 1+uU; LU = lg2 sU
//   w0 = (kU - r uU)/(sU m); gU = w0 m; w = 1[a>0] w0      (1/s and 1/(sU m) share one MUFU.RCP)
//   log-prob pieces: kS (y-LS), LS, kU (y+lg2 mb-LU), LU   (times ln2, plus per-gene terms, in the epilogue)
#pragma once
#include "vcb_common.cuh"

namespace vcb {

constexpr int kGroupCells = 8;  // cells per ring stage = the n extent (forward) / k extent (backward) of the MMAs
constexpr int kCountDepth = 4;  // count groups in flight per thread (4 x 32 KB per 512-thread CTA)
constexpr int kMaxStages = 16;  // the table ring is as deep as the rest of shared memory allows, up to this
constexpr int kSmemHeader = 512;

// A warp owns 32*NPAIR genes.  NPAIR = 1: up to 512 threads/CTA at <= 128 registers; NPAIR = 2: up to 256.
__host__ __device__ constexpr int max_threads(int NPAIR) { return NPAIR == 1 ? 512 : 256; }

// per-gene partial rows written by the streaming kernel
enum GeneRow { ROW_AS = 0, ROW_LS = 1, ROW_AU = 2, ROW_LU = 3, ROW_GU = 4, ROW_W = 5, ROW_PSI = 6, ROW_DNU = 7 };
__host__ __device__ constexpr int gene_rows(int H) { return ROW_DNU + 2 * H + 1; }

// Operand table of one 8-cell group: sections of [k-step][lane][4] words = {TF32 b0, TF32 b1 of the main MMA,
// BF16x2 b0, BF16x2 b1 of the cross-term MMA}, then omega[8], batch id[8] and {the batch id if all 8 agree else -1, 0, 0, 0}.
// Slots of a k-step row: 0 -> constant, 1..2H -> harmonics, 2H+1 -> spare.
enum TabSection { SEC_F0 = 0, SEC_F1 = 1, SEC_B0 = 2, SEC_F2 = 3, SEC_B1 = 4 };
__host__ __device__ constexpr int ksteps(int H) { return (2 * H + 2 + 7) / 8; }
__host__ __device__ constexpr int table_sections(bool velo) { return velo ? 5 : 3; }
__host__ __device__ constexpr int table_tail(int H, bool velo) { return table_sections(velo) * ksteps(H) * 128; }
__host__ __device__ constexpr int table_group_floats(int H, bool velo) { return table_tail(H, velo) + 20; }
// forward B column n of a group holds cell fwd_cell(n): the accumulator columns {2q, 2q+1} of a lane are then
// the cells {q, q+4} that the same lane must supply as backward A columns
__host__ __device__ constexpr int fwd_cell(int n) { return (n >> 1) + 4 * (n & 1); }

struct StreamParams {
  const float* S;
  const float* U;
  const float* tab;  // [n_groups][table_group_floats]
  const float* nu;
  const float* dnu;
  const float* shape_inv;
  const float* logbeta;
  const float* gamma;
  float* genepart;  // [n_split][rows][ld]
  float* cellpart;  // [n_tiles][Ncp][NQ]
  float* d_dnu;     // [Nb][Ng], zeroed, atomically accumulated at batch boundaries
  long long Nc, Ng, ld, Ncp;
  int n_split;
  int Nb;
  int n_ring;              // depth of the table ring
};

struct StreamSmem {
  int part_off, gene_off, aop_off, tab_off, cnt_off, total;  // byte offsets inside dynamic shared memory
};

__host__ __device__ inline StreamSmem stream_smem_layout(int H, bool velo, int nwarps, int npair, int n_ring) {
  StreamSmem L;
  int off = kSmemHeader;  // mbarriers
  L.part_off = off;
  off += n_ring * nwarps * (velo ? 3 : 2) * 32 * 4;  // cell partials: [slot][warp][quantity][lane]
  off = (off + 127) / 128 * 128;
  L.gene_off = off;  // per-gene parameters: [warp][row tile][2][grp] float4
  off += nwarps * 2 * npair * 2 * 8 * 16;
  L.aop_off = off;  // forward A operands (nu fragments): [warp][row tile][k-step][main / cross][lane] uint4
  off += nwarps * 2 * npair * ksteps(H) * 2 * 32 * 16;
  L.tab_off = off;
  off += n_ring * table_group_floats(H, velo) * 4;
  off = (off + 127) / 128 * 128;
  L.cnt_off = off;
  off += kCountDepth * (velo ? 2 : 1) * 2 * npair * (32 * nwarps) * 16;  // [depth][matrix][cell q / q+4][pair][thread] x 16 B
  L.total = off;
  return L;
}

// ---- packed helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
#ifdef VCB_EXP_NOMUFU
__device__ __forceinline__ float2 ex2_2(float2 a) { return f2(a.x * 0.001f + 1.f, a.y * 0.001f + 1.f); }
__device__ __forceinline__ float2 lg2_2(float2 a) { return f2(a.x * 0.5f - 0.5f, a.y * 0.5f - 0.5f); }
__device__ __forceinline__ float2 rcp_2(float2 a) { return f2(2.f - a.x, 2.f - a.y); }
#else
__device__ __forceinline__ float2 ex2_2(float2 a) { return f2(ex2_approx(a.x), ex2_approx(a.y)); }
__device__ __forceinline__ float2 lg2_2(float2 a) { return f2(lg2_approx(a.x), lg2_approx(a.y)); }
__device__ __forceinline__ float2 rcp_2(float2 a) { return f2(rcp_approx(a.x), rcp_approx(a.y)); }
#endif

// ---- cp.async (LDGSTS.128): 16 bytes global -> shared, zero-filled when src_bytes == 0 ---------------------------
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gmem_src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- tensor-core helpers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(tf32_rna(x)); }
// two floats -> one register of two BF16; `k_even` lands in the low half (the lower k index of an MMA operand pair)
__device__ __forceinline__ uint32_t pack_bf16(float k_even, float k_odd) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(k_odd), "f"(k_even));
  return r;
}
// same for FP16 (3 more mantissa bits than BF16; used where the operands are bounded: nu and the basis tables);
// values beyond the FP16 range saturate instead of turning into infinities
__device__ __forceinline__ uint32_t pack_f16(float k_even, float k_odd) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(k_odd), "f"(k_even));
  return r;
}
// Fragments (grp = lane/4, q = lane%4).  m16n8k8 TF32:
//   a0 (grp, q)  a1 (grp+8, q)  a2 (grp, q+4)  a3 (grp+8, q+4);  b0 (k=q, n=grp)  b1 (k=q+4, n=grp);
//   c0 (grp, 2q)  c1 (grp, 2q+1)  c2 (grp+8, 2q)  c3 (grp+8, 2q+1)
// m16n8k16 BF16 (two k per register, lower k in the low half):
//   a0 (grp, k=2q,2q+1)  a1 (grp+8, 2q,2q+1)  a2 (grp, 2q+8,2q+9)  a3 (grp+8, 2q+8,2q+9);
//   b0 (k=2q,2q+1, n=grp)  b1 (k=2q+8,2q+9, n=grp);  c as above
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// C += A.B with A = amain (TF32 hi) and the cross terms folded into across = [Alo | A], small terms first;
// b = one table entry {main b0, main b1, cross b0, cross b1}.  Forward (nu x basis tables, bounded operands) the
// cross operands are FP16: error 2^-11 * 2^-12 of the product, i.e. fp32 level.  Backward (the A operand is a
// per-element gradient of unbounded range) they are BF16: 2^-20 per element, random over the cells of a sum.
__device__ __forceinline__ void mma_split_fwd(float (&c)[4], const uint32_t (&amain)[4], const uint32_t (&across)[4],
                                              const float4 b) {
  mma_f16(c, across, __float_as_uint(b.z), __float_as_uint(b.w));
  mma_tf32(c, amain, __float_as_uint(b.x), __float_as_uint(b.y));
}
__device__ __forceinline__ void mma_split_bwd(float (&c)[4], const uint32_t (&amain)[4], const uint32_t (&across)[4],
                                              const float4 b) {
  mma_bf16(c, across, __float_as_uint(b.z), __float_as_uint(b.w));
  mma_tf32(c, amain, __float_as_uint(b.x), __float_as_uint(b.y));
}
// backward A operands of one row tile from the gene-pair values of the cells q (v0) and q+4 (v1)
__device__ __forceinline__ void split_operand(const float2 v0, const float2 v1, uint32_t (&amain)[4],
                                              uint32_t (&across)[4]) {
  amain[0] = __float_as_uint(v0.x) & 0xffffe000u;
  amain[1] = __float_as_uint(v0.y) & 0xffffe000u;
  amain[2] = __float_as_uint(v1.x) & 0xffffe000u;
  amain[3] = __float_as_uint(v1.y) & 0xffffe000u;
  const float2 l0 = __fadd2_rn(v0, make_float2(-__uint_as_float(amain[0]), -__uint_as_float(amain[1])));
  const float2 l1 = __fadd2_rn(v1, make_float2(-__uint_as_float(amain[2]), -__uint_as_float(amain[3])));
  across[0] = pack_bf16(l0.x, l1.x);
  across[1] = pack_bf16(l0.y, l1.y);
  across[2] = pack_bf16(v0.x, v1.x);
  across[3] = pack_bf16(v0.y, v1.y);
}

template <bool B>
struct BoolTag {
  static constexpr bool value = B;
};

template <int H, bool VELO, bool GRAD, bool LGINLINE, int NPAIR>
__global__ void __launch_bounds__(max_threads(NPAIR), 1) vcb_stream_kernel(const StreamParams P) {
  constexpr int K = 2 * H + 1;
  constexpr int KS = ksteps(H);
  constexpr int NT = 2 * NPAIR;  // MMA row tiles per warp
  constexpr int NMAT = VELO ? 2 : 1;
  constexpr int NQ = VELO ? 3 : 2;
  constexpr int R = kGroupCells;
  constexpr int D = kCountDepth;
  constexpr int NLD = NMAT * 2 * NPAIR;  // 16-byte count loads per thread and group: [matrix][cell q / q+4][pair]
  constexpr int TABG = table_group_floats(H, VELO);
  constexpr int TAIL = table_tail(H, VELO);
  constexpr bool NEED_D = GRAD || VELO;
  constexpr bool NEED_E = GRAD && VELO;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  const int grp = lane >> 2, q = lane & 3;
  const int NS = P.n_ring;
  const int tile = blockIdx.x;
  const int split = blockIdx.y;
  const int WT = 32 * NPAIR * nwarps;  // genes per CTA tile
  const long long g_base = (long long)tile * WT;
  const long long rem = P.ld - g_base;
  const int W = (int)(rem < (long long)WT ? rem : (long long)WT);  // genes of this tile that exist in a row
  const StreamSmem L = stream_smem_layout(H, VELO, nwarps, NPAIR, NS);
  const uint32_t full0 = smem_u32(smem_raw);     // shared-window addresses of full[0] / done[0] ([kMaxStages] each;
  const uint32_t done0 = full0 + 8 * kMaxStages;  // roles documented where they are initialised)
  float* s_part = reinterpret_cast<float*>(smem_raw + L.part_off);
  float* s_tab = reinterpret_cast<float*>(smem_raw + L.tab_off);

  // this CTA's cell groups
  const long long n_groups = P.Ncp / R;
  const long long G0 = (n_groups * split) / P.n_split;
  const long long G1 = (n_groups * (split + 1)) / P.n_split;
  const int n_stages = (int)(G1 - G0);

  // ---- per-gene state in registers ------------------------------------------------------------------
  // pair p, gene j in 0..3: g = g_base + (warp*NPAIR + p)*32 + 4*grp + j; row tile mt = 2p + (j>>1) holds
  // gene j in fragment rows grp (j even) and grp+8 (j odd).
  int gl[NPAIR];       // tile-local index of the lane's first gene of pair p
  bool gvalid[NPAIR];  // the lane's 4 genes of pair p lie inside the row pitch
#pragma unroll
  for (int p = 0; p < NPAIR; ++p) {
    gl[p] = (warp * NPAIR + p) * 32 + 4 * grp;
    gvalid[p] = gl[p] < W;
  }
  auto gene_of = [&](int mt, int odd) -> long long { return g_base + gl[mt >> 1] + 2 * (mt & 1) + odd; };
  // forward A operand (gene g, slot): [0, nu_1..nu_2H, 1, 0...]; a padding gene is all zero
  auto a_value = [&](long long g, int slot) -> float {
    if (g >= P.Ng || slot == 0 || slot > K) return 0.f;
    return slot == K ? 1.f : P.nu[g * K + slot];
  };
  // the constant addend nu_0 - ln r (+ batch offset); -inf on a padding gene makes every contribution zero
  auto const_term = [&](long long g, int b) -> float {
    if (g >= P.Ng) return -1e30f;
    float v = P.nu[g * K] + logf(P.shape_inv[g]);
    if (b >= 0) v += P.dnu[(long long)b * P.Ng + g];
    return v;
  };
  // The nu fragments of the forward MMAs (TF32 hi for the main product, FP16 [lo | value] for the cross terms)
  // are this lane's own, but 16 registers is more than the loop can spare: they wait in shared memory and come back
  // with two LDS.128 per row tile and group.
  uint4* const s_aop = reinterpret_cast<uint4*>(smem_raw + L.aop_off) + (size_t)warp * (NT * KS * 2 * 32) + lane;
  // Dispersion, kinetics and the constant term live in shared memory (the 4 lanes of a grp share them; one
  // broadcast LDS.128 per row tile brings {-r, c0} and {gamma, 1/beta} as gene pairs): 16 registers fewer.
  float4* const s_gene = reinterpret_cast<float4*>(smem_raw + L.gene_off) + (size_t)warp * (NT * 2 * 8) + grp;
#pragma unroll
  for (int mt = 0; mt < NT; ++mt) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t am[4], ax[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) am[i] = tf32_rna(a_value(gene_of(mt, i & 1), 8 * ks + q + 4 * (i >> 1)));
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        const float x0 = a_value(gene_of(mt, o), 8 * ks + 2 * q), x1 = a_value(gene_of(mt, o), 8 * ks + 2 * q + 1);
        ax[o] = pack_f16(tf32_lo(x0), tf32_lo(x1));
        ax[2 + o] = pack_f16(x0, x1);
      }
      s_aop[((mt * KS + ks) * 2 + 0) * 32] = make_uint4(am[0], am[1], am[2], am[3]);
      s_aop[((mt * KS + ks) * 2 + 1) * 32] = make_uint4(ax[0], ax[1], ax[2], ax[3]);
    }
    if (q == 0) {
      float rs[2], gs[2], ibs[2];
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        const long long g = gene_of(mt, o);
        const bool ok = g < P.Ng;
        rs[o] = ok ? 1.0f / P.shape_inv[g] : 1.f;
        gs[o] = (ok && VELO) ? P.gamma[g] : 1.f;
        ibs[o] = (ok && VELO) ? expf(-P.logbeta[g]) : 1.f;
      }
      s_gene[(mt * 2 + 0) * 8] = make_float4(-rs[0], -rs[1], const_term(gene_of(mt, 0), -1), const_term(gene_of(mt, 1), -1));
      s_gene[(mt * 2 + 1) * 8] = make_float4(gs[0], gs[1], ibs[0], ibs[1]);
    }
  }
  __syncwarp();
  auto set_batch = [&](int b) {
    __syncwarp();
    if (q == 0) {
#pragma unroll
      for (int mt = 0; mt < NT; ++mt) {
        float4 v = s_gene[(mt * 2 + 0) * 8];
        v.z = const_term(gene_of(mt, 0), b);
        v.w = const_term(gene_of(mt, 1), b);
        s_gene[(mt * 2 + 0) * 8] = v;
      }
    }
    __syncwarp();
  };
  int cur_b = -1;

  const float2 zero2 = f2s(0.f);
  float2 accAS[NT], accLS[NT], accAU[NT], accLU[NT], accGU[NT], accPsi[NT];
  // d/dnu fragments.  Tensor-core accumulation truncates, and a long chain of small addends into a large sum
  // would drift: the MMA accumulators start from zero in every group and are added here with fp32 rounding.
  float accNu[NT][KS][4];
#pragma unroll
  for (int mt = 0; mt < NT; ++mt) {
    accAS[mt] = accLS[mt] = accAU[mt] = accLU[mt] = accGU[mt] = accPsi[mt] = zero2;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int i = 0; i < 4; ++i) accNu[mt][ks][i] = 0.f;
  }
  // batch boundary: the constant column of d/dnu (slot 0: lanes q == 0, c0 / c2) is the batch's d/ddnu
  auto flush_batch = [&]() {
    if (q == 0 && cur_b >= 0) {
#pragma unroll
      for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const long long g = gene_of(mt, o);
          if (g < P.Ng) atomicAdd(&P.d_dnu[(long long)cur_b * P.Ng + g], accNu[mt][0][2 * o]);
          accNu[mt][0][2 * o] = 0.f;
        }
    }
  };

  // ---- count ring: warp-private cp.async pipeline -----------------------------------------------------------
  // Consumer lane (grp, q) reads slot (d, j) = s_cnt[(d*NLD + j)*nthr], j = (matrix*2 + cc)*NPAIR + p: the 16 bytes
  // of its 4 genes at cell q + 4cc.  The LOADS use another lane mapping: 8 consecutive lanes fetch one full
  // 128-byte line (row lane/8, 16-byte chunk lane%8) -- with the consumer mapping a quarter-warp would touch 4
  // rows x 32 B, four times the memory requests (measured: 4.5 vs 6.9 TB/s, tools/ubench_tma.cu) -- and write
  // into the slot of the lane that will consume them, so a __syncwarp() separates loads from reads.  Loads outside
  // the matrix (cells >= Nc, genes past the pitch) are zero-filled by the copy itself: the consumer never masks.
  const int l_row = lane >> 3, l_chunk = lane & 7;
  const uint32_t s_cnt_ld = smem_u32(reinterpret_cast<float4*>(smem_raw + L.cnt_off) + (warp * 32 + 4 * l_chunk + l_row));
  const uint32_t slot_bytes = (uint32_t)nthr * 16u;  // distance between two of a thread's slots
  const uint32_t half_off = (uint32_t)(16 * P.ld);   // bytes from row r to row r+4 (ld <= 2^24)
  const long long row_step = 4ll * R * P.ld;         // bytes from one group to the next
  // Row pointers of the next stage to load.  A lane whose genes lie past the row pitch copies 0 bytes (= zero
  // fill) from a clamped, valid address, so the steady-state path has no per-lane branches.
  const char* ldS[NPAIR];
  const char* ldU[NPAIR];
  uint32_t ld_sz[NPAIR];
#pragma unroll
  for (int p = 0; p < NPAIR; ++p) {
    const int lgl = (warp * NPAIR + p) * 32 + 4 * l_chunk;
    const long long off = (G0 * R + l_row) * P.ld + g_base + (lgl < W ? lgl : 0);
    ld_sz[p] = lgl < W ? 16u : 0u;
    ldS[p] = reinterpret_cast<const char*>(P.S + off);
    ldU[p] = reinterpret_cast<const char*>((VELO ? P.U : P.S) + off);
  }
  // stages [0, n_whole) of this CTA lie entirely inside the matrix; only the very last group of the matrix can be ragged
  const int n_whole = (int)((P.Nc / R > G0 ? P.Nc / R - G0 : 0) < (long long)n_stages ? (P.Nc / R > G0 ? P.Nc / R - G0 : 0)
                                                                                       : (long long)n_stages);
  auto load_counts = [&](int st, int d) {  // stage st (loaded in order) into depth slot d
    if (st < n_stages) {
      uint32_t dst = s_cnt_ld + (uint32_t)d * (NLD * slot_bytes);
      if (st < n_whole) {  // CTA-uniform
#pragma unroll
        for (int mat = 0; mat < NMAT; ++mat)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc)
#pragma unroll
            for (int p = 0; p < NPAIR; ++p) {
              cp_async16(dst, (mat ? ldU[p] : ldS[p]) + (cc ? half_off : 0u), ld_sz[p]);
              dst += slot_bytes;
            }
      } else {
        const long long rows_left = P.Nc - (G0 + st) * R;
#pragma unroll
        for (int mat = 0; mat < NMAT; ++mat)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc)
#pragma unroll
            for (int p = 0; p < NPAIR; ++p) {
              const bool ok = 4 * cc + l_row < rows_left;
              cp_async16(dst, ok ? (mat ? ldU[p] : ldS[p]) + (cc ? half_off : 0u) : reinterpret_cast<const char*>(P.S),
                         ok ? ld_sz[p] : 0u);
              dst += slot_bytes;
            }
      }
#pragma unroll
      for (int p = 0; p < NPAIR; ++p) {
        ldS[p] += row_step;
        ldU[p] += row_step;
      }
    }
    cp_async_commit();  // one group per stage, empty past the end: wait_group counts stay uniform
  };

  // ---- table ring (TMA) and the per-cell partial sums -------------------------------------------------------
  //   full[s] : the operand table in slot s has landed              (arrive.expect_tx + complete_tx)
  //   done[s] : every warp finished with slot s (table, partials)   (one arrival per warp)
  // Stage v is issued by warp v % nwarps, which first drains the slot's cell partials (summed in warp order:
  // deterministic).  Each warp walks its own stages with a private cursor: no shared cursor, no CAS.
  // The ring is deep (up to 16 groups), so a refill is never urgent and nobody spins on done[].
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(done0 + 8 * s, (uint32_t)nwarps);
    }
    mbar_fence_init();
  }
  __syncthreads();
  float* const cellpart_t = P.cellpart + (long long)tile * P.Ncp * NQ;
  // Cell partials of a stage: every lane parks NQ values for ONE cell (q + 4*(lane/16)), summed over its 4 genes and
  // one shuffle level; the warp that re-issues the slot adds the 4 lanes x nwarps terms of each of the 8*NQ outputs in
  // a fixed order (deterministic) and stores them.  (A full butterfly per warp and group cost 15 % of the loop.)
  auto flush_partials = [&](int st, int slot) {
    if (lane < R * NQ) {
      const int cell = lane / NQ, i = lane - cell * NQ;
      const float* src = s_part + (size_t)slot * nwarps * (NQ * 32) + i * 32 + (cell & 3) + 16 * (cell >> 2);
      float s = 0.f;
      for (int w = 0; w < nwarps; ++w) {
        const float* sw = src + w * (NQ * 32);
        s += (sw[0] + sw[4]) + (sw[8] + sw[12]);
      }
      cellpart_t[(G0 + st) * (R * NQ) + lane] = s;
    }
  };
  // issue cursor over this warp's stages: stage i_next = i_k*NS + i_slot; i_need = the stage that must have left
  // the slot (i_next - NS), pushed out of reach once the warp has nothing more to issue
  int i_next = warp, i_slot = warp % NS, i_k = warp / NS;
  int i_need = i_next < n_stages ? i_next - NS : 0x3fffffff;
  const uint32_t tab_bytes = (uint32_t)TABG * 4u;
  auto issue_table = [&](int consumed) -> bool {  // consumed = stages this warp has finished; warp-uniform; never blocks
    // A parity wait only distinguishes adjacent phases: do not look at done[] before this warp itself has
    // arrived for stage i_need (with fewer slots than warps its next stage is two phases ahead).
    if (i_need >= consumed) return false;
    if (i_k > 0) {  // the slot held stage i_need: wait until every warp has left it, then drain its partials
      int ready = 0;
      if (lane == 0) ready = mbar_test_wait(done0 + 8 * i_slot, (uint32_t)((i_k - 1) & 1)) ? 1 : 0;
      if (!__shfl_sync(0xffffffffu, ready, 0)) return false;
      if (GRAD) flush_partials(i_need, i_slot);
      __syncwarp();
    }
    if (lane == 0) {
      mbar_expect_tx(full0 + 8 * i_slot, tab_bytes);
      bulk_g2s(smem_u32(s_tab) + (uint32_t)i_slot * tab_bytes, P.tab + (G0 + i_next) * TABG, tab_bytes, full0 + 8 * i_slot);
    }
    i_next += nwarps;
    i_slot += nwarps;
    while (i_slot >= NS) {
      i_slot -= NS;
      ++i_k;
    }
    i_need = i_next < n_stages ? i_next - NS : 0x3fffffff;
    return true;
  };
  while (i_k == 0 && issue_table(0)) {}  // prologue: the slots are fresh
#pragma unroll
  for (int s = 0; s < D; ++s) load_counts(s, s);

  const float2 one2 = f2s(1.f), neg1 = f2s(-1.f), l2e = f2s(kLog2e), eps2 = f2s(1e-5f);
  int c_slot = 0, c_phase = 0, c_d = 0;  // consumer cursor: table slot / parity, count depth slot
  const float4* const s_cnt = reinterpret_cast<const float4*>(smem_raw + L.cnt_off) + tid;  // this lane's column

  // One 8-cell group.  MIXED (compile-time copy of the loop body): the group's cells belong to different batches
  // (unsorted input), so batch offsets are added per element and d/ddnu goes through atomics -- rare and slow.
  auto process = [&](auto mixed_tag, const int st, const float* tb, const float4* cnt, float (&pcf)[2], float (&pphi)[2],
                     float (&pom)[2]) {
    constexpr bool MIXED = decltype(mixed_tag)::value;
    const float4* tb4 = reinterpret_cast<const float4*>(tb);
    int bc[2] = {0, 0};
    if (MIXED) {
      bc[0] = __float_as_int(tb[TAIL + 8 + q]);
      bc[1] = __float_as_int(tb[TAIL + 12 + q]);
    }
    float om[2] = {0.f, 0.f};
    if (VELO) {
      om[0] = tb[TAIL + q];
      om[1] = tb[TAIL + q + 4];
    }
    float2 pcf2[2] = {zero2, zero2}, pphi2[2] = {zero2, zero2}, pom2[2] = {zero2, zero2};
    const float2* cnt2 = reinterpret_cast<const float2*>(cnt);

    // Row tiles (16 genes x 8 cells) are software-pipelined: the forward MMAs of tile mt+1 are issued before the
    // element terms of tile mt, so that the tensor pipe's latency hides behind fp32 / MUFU work of the same warp;
    // element terms -> backward MMAs of a tile stay together so that only two tiles' fragments are ever live.
    float Cf[2][3][4];  // [pipeline slot][eta, d, omega*d''][fragment]
    auto forward = [&](int mt, float (&C)[3][4]) {
#pragma unroll
      for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) C[j][i] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const uint4 a0 = s_aop[((mt * KS + ks) * 2 + 0) * 32], a1 = s_aop[((mt * KS + ks) * 2 + 1) * 32];
        const uint32_t am[4] = {a0.x, a0.y, a0.z, a0.w}, ax[4] = {a1.x, a1.y, a1.z, a1.w};
        mma_split_fwd(C[0], am, ax, tb4[(SEC_F0 * KS + ks) * 32 + lane]);
        if (NEED_D) mma_split_fwd(C[1], am, ax, tb4[(SEC_F1 * KS + ks) * 32 + lane]);
        if (NEED_E) mma_split_fwd(C[2], am, ax, tb4[(SEC_F2 * KS + ks) * 32 + lane]);
      }
    };
    forward(0, Cf[0]);
#pragma unroll
    for (int mt = 0; mt < NT; ++mt) {
      const int p = mt >> 1, t = mt & 1;
      if (mt + 1 < NT) forward(mt + 1, Cf[(mt + 1) & 1]);
      float(&Ce)[4] = Cf[mt & 1][0];
      float(&Cd)[4] = Cf[mt & 1][1];
      float(&Cw)[4] = Cf[mt & 1][2];
      const float4 ga = s_gene[(mt * 2 + 0) * 8];
      const float2 nr_mt = f2(ga.x, ga.y);
      float2 gam_mt = one2, invb_mt = one2;
      if (VELO) {
        const float4 gb = s_gene[(mt * 2 + 1) * 8];
        gam_mt = f2(gb.x, gb.y);
        invb_mt = f2(gb.z, gb.w);
      }

      // ---- per-element negative-binomial terms (gene pair x cells q, q+4) ------------------------------------
      float2 Gg[2], Gw[2];  // backward A operands of this row tile
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        // this lane's slot j holds its 4 genes of pair p: .xy row tile 2p, .zw row tile 2p+1
        const float2 kS = cnt2[((size_t)(cc * NPAIR + p) * nthr) * 2 + t];
        const float2 kU = VELO ? cnt2[((size_t)((2 + cc) * NPAIR + p) * nthr) * 2 + t] : kS;
        float2 c0 = f2(ga.z, ga.w);
        if (MIXED) {  // the table's constant term carries the offset of cur_b: use the cell's own batch instead
          const long long g0 = gene_of(mt, 0);
          c0 = f2(const_term(g0, bc[cc]), const_term(g0 + 1, bc[cc]));
        }
        // (the accumulator fragment holds the gene pair in registers c[cc], c[2+cc]: the first operation on it is
        //  issued per gene, its result lands in an aligned pair and everything after is packed f32x2)
        const float2 d = f2(Cd[cc], Cd[2 + cc]);
        const float2 y = f2((Ce[cc] + c0.x) * kLog2e, (Ce[2 + cc] + c0.y) * kLog2e);
        const float2 u = ex2_2(y);
        const float2 s = add2(u, one2);
        const float2 LS = lg2_2(s);
        accAS[mt] = fma2(kS, fma2(LS, neg1, y), accAS[mt]);
        accLS[mt] = add2(accLS[mt], LS);
        if (LGINLINE) {
          float2 psi, lg;
          lg.x = lgamma_terms_inline(-nr_mt.x, kS.x, psi.x);
          lg.y = lgamma_terms_inline(-nr_mt.y, kS.y, psi.y);
          accAS[mt] = fma2(lg, l2e, accAS[mt]);
          accPsi[mt] = add2(accPsi[mt], psi);
        }
        float2 g = zero2, w = zero2;
        if (VELO) {
          const float2 a = f2(fmaf(Cd[cc], om[cc], gam_mt.x), fmaf(Cd[2 + cc], om[cc], gam_mt.y));
          const float2 m = add2(f2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f)), eps2);
          const float2 mb = mul2(m, invb_mt);
          const float2 uU = mul2(u, mb);
          const float2 sU = add2(uU, one2);
          const float2 LU = lg2_2(sU);
          const float2 lmb = lg2_2(mb);
          accAU[mt] = fma2(kU, fma2(LU, neg1, add2(y, lmb)), accAU[mt]);
          accLU[mt] = add2(accLU[mt], LU);
          if (LGINLINE) {
            float2 psi, lg;
            lg.x = lgamma_terms_inline(-nr_mt.x, kU.x, psi.x);
            lg.y = lgamma_terms_inline(-nr_mt.y, kU.y, psi.y);
            accAU[mt] = fma2(lg, l2e, accAU[mt]);
            accPsi[mt] = add2(accPsi[mt], psi);
          }
          if (GRAD) {
            const float2 sUm = mul2(sU, m);
            const float2 rc = rcp_2(mul2(s, sUm));  // one MUFU for 1/s and 1/(sU m)
            const float2 inv_s = mul2(rc, sUm), inv_sUm = mul2(rc, s);
            const float2 gS = mul2(fma2(nr_mt, u, kS), inv_s);
            const float2 w0 = mul2(fma2(nr_mt, uU, kU), inv_sUm);
            const float2 gU = mul2(w0, m);
            w = f2(a.x > 0.f ? w0.x : 0.f, a.y > 0.f ? w0.y : 0.f);
            g = add2(gS, gU);
            accGU[mt] = add2(accGU[mt], gU);
            pom2[cc] = fma2(w, d, pom2[cc]);
            pphi2[cc] = fma2(w, f2(Cw[cc], Cw[2 + cc]), pphi2[cc]);
          }
        } else if (GRAD) {
          g = mul2(fma2(nr_mt, u, kS), rcp_2(s));
        }
        if (GRAD) {
          pcf2[cc] = add2(pcf2[cc], g);
          pphi2[cc] = fma2(g, d, pphi2[cc]);
          if (MIXED) {
            if ((G0 + st) * R + q + 4 * cc < P.Nc) {
              const long long g0 = gene_of(mt, 0), brow = (long long)bc[cc] * P.Ng;
              if (g0 < P.Ng) atomicAdd(&P.d_dnu[brow + g0], g.x);
              if (g0 + 1 < P.Ng) atomicAdd(&P.d_dnu[brow + g0 + 1], g.y);
            }
          }
        }
        Gg[cc] = g;
        Gw[cc] = w;
      }

      // ---- backward contractions on the tensor pipe: acc[gene][slot] = sum_cells G[gene][cell] Z[cell][slot] -----
      if (GRAD) {
        uint32_t gm[4], gx[4], wm[4], wx[4];
        split_operand(Gg[0], Gg[1], gm, gx);
        if (VELO) split_operand(Gw[0], Gw[1], wm, wx);
#pragma unroll
        for (int nt = 0; nt < KS; ++nt) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          float4 bz = tb4[(SEC_B0 * KS + nt) * 32 + lane];
          if (MIXED && nt == 0 && grp == 0) bz.x = bz.y = bz.z = 0.f;  // constant column off: d/ddnu went through atomics
          mma_split_bwd(acc, gm, gx, bz);
          if (VELO) mma_split_bwd(acc, wm, wx, tb4[(SEC_B1 * KS + nt) * 32 + lane]);
#pragma unroll
          for (int i = 0; i < 4; ++i) accNu[mt][nt][i] += acc[i];
        }
      }
    }
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      pcf[cc] = pcf2[cc].x + pcf2[cc].y;
      pphi[cc] = pphi2[cc].x + pphi2[cc].y;
      pom[cc] = pom2[cc].x + pom2[cc].y;
    }
  };

  // ---- main loop ----------------------------------------------------------------------------------
  for (int st = 0; st < n_stages; ++st) {
    {
      uint32_t spins = 0;
      while (!__all_sync(0xffffffffu, mbar_try_wait(full0 + 8 * c_slot, (uint32_t)c_phase))) {
        issue_table(st);
        __nanosleep(128);  // the table is not there yet: leave the issue slots to the warps that have work
        if (++spins > (1u << 22)) __trap();
      }
    }
    cp_async_wait<D - 1>();  // this lane's loads of stage st have landed ...
    __syncwarp();            // ... and so have the other lanes' (counts never cross a warp)
    const float* tb = s_tab + (size_t)c_slot * TABG;
    const float4* cnt = s_cnt + (size_t)c_d * NLD * nthr;
    float pcf[2] = {0.f, 0.f}, pphi[2] = {0.f, 0.f}, pom[2] = {0.f, 0.f};

    {
      // batch bookkeeping (CTA-uniform decisions: every warp reads the same table)
      bool mixed = false;
      if (P.Nb > 0) {
        const int gb = __float_as_int(tb[TAIL + 16]);  // the group's batch, -1 if its cells disagree
        mixed = gb < 0;
        if (!mixed && gb != cur_b) {
          if (GRAD) flush_batch();
          cur_b = gb;
          set_batch(cur_b);
        }
      }
      if (mixed)
        process(BoolTag<true>{}, st, tb, cnt, pcf, pphi, pom);
      else
        process(BoolTag<false>{}, st, tb, cnt, pcf, pphi, pom);
    }

    if (GRAD) {
      // the lane holds partial sums for the cells q and q+4: swap halves so that lanes 0-15 keep cell q and lanes
      // 16-31 cell q+4, and park the result; the remaining 4 lanes x nwarps terms are added when the slot is drained
      const bool up = (lane & 16) != 0;
      float* dst = s_part + ((size_t)c_slot * nwarps + warp) * (NQ * 32) + lane;
      dst[0] = (up ? pcf[1] : pcf[0]) + __shfl_xor_sync(0xffffffffu, up ? pcf[0] : pcf[1], 16);
      dst[32] = (up ? pphi[1] : pphi[0]) + __shfl_xor_sync(0xffffffffu, up ? pphi[0] : pphi[1], 16);
      if (VELO) dst[(NQ - 1) * 32] = (up ? pom[1] : pom[0]) + __shfl_xor_sync(0xffffffffu, up ? pom[0] : pom[1], 16);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(done0 + 8 * c_slot);  // this warp no longer needs the slot (release: partials are visible)
    if (++c_slot == NS) {
      c_slot = 0;
      c_phase ^= 1;
    }
    load_counts(st + D, c_d);  // refill the count slot this warp has just consumed
    if (++c_d == D) c_d = 0;
    issue_table(st + 1);
  }
  cp_async_wait<0>();

  // ---- drain: the last ring-depth stages were never re-issued, their cell partials are still parked ----------
  if (GRAD) {
    __syncthreads();
    const int first = n_stages > NS ? n_stages - NS : 0;
    for (int st = first + warp; st < n_stages; st += nwarps) flush_partials(st, st % NS);
  }

  // ---- flush per-gene partial sums ----------------------------------------------------------------
  if (GRAD && P.Nb > 0) flush_batch();
  constexpr int ROWS = gene_rows(H);
  float* gp = P.genepart + ((long long)split * ROWS) * P.ld;
  auto lane_sum4 = [&](float2 v) -> float2 {  // over the 4 lanes (cells) that share a gene pair
    v.x += __shfl_xor_sync(0xffffffffu, v.x, 1);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, 1);
    v.x += __shfl_xor_sync(0xffffffffu, v.x, 2);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, 2);
    return v;
  };
#pragma unroll
  for (int mt = 0; mt < NT; ++mt) {
    const bool ok = gvalid[mt >> 1];
    const long long g0 = gene_of(mt, 0);  // even: the pair (g0, g0+1) is one aligned float2
    auto put = [&](int row, float2 v) {
      v = lane_sum4(v);
      if (ok && q == 0) *reinterpret_cast<float2*>(gp + (long long)row * P.ld + g0) = v;
    };
    put(ROW_AS, accAS[mt]);
    put(ROW_LS, accLS[mt]);
    if (VELO) {
      put(ROW_AU, accAU[mt]);
      put(ROW_LU, accLU[mt]);
      if (GRAD) put(ROW_GU, accGU[mt]);
    }
    if (LGINLINE) put(ROW_PSI, accPsi[mt]);
    if (GRAD && ok) {
      // accumulator fragment: c0 (gene g0, slot 2q) c1 (g0, 2q+1) c2 (g0+1, 2q) c3 (g0+1, 2q+1)
#pragma unroll
      for (int nt = 0; nt < KS; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int slot_k = 8 * nt + 2 * q + (i & 1);
          const long long g = g0 + (i >> 1);
          if (slot_k < K)
            gp[(long long)(ROW_DNU + slot_k) * P.ld + g] = accNu[mt][nt][i];
          else if (VELO && slot_k == K)
            gp[(long long)ROW_W * P.ld + g] = accNu[mt][nt][i];
        }
    }
  }
}

}  // namespace vcb
