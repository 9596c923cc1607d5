// The streaming kernel, second generation (round 2): same mapping and arithmetic as vcb_stream.cuh -- one pass over
// the [cells][genes] count tiles of S (and U) that yields the per-gene log-prob partial sums AND every gradient
// partial sum, contractions on the tensor pipe, counts through a warp-private cp.async ring, operand tables staged
// by the TMA engine -- rebuilt around the round-1 profile (profiles/r01_stream_kernel_ncu_summary.txt): that kernel
// was issue bound, 475 warp instructions per warp and 8-cell group of which ~190 were bookkeeping.  What changed:
//
//   * a ring stage is kGPS = 2 8-cell groups (16 cells): one table copy, one full[] wait, one done[] arrival, one batch
//     test and one issue test per 16 cells;
//   * stage service by rotation: warp (x + 2) % 16 services stage x when it enters stage x + 2 -- it waits on the done[]
//     mbarrier of x (all 16 warps have arrived; normally long ago), re-issues the slot's table copy for stage x + NS and
//     drains the stage's parked per-cell sums in warp order.  No polling, no completion counter, no warp spins;
//   * the backward MMAs run as two independent accumulation chains (g terms, w terms): a dependent mma.sync issues ~35
//     cycles after its predecessor, one six-deep chain per row tile cost 8 % of the kernel;
//   * nu, the constant term and the size-factor slot are pre-scaled by log2(e) in the A operands, omega by ln 2 in the
//     table, so the MMA result IS the base-2 exponent; the relu offset 1e-5 rides gamma: m = max(a + eps, eps);
//   * the backward operands use the 3xTF32 form  G.Z ~ Ghi.Zhi + Ghi.Zlo + Glo.Zhi  with Ghi = G & ~0x1fff and
//     Glo = G - Ghi passed as raw fp32: 6 instructions per operand and row tile instead of 10 (no 16-bit packing),
//     one more MMA on a tensor pipe that was 20 % busy;
//   * per-cell sums (d/dphi, d/dcf, d/domega) are accumulated per cell in scalar registers straight from the
//     accumulator fragments: no re-pairing moves;
//   * batch offsets: cells arrive sorted by batch (PackedCounts sorts the rows once per dataset), the offsets are
//     switched at stage boundaries, d/dDelta-nu of a (split, batch, gene) is written by the one thread that owns the
//     gene: no atomics, bit-reproducible.  A stage whose 16 cells disagree (the Nb - 1 batch boundaries of sorted data,
//     or every stage of a C-ABI caller with unsorted ids) is processed once per batch present with the other cells
//     masked out -- slow, correct and still deterministic;
//   * no debug branches, no NPAIR / inline-lgamma variants (those calls keep using vcb_stream.cuh).
#pragma once
#include "vcb_stream.cuh"

namespace vcb {
namespace s2 {

#ifdef VCB_EXP_GPS
constexpr int kGPS = VCB_EXP_GPS;
#else
// 8-cell groups per ring stage.  4 (with the ring depth halved: same shared memory, half the per-stage bookkeeping) loses 4 %:
// the servicing warp then waits for the slowest warp every stage and the warps run in lockstep through the same phases.
constexpr int kGPS = 2;
#endif
constexpr int kStageCells = kGPS * kGroupCells;  // 16
constexpr int kD = 4;                            // count groups in flight per warp
constexpr int kMaxNS = 8;                        // table-ring depth limit (mbarrier slots)
#ifdef VCB_EXP_THREADS
constexpr int kThreads = VCB_EXP_THREADS;
#else
constexpr int kThreads = 512;  // 16 warps x 128 registers: 20-24 warps at 80-96 registers spill and lose 30 % (tried)
#endif
constexpr int kHeader = 256;
// a stage is serviced (table slot refilled, parked cell partials drained) this many stages later
constexpr int kSvcDist = kGPS >= 4 ? 1 : 2;
constexpr float kRelEps = 1e-5f;

struct Params {
  const float* S;
  const float* U;
  const float* tab;  // [n_stages_total][kGPS][table_group_floats]
  const float* nu;
  const float* dnu;
  const float* shape_inv;
  const float* logbeta;
  const float* gamma;
  float* genepart;  // [n_split][rows][ld]
  float* cellpart;  // [n_tiles][Ncp][NQ]
  float* dnupart;   // [n_split][Nb][ld], zeroed by the caller; (split, gene) has exactly one writer
  long long Nc, Ng, ld, Ncp;
  long long n_stages_total;
  int n_split;
  int Nb;
  int n_ring;
};

template <int N> struct GroupVec;
template <> struct GroupVec<2> { using type = float2; };
template <> struct GroupVec<4> { using type = float4; };

struct Smem {
  int part_off, gene_off, aop_off, tab_off, cnt_off, total;
};

__host__ __device__ constexpr Smem smem_layout(int H, bool velo, int nwarps, int n_ring) {
  Smem L{};
  int off = kHeader;
  L.part_off = off;  // cell partials: [park slot][group][warp][quantity][cell], park ring = table ring + 2 stages
  off += (n_ring + kSvcDist) * kGPS * nwarps * (velo ? 3 : 2) * 8 * 4;
  L.gene_off = off;  // per-gene parameters: [warp][row tile][2][grp] float4
  off += nwarps * 2 * 2 * 8 * 16;
  L.aop_off = off;  // forward A operands: [warp][row tile][k-step][main / cross][lane] uint4
  off += nwarps * 2 * ksteps(H) * 2 * 32 * 16;
  L.tab_off = off;
  off += n_ring * kGPS * table_group_floats(H, velo) * 4;
  off = (off + 127) / 128 * 128;
  L.cnt_off = off;  // [depth][matrix][cell q / q+4][thread] x 16 B
  off += kD * (velo ? 2 : 1) * 2 * (32 * nwarps) * 16;
  L.total = off;
  return L;
}

// Depth of the table ring: as many stages as the shared memory of one CTA per SM allows -- a function of the template
// arguments only, so the ring arithmetic of the kernel is compile-time.
constexpr int kSmemBudget = 226 * 1024;  // of the 227 KB opt-in dynamic shared memory per CTA on sm_100
__host__ __device__ constexpr int ring_depth(int H, bool velo) {
  for (int r = kMaxNS; r > 2; --r)
    if (smem_layout(H, velo, kThreads / 32, r).total <= kSmemBudget) return r;
  return 2;
}

template <int H, bool VELO, bool GRAD>
__global__ void __launch_bounds__(kThreads, 1) vcb_stream2_kernel(const Params P) {
  constexpr int K = 2 * H + 1;
  constexpr int KS = ksteps(H);
  constexpr int NT = 2;  // MMA row tiles per warp (32 genes)
  constexpr int NMAT = VELO ? 2 : 1;
  constexpr int NQ = VELO ? 3 : 2;
  constexpr int R = kGroupCells;
  constexpr int D = kD;
  constexpr int NLD = NMAT * 2;
  constexpr int TABG = table_group_floats(H, VELO);
  constexpr int TAIL = table_tail(H, VELO);
  constexpr bool NEED_D = GRAD || VELO;
  constexpr bool NEED_E = GRAD && VELO;
  constexpr uint32_t kStageTabBytes = (uint32_t)(kGPS * TABG * 4);

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  constexpr int nthr = kThreads, nwarps = nthr >> 5;
  constexpr uint32_t kSlotBytes = (uint32_t)nthr * 16u;  // distance between two of a thread's count slots
  const int warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 2, q = lane & 3;
  constexpr int NS = ring_depth(H, VELO);
  const int tile = blockIdx.x, split = blockIdx.y;
  constexpr int tile_genes = 32 * nwarps;
  const long long g_base = (long long)tile * tile_genes;
  const long long rem = P.ld - g_base;
  const int W = (int)(rem < (long long)tile_genes ? rem : (long long)tile_genes);
  constexpr Smem L = smem_layout(H, VELO, nwarps, NS);
  const uint32_t full0 = smem_u32(smem_raw);
  const uint32_t done0 = full0 + 8 * kMaxNS;
  float* const s_part = reinterpret_cast<float*>(smem_raw + L.part_off);
  float* const s_tab = reinterpret_cast<float*>(smem_raw + L.tab_off);

  // this CTA's stages
  const long long T0 = (P.n_stages_total * split) / P.n_split;
  const long long T1 = (P.n_stages_total * (split + 1)) / P.n_split;
  const int n_stages = (int)(T1 - T0);
  const long long G0 = T0 * kGPS;  // first 8-cell group

  // ---- per-gene constants -> shared memory ---------------------------------------------------------------
  // gene j in 0..3 of the lane: g = g_base + warp*32 + 4*grp + j; row tile mt = j>>1 holds it in fragment row grp (j even)
  // or grp+8 (j odd).
  const int gl = warp * 32 + 4 * grp;
  const bool gvalid = gl < W;
  auto gene_of = [&](int mt, int odd) -> long long { return g_base + gl + 2 * mt + odd; };
  auto a_value = [&](long long g, int slot) -> float {  // forward A operand, already in base-2 units
    if (g >= P.Ng || slot == 0 || slot > K) return 0.f;
    return slot == K ? kLog2e : P.nu[g * K + slot] * kLog2e;
  };
  auto const_term = [&](long long g, int b) -> float {  // (nu_0 - ln r + batch offset) log2 e; a padding gene contributes nothing
    if (g >= P.Ng) return -1e30f;
    float v = P.nu[g * K] + logf(P.shape_inv[g]);
    if (b >= 0) v += P.dnu[(long long)b * P.Ng + g];
    return v * kLog2e;
  };
  uint4* const s_aop = reinterpret_cast<uint4*>(smem_raw + L.aop_off) + (size_t)warp * (NT * KS * 2 * 32) + lane;
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
        gs[o] = ((ok && VELO) ? P.gamma[g] : 1.f) + kRelEps;
        ibs[o] = (ok && VELO) ? expf(-P.logbeta[g]) : 1.f;
      }
      s_gene[(mt * 2 + 0) * 8] = make_float4(-rs[0], -rs[1], const_term(gene_of(mt, 0), -1), const_term(gene_of(mt, 1), -1));
      s_gene[(mt * 2 + 1) * 8] = make_float4(gs[0], gs[1], ibs[0], ibs[1]);
    }
  }
  __syncwarp();
#ifdef VCB_EXP_REGA
  // forward A operands and per-gene parameters held in registers (needs the 255-register configuration)
  uint4 rA0[NT][KS], rA1[NT][KS];
  float4 rGa[NT], rGb[NT];
  auto reload_gene_regs = [&]() {
#pragma unroll
    for (int mt = 0; mt < NT; ++mt) {
      rGa[mt] = s_gene[(mt * 2 + 0) * 8];
      rGb[mt] = s_gene[(mt * 2 + 1) * 8];
    }
  };
#pragma unroll
  for (int mt = 0; mt < NT; ++mt)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      rA0[mt][ks] = s_aop[((mt * KS + ks) * 2 + 0) * 32];
      rA1[mt][ks] = s_aop[((mt * KS + ks) * 2 + 1) * 32];
    }
  reload_gene_regs();
#endif
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
#ifdef VCB_EXP_REGA
    reload_gene_regs();
#endif
  };
  // -2 = no batch selected yet (a mixed stage is stamped -1 and must never compare equal); without batches the table kernel
  // stamps every stage with 0
  int cur_b = P.Nb > 0 ? -2 : 0;

  const float2 zero2 = f2s(0.f);
  // per gene pair: sum kS (y - LS), sum LS, sum kU (y + lg2 mb - LU), sum LU, sum gU   (base-2 units; folded at the end)
  float2 accKS[NT], accLS[NT], accKU[NT], accLU[NT], accGU[NT];
  float accNu[NT][KS][4];
#pragma unroll
  for (int mt = 0; mt < NT; ++mt) {
    accKS[mt] = accLS[mt] = accKU[mt] = accLU[mt] = accGU[mt] = zero2;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
      for (int i = 0; i < 4; ++i) accNu[mt][ks][i] = 0.f;
  }
  // batch boundary: the constant column of d/dnu (slot 0: lanes q == 0, c0 / c2) is the batch's d/dDelta-nu.  Every
  // (split, gene) has exactly one owner, so a plain read-modify-write is race free and the result reproducible.
  auto flush_batch = [&]() {
    if (q == 0 && cur_b >= 0 && P.Nb > 0) {
#pragma unroll
      for (int mt = 0; mt < NT; ++mt)
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const long long g = gene_of(mt, o);
          if (g < P.ld) {
            float* dst = P.dnupart + ((long long)split * P.Nb + cur_b) * P.ld + g;
            *dst += accNu[mt][0][2 * o];
          }
          accNu[mt][0][2 * o] = 0.f;
        }
    }
  };

  // ---- count ring: warp-private cp.async pipeline (layout and loader mapping as in vcb_stream.cuh) -----------
  // Loader lane (row l_row = lane/8 of each half group, 16-byte chunk lane%8) copies into the slot of the lane that will
  // consume the data.  Rows outside the matrix and gene chunks past the row pitch are zero-filled by the copy itself
  // (src-size 0: nothing is read, so the source address needs no clamping): `rows_left` counts the matrix rows from this
  // lane's row of the next group on, and is pinned to 0 for a lane without genes.
  const int l_row = lane >> 3, l_chunk = lane & 7;
  // Slot of (row r in 0..3, 16-byte gene chunk c in 0..7) inside a warp's 512-byte block: r*8 + (c ^ 2r).  A loader quarter-warp
  // (one row, chunks 0..7) and a consumer quarter-warp (rows 0..3 x two chunks) both touch eight different 16-byte bank
  // groups: neither the copy's shared-memory write nor the consumer's LDS.128 has a bank conflict.  (With the plain
  // consumer-order layout the copy wrote at a 64-byte stride: a 4-way conflict on every LDGSTS.)
  const uint32_t s_cnt_ld = smem_u32(reinterpret_cast<float4*>(smem_raw + L.cnt_off) + (warp * 32 + l_row * 8 + (l_chunk ^ (2 * l_row))));
  const int lgl = warp * 32 + 4 * l_chunk;
  const char* ldS;
  long long u_minus_s = 0;
  int rows_left;
  {
    const long long off = (G0 * R + l_row) * P.ld + g_base + lgl;
    ldS = reinterpret_cast<const char*>(P.S + off);
    if (VELO) u_minus_s = reinterpret_cast<const char*>(P.U) - reinterpret_cast<const char*>(P.S);
    const long long c_end = T1 * kStageCells < P.Nc ? T1 * kStageCells : P.Nc;  // this CTA's rows end here
    long long rl = c_end - (G0 * R + l_row);
    rl = rl < 0 ? 0 : (rl > (1ll << 30) ? (1ll << 30) : rl);
    rows_left = lgl < W ? (int)rl : 0;
  }
  auto load_counts = [&](int d) {  // the next group (in order) into depth slot d; past the CTA's range the copies are all zero-fill
    uint32_t dst = s_cnt_ld + (uint32_t)d * (NLD * kSlotBytes);
    const uint32_t sz0 = rows_left > 0 ? 16u : 0u, sz1 = rows_left > 4 ? 16u : 0u;
    const char* hi = ldS + 16 * P.ld;  // row r + 4
    cp_async16(dst, ldS, sz0);
    cp_async16(dst + kSlotBytes, hi, sz1);
    if (VELO) {
      cp_async16(dst + 2 * kSlotBytes, ldS + u_minus_s, sz0);
      cp_async16(dst + 3 * kSlotBytes, hi + u_minus_s, sz1);
    }
    ldS += 4ll * R * P.ld;
    rows_left = rows_left > R ? rows_left - R : 0;
    cp_async_commit();
  };

  // ---- table ring ------------------------------------------------------------------------------------------
  //   full[s] : mbarrier, the two operand tables of the stage in slot s have landed  (arrive.expect_tx + complete_tx)
  //   done[s] : mbarrier, every warp has left the stage in slot s (its tables and the parked cell partials are final)
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(done0 + 8 * s, nwarps);
    }
    mbar_fence_init();
  }
  __syncthreads();
  auto issue_stage = [&](int stage, int slot) {  // one elected lane
    mbar_expect_tx(full0 + 8 * slot, kStageTabBytes);
    bulk_g2s(smem_u32(s_tab) + (uint32_t)slot * kStageTabBytes, P.tab + (T0 + stage) * (long long)(kGPS * TABG), kStageTabBytes,
             full0 + 8 * slot);
  };
  if (tid == 0)
    for (int s = 0; s < NS && s < n_stages; ++s) issue_stage(s, s);
  float* const cellpart_t = P.cellpart + (long long)tile * P.Ncp * NQ;
  // Cell partials of a group: a warp reduces its NQ x 8 per-cell sums over its 32 genes (three shuffle levels) and parks
  // them, [park slot][warp][quantity][cell][group of the stage].  Stage x is SERVICED by warp (x + 2) % nwarps at the top
  // of its own stage x + 2: it waits for done[slot of x] (every warp has left stage x -- they are normally long past it, so
  // the wait returns at once and nobody spins), re-issues the table copy of that slot for stage x + NS, adds the 16 warps'
  // parked terms in a fixed order (deterministic) and stores them.  A warp can be at most NS - 1 stages ahead of the slowest
  // one and the servicing warp two stages behind the stage it serves: the park ring has NS + 2 slots.
  constexpr int NP = NS + kSvcDist;
  auto service = [&](int x) {
    const int xs = x % NS;
    mbar_wait(done0 + 8 * xs, (uint32_t)((x / NS) & 1));
    if (lane == 0 && x + NS < n_stages) issue_stage(x + NS, xs);
    if (GRAD && lane < R * NQ) {
      const int i = lane >> 3, cell = lane & 7;
      using VecG = typename GroupVec<kGPS>::type;  // one (quantity, cell) of every group of the stage
      const VecG* src = reinterpret_cast<const VecG*>(s_part) + (size_t)(x % NP) * nwarps * (NQ * 8) + lane;
      float s[kGPS];
#pragma unroll
      for (int g = 0; g < kGPS; ++g) s[g] = 0.f;
#pragma unroll
      for (int w = 0; w < nwarps; ++w) {
        const VecG v = src[w * (NQ * 8)];
        const float* vf = reinterpret_cast<const float*>(&v);
#pragma unroll
        for (int g = 0; g < kGPS; ++g) s[g] += vf[g];
      }
      float* dst = cellpart_t + (T0 + x) * (long long)(kGPS * R * NQ) + cell * NQ + i;
#pragma unroll
      for (int g = 0; g < kGPS; ++g) dst[g * R * NQ] = s[g];
    }
  };
#pragma unroll
  for (int s = 0; s < D; ++s) load_counts(s);

  const float2 one2 = f2s(1.f), neg1 = f2s(-1.f), eps2 = f2s(kRelEps);
  const float4* const s_cnt = reinterpret_cast<const float4*>(smem_raw + L.cnt_off) + (warp * 32 + q * 8 + (grp ^ (2 * q)));

  // One 8-cell group of one batch.  MASKED: compile-time copy used for stages whose cells belong to several batches;
  // cells outside batch `pass_b` are switched off (their counts read as zero, their exponent as -inf).
  auto process = [&](auto masked_tag, const float* tb, const float4* cnt, const int pass_b, float (&pcf)[2], float (&pphi)[2],
                     float (&pom)[2]) {
    constexpr bool MASKED = decltype(masked_tag)::value;
    const float4* tb4 = reinterpret_cast<const float4*>(tb);
    float om[2] = {0.f, 0.f};
    if (VELO) {
      om[0] = tb[TAIL + q];
      om[1] = tb[TAIL + q + 4];
    }
    float moff[2] = {0.f, 0.f}, mk[2] = {1.f, 1.f};
    if (MASKED) {
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const bool in = __float_as_int(tb[TAIL + 8 + q + 4 * cc]) == pass_b;
        moff[cc] = in ? 0.f : -60000.f;
        mk[cc] = in ? 1.f : 0.f;
      }
    }
    float2 pcf2[2] = {zero2, zero2};
    const float2* cnt2 = reinterpret_cast<const float2*>(cnt);
#pragma unroll
    for (int mt = 0; mt < NT; ++mt) {
      // ---- forward contractions: y (base-2 exponent without the constant), d log2e, omega d'' log2e ----------
      float Ce[4] = {0.f, 0.f, 0.f, 0.f}, Cd[4] = {0.f, 0.f, 0.f, 0.f}, Cw[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
#ifdef VCB_EXP_REGA
        const uint4 a0 = rA0[mt][ks], a1 = rA1[mt][ks];
#else
        const uint4 a0 = s_aop[((mt * KS + ks) * 2 + 0) * 32], a1 = s_aop[((mt * KS + ks) * 2 + 1) * 32];
#endif
        const uint32_t am[4] = {a0.x, a0.y, a0.z, a0.w}, ax[4] = {a1.x, a1.y, a1.z, a1.w};
#ifdef VCB_EXP_NOFWD
        Ce[0] = __uint_as_float(am[0]) * tb[q]; Ce[1] = __uint_as_float(am[1]) * tb[q]; Ce[2] = __uint_as_float(am[2]) * tb[q]; Ce[3] = __uint_as_float(am[3]) * tb[q];
        Cd[0] = __uint_as_float(ax[0]) * tb[q]; Cd[1] = __uint_as_float(ax[1]) * tb[q]; Cd[2] = __uint_as_float(ax[2]) * tb[q]; Cd[3] = __uint_as_float(ax[3]) * tb[q];
        Cw[0] = Ce[1]; Cw[1] = Cd[2]; Cw[2] = Ce[3]; Cw[3] = Cd[0];
#else
        mma_split_fwd(Ce, am, ax, tb4[(SEC_F0 * KS + ks) * 32 + lane]);
        if (NEED_D) mma_split_fwd(Cd, am, ax, tb4[(SEC_F1 * KS + ks) * 32 + lane]);
        if (NEED_E) mma_split_fwd(Cw, am, ax, tb4[(SEC_F2 * KS + ks) * 32 + lane]);
#endif
      }
#ifdef VCB_EXP_REGA
      const float4 ga = rGa[mt];
#else
      const float4 ga = s_gene[(mt * 2 + 0) * 8];
#endif
      const float2 nr_mt = f2(ga.x, ga.y);
      float2 gam_mt = one2, invb_mt = one2;
      if (VELO) {
#ifdef VCB_EXP_REGA
        const float4 gb = rGb[mt];
#else
        const float4 gb = s_gene[(mt * 2 + 1) * 8];
#endif
        gam_mt = f2(gb.x, gb.y);
        invb_mt = f2(gb.z, gb.w);
      }
      float2 Gg[2], Gw[2];
#ifdef VCB_EXP_NOELEM
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const float2 kS = cnt2[((size_t)cc * nthr) * 2 + mt], kU = cnt2[((size_t)(2 + cc) * nthr) * 2 + mt];
        Gg[cc] = add2(f2(Ce[cc], Ce[2 + cc]), kS);
        Gw[cc] = add2(f2(Cd[cc] + Cw[cc], Cd[2 + cc] + Cw[2 + cc]), kU);
        accKS[mt] = add2(accKS[mt], Gg[cc]);
        pcf[cc] += Gw[cc].x;
      }
      if (false)
#endif
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        float2 kS = cnt2[((size_t)cc * nthr) * 2 + mt];
        float2 kU = VELO ? cnt2[((size_t)(2 + cc) * nthr) * 2 + mt] : kS;
        float2 y = f2(Ce[cc] + ga.z, Ce[2 + cc] + ga.w);
        if (MASKED) {
          y = add2(y, f2s(moff[cc]));
          kS = mul2(kS, f2s(mk[cc]));
          kU = mul2(kU, f2s(mk[cc]));
        }
        const float2 u = ex2_2(y);
        const float2 s = add2(u, one2);
        const float2 LS = lg2_2(s);
        accKS[mt] = fma2(kS, fma2(LS, neg1, y), accKS[mt]);
        accLS[mt] = add2(accLS[mt], LS);
        float2 g = zero2, w = zero2;
        if (VELO) {
          const float2 a = f2(fmaf(Cd[cc], om[cc], gam_mt.x), fmaf(Cd[2 + cc], om[cc], gam_mt.y));  // a + eps
          const float2 m = f2(fmaxf(a.x, kRelEps), fmaxf(a.y, kRelEps));
          const float2 mb = mul2(m, invb_mt);
          const float2 uU = mul2(u, mb);
          const float2 sU = add2(uU, one2);
          const float2 LU = lg2_2(sU);
          const float2 lmb = lg2_2(mb);
          accKU[mt] = fma2(kU, fma2(LU, neg1, add2(y, lmb)), accKU[mt]);
          accLU[mt] = add2(accLU[mt], LU);
          if (GRAD) {
            const float2 sUm = mul2(sU, m);
            const float2 rc = rcp_2(mul2(s, sUm));  // one MUFU for 1/s and 1/(sU m)
            const float2 inv_s = mul2(rc, sUm), inv_sUm = mul2(rc, s);
            const float2 gS = mul2(fma2(nr_mt, u, kS), inv_s);
            const float2 w0 = mul2(fma2(nr_mt, uU, kU), inv_sUm);
            const float2 gU = mul2(w0, m);
            w = mul2(w0, f2(a.x > kRelEps ? 1.f : 0.f, a.y > kRelEps ? 1.f : 0.f));
            g = add2(gS, gU);
            accGU[mt] = add2(accGU[mt], gU);
            pom[cc] = fmaf(w.x, Cd[cc], pom[cc]);
            pom[cc] = fmaf(w.y, Cd[2 + cc], pom[cc]);
            pphi[cc] = fmaf(w.x, Cw[cc], pphi[cc]);
            pphi[cc] = fmaf(w.y, Cw[2 + cc], pphi[cc]);
          }
        } else if (GRAD) {
          g = mul2(fma2(nr_mt, u, kS), rcp_2(s));
        }
        if (GRAD) {
          pcf2[cc] = mt == 0 ? g : add2(pcf2[cc], g);
          pphi[cc] = fmaf(g.x, Cd[cc], pphi[cc]);
          pphi[cc] = fmaf(g.y, Cd[2 + cc], pphi[cc]);
        }
        Gg[cc] = g;
        Gw[cc] = w;
      }
      // ---- backward contractions, 3xTF32: acc[gene][slot] = sum_cells G[gene][cell] Z[cell][slot] ----------------
#ifdef VCB_EXP_NOBWD
      if (GRAD) {
        accNu[mt][0][0] += Gg[0].x + Gw[0].x; accNu[mt][0][1] += Gg[0].y + Gw[0].y; accNu[mt][0][2] += Gg[1].x + Gw[1].x; accNu[mt][0][3] += Gg[1].y + Gw[1].y;
      }
      if (false) {
#else
      if (GRAD) {
#endif
        auto split3 = [&](const float2 v0, const float2 v1, uint32_t (&hi)[4], uint32_t (&lo)[4]) {
          hi[0] = __float_as_uint(v0.x) & 0xffffe000u;
          hi[1] = __float_as_uint(v0.y) & 0xffffe000u;
          hi[2] = __float_as_uint(v1.x) & 0xffffe000u;
          hi[3] = __float_as_uint(v1.y) & 0xffffe000u;
          const float2 l0 = __fadd2_rn(v0, make_float2(-__uint_as_float(hi[0]), -__uint_as_float(hi[1])));
          const float2 l1 = __fadd2_rn(v1, make_float2(-__uint_as_float(hi[2]), -__uint_as_float(hi[3])));
          lo[0] = __float_as_uint(l0.x);
          lo[1] = __float_as_uint(l0.y);
          lo[2] = __float_as_uint(l1.x);
          lo[3] = __float_as_uint(l1.y);
        };
        uint32_t gh[4], gl_[4], wh[4], wl[4];
        split3(Gg[0], Gg[1], gh, gl_);
        if (VELO) split3(Gw[0], Gw[1], wh, wl);
#pragma unroll
        for (int nt = 0; nt < KS; ++nt) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
          const float4 bz = tb4[(SEC_B0 * KS + nt) * 32 + lane];  // {Zhi b0, Zhi b1, Zlo b0, Zlo b1}
          // two independent accumulation chains (the g terms and the w terms): a dependent mma.sync costs ~35 cycles, a
          // single six-deep chain per row tile was 15 % of the kernel
          float acw[4] = {0.f, 0.f, 0.f, 0.f};
          float4 bw = make_float4(0.f, 0.f, 0.f, 0.f);
          if (VELO) bw = tb4[(SEC_B1 * KS + nt) * 32 + lane];
          mma_tf32(acc, gl_, __float_as_uint(bz.x), __float_as_uint(bz.y));
          if (VELO) mma_tf32(acw, wl, __float_as_uint(bw.x), __float_as_uint(bw.y));
          mma_tf32(acc, gh, __float_as_uint(bz.z), __float_as_uint(bz.w));
          if (VELO) mma_tf32(acw, wh, __float_as_uint(bw.z), __float_as_uint(bw.w));
          mma_tf32(acc, gh, __float_as_uint(bz.x), __float_as_uint(bz.y));
          if (VELO) mma_tf32(acw, wh, __float_as_uint(bw.x), __float_as_uint(bw.y));
          // packed adds on the accumulator pairs (c0,c1), (c2,c3): with scalar adds the register allocator scattered accNu over
          // odd registers and re-paired every FFMA2 operand around them with MOVs (97 -> 22 MOVs per stage, -6 % time)
#pragma unroll
          for (int i = 0; i < 4; i += 2) {
            float2 t = f2(acc[i], acc[i + 1]);
            if (VELO) t = add2(t, f2(acw[i], acw[i + 1]));
            const float2 r2 = add2(f2(accNu[mt][nt][i], accNu[mt][nt][i + 1]), t);
            accNu[mt][nt][i] = r2.x;
            accNu[mt][nt][i + 1] = r2.y;
          }
        }
      }
    }
    if (GRAD) {
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {  // a masked group makes one pass per batch present: those accumulate
        const float t = pcf2[cc].x + pcf2[cc].y;
        pcf[cc] = MASKED ? pcf[cc] + t : t;
      }
    }
  };

  // ---- main loop ---------------------------------------------------------------------------------------
  int c_slot = 0, c_phase = 0;  // table slot / parity of the current stage
  int p_slot = 0;               // park slot of the current stage
  // one 8-cell group of the current stage: consume, reduce the per-cell sums over the warp's genes, park them, refill the
  // count slot
  int c_ring = 0;  // count-ring slot of the next group to consume (and, once consumed, to refill)
  auto do_group = [&](auto masked_tag, const int h, const float* tb_stage, float* part_stage) {
    constexpr bool MASKED = decltype(masked_tag)::value;
    cp_async_wait<D - 1>();
    __syncwarp();
    const float* tb = tb_stage + h * TABG;
    const float4* cnt = s_cnt + (size_t)c_ring * NLD * nthr;
    float pcf[2] = {0.f, 0.f}, pphi[2] = {0.f, 0.f}, pom[2] = {0.f, 0.f};
    if (MASKED) {
      // mixed stage: one masked pass per batch present among the 16 cells (warp-uniform decisions)
      const int my_id = __float_as_int(tb_stage[((lane >> 3) & (kGPS - 1)) * TABG + TAIL + 8 + (lane & 7)]);
      for (int b = 0; b < P.Nb; ++b) {
        if (!__any_sync(0xffffffffu, my_id == b)) continue;
        if (b != cur_b) {
          if (GRAD) flush_batch();
          cur_b = b;
          set_batch(b);
        }
        process(BoolTag<true>{}, tb, cnt, b, pcf, pphi, pom);
      }
    } else {
      process(BoolTag<false>{}, tb, cnt, 0, pcf, pphi, pom);
    }
    if (GRAD) {
      // the lane holds partial sums for the cells q and q+4: swap halves so that lanes 0-15 keep cell q and lanes
      // 16-31 cell q+4, add the four gene groups of each half, and park 8 values per quantity
      const bool up = (lane & 16) != 0;
      float v0 = (up ? pcf[1] : pcf[0]) + __shfl_xor_sync(0xffffffffu, up ? pcf[0] : pcf[1], 16);
      float v1 = (up ? pphi[1] : pphi[0]) + __shfl_xor_sync(0xffffffffu, up ? pphi[0] : pphi[1], 16);
      float v2 = VELO ? (up ? pom[1] : pom[0]) + __shfl_xor_sync(0xffffffffu, up ? pom[0] : pom[1], 16) : 0.f;
      v0 += __shfl_xor_sync(0xffffffffu, v0, 4);
      v1 += __shfl_xor_sync(0xffffffffu, v1, 4);
      if (VELO) v2 += __shfl_xor_sync(0xffffffffu, v2, 4);
      v0 += __shfl_xor_sync(0xffffffffu, v0, 8);
      v1 += __shfl_xor_sync(0xffffffffu, v1, 8);
      if (VELO) v2 += __shfl_xor_sync(0xffffffffu, v2, 8);
      if ((lane & 12) == 0) {
        part_stage[h] = v0;
        part_stage[8 * kGPS + h] = v1;
        if (VELO) part_stage[16 * kGPS + h] = v2;
      }
    }
    __syncwarp();  // every lane has read its counts: the slot may be refilled
    load_counts(c_ring);
    c_ring = c_ring + 1 == D ? 0 : c_ring + 1;
  };
  // Leave stage st: the tables of the slot and this warp's parked partials are final (the __syncwarp in do_group orders the
  // other lanes' accesses before lane 0's arrival); then the rotating service duty for the stage that starts next.
  auto leave_stage = [&](const int st) {
    if (lane == 0) mbar_arrive(done0 + 8 * c_slot);
    if (++p_slot == NP) p_slot = 0;
    if (++c_slot == NS) {
      c_slot = 0;
      c_phase ^= 1;
    }
    const int nx = st + 1;
    if (nx < n_stages && nx >= kSvcDist && (nx - warp) % nwarps == 0) service(nx - kSvcDist);
  };
  for (int st = 0; st < n_stages; ++st) {
    mbar_wait(full0 + 8 * c_slot, (uint32_t)c_phase);
    const float* tb_stage = s_tab + (size_t)c_slot * (kGPS * TABG);
    // park address of this lane's cell: [p_slot][warp][quantity][cell][group]
    float* part_stage = s_part + (((size_t)p_slot * nwarps + warp) * (NQ * 8) + (q + 4 * (lane >> 4))) * kGPS;
    const int stage_b = __float_as_int(tb_stage[TAIL + 16]);  // the stage's batch (0 without batches), -1 if its cells disagree
    if (stage_b != cur_b && stage_b >= 0) {
      if (GRAD) flush_batch();
      cur_b = stage_b;
      set_batch(cur_b);
    }
    if (stage_b >= 0) {
#pragma unroll
      for (int h = 0; h < kGPS; ++h) do_group(BoolTag<false>{}, h, tb_stage, part_stage);
    } else {
      for (int h = 0; h < kGPS; ++h) do_group(BoolTag<true>{}, h, tb_stage, part_stage);
    }
    leave_stage(st);
  }
  cp_async_wait<0>();
  // the last stages have no successor stage to be serviced from: same rotation, after the loop
  for (int x = (n_stages >= kSvcDist ? n_stages - kSvcDist : 0); x < n_stages; ++x)
    if ((x + kSvcDist - warp + nwarps) % nwarps == 0) service(x);

  // ---- flush per-gene partial sums -------------------------------------------------------------------------
  if (GRAD) flush_batch();
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
    const long long g0 = gene_of(mt, 0);
    auto put = [&](int row, float2 v) {
      v = lane_sum4(v);
      if (gvalid && q == 0) *reinterpret_cast<float2*>(gp + (long long)row * P.ld + g0) = v;
    };
    const float4 ga = s_gene[(mt * 2 + 0) * 8];
    const float2 nr_mt = f2(ga.x, ga.y);
    put(ROW_AS, fma2(nr_mt, accLS[mt], accKS[mt]));                    // sum kS (y - LS) - r sum LS
    put(ROW_LS, VELO ? add2(accLS[mt], accLU[mt]) : accLS[mt]);
    if (VELO) {
      put(ROW_AU, fma2(nr_mt, accLU[mt], accKU[mt]));
      if (GRAD) put(ROW_GU, accGU[mt]);
    }
    if (GRAD && gvalid) {
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

}  // namespace s2
}  // namespace vcb
