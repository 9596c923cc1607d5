// The streaming kernel: one pass over the [cells][genes] count tiles of S (and U) that produces the
// per-gene log-prob partial sums AND every gradient partial sum.
//
// Mapping (genes are the contiguous axis of the counts, preprocessing.py:193-194):
//   * a CTA owns a gene tile of 2*NP*blockDim genes and a contiguous range of cells;
//   * a thread owns NP packed pairs of adjacent genes: their Fourier coefficients, dispersion and
//     kinetics stay in registers for the whole kernel, as do the per-gene accumulators (no atomics on
//     the hot path).  All fp32 arithmetic is issued as packed f32x2 (FFMA2/FADD2/FMUL2) over a gene
//     pair: the fp32 pipe, not HBM, bounds this kernel, so halving issue slots matters;
//   * cells are streamed through a shared-memory ring filled by the TMA engine with 1-D bulk copies
//     (cp.async.bulk + mbarrier complete_tx); a stage = kCellsPerStage count rows of S and U plus the
//     per-cell table rows (Fourier basis, its derivatives, omega, size factor -- each stored twice so a
//     broadcast LDS.128 yields ready-made {z,z} operands -- and the batch id) built by
//     vcb_cell_tables_kernel; counts are read with one conflict-free LDS.128 per matrix;
//   * per-cell sums (d/dphi, d/dcf, d/domega run over genes, i.e. across threads) are reduced inside each
//     warp once per stage with a transposed butterfly (reduce-scatter over the stage's cells, ~2 shuffles
//     per value instead of 5) and written as per-warp partials; the cell epilogue kernel adds the warps.
//     No warp ever waits for another warp: the only cross-warp state is the ring's done[] counters that
//     tell the producer thread when a slot may be refilled.
//
// Arithmetic per (cell, gene), SURVEY.md Appendix A, in units of mu/r and base-2 logs so that every
// transcendental is a single MUFU op and the n r log r terms cancel analytically:
//   y = (etaS - ln r) log2e; u = 2^y = muS/r; s = 1+u; LS = lg2 s; gS = (kS - r u)/s
//   a = d*omega+gamma; m = relu(a)+1e-5; mb = m/beta; uU = u mb; sU = 1+uU; LU = lg2 sU
//   w0 = (kU - r uU)/(sU m); gU = w0 m; w = 1[a>0] w0
//   log-prob pieces: kS (y-LS), LS, kU (y+lg2 mb-LU), LU   (times ln2, plus per-gene terms, in the epilogue)
#pragma once
#include "vcb_common.cuh"

namespace vcb {

constexpr int kCellsPerStage = 4;  // R (power of two <= 32: the warp reduce-scatter splits lanes by cell)
constexpr int kStages = 6;         // ring depth (6 x 32 KB of counts in flight per SM at 1024-gene tiles)
// A thread owns NP packed gene pairs.  NP = 1: up to 512 threads/CTA at <=128 registers (16 warps/SM);
// NP = 2: up to 256 threads/CTA at <=255 registers (8 warps/SM, half the per-cell table traffic per gene).
__host__ __device__ constexpr int max_threads(int NP) { return NP == 1 ? 512 : 256; }

// per-gene partial rows written by the streaming kernel
enum GeneRow { ROW_AS = 0, ROW_LS = 1, ROW_AU = 2, ROW_LU = 3, ROW_GU = 4, ROW_W = 5, ROW_PSI = 6, ROW_DNU = 7 };

// per-cell table row, in floats: pairs {z,z} for zeta[1..2H], zeta'[1..2H], zeta''[1..2H], omega, cf; then batch id
__host__ __device__ constexpr int table_width(int H) { return ((12 * H + 5) + 3) / 4 * 4; }
__host__ __device__ constexpr int gene_rows(int H) { return ROW_DNU + 2 * H + 1; }

struct StreamParams {
  const float* S;
  const float* U;
  const float* tab;  // [Nc][TABW]
  const float* nu;
  const float* dnu;
  const float* shape_inv;
  const float* logbeta;
  const float* gamma;
  float* genepart;  // [n_split][rows][ld]
  float* cellpart;  // [n_tiles * warps_per_cta][Nc][NQ]
  float* d_dnu;     // [Nb][Ng], zeroed, atomically accumulated at batch boundaries
  long long Nc, Ng, ld;
  int n_split;
  int Nb;
  int debug_skip_compute;  // profiling aid: stream the tiles but skip the arithmetic
};

struct StreamSmem {
  int tab_off, cnt_off, total;  // byte offsets inside dynamic shared memory
};

__host__ __device__ inline StreamSmem stream_smem_layout(int H, bool velo, bool grad, int nthr, int W) {
  StreamSmem L;
  int off = 128;  // mbarriers
  L.tab_off = off;
  off += kStages * kCellsPerStage * table_width(H) * 4;
  off = (off + 127) / 128 * 128;
  L.cnt_off = off;
  off += kStages * (velo ? 2 : 1) * kCellsPerStage * W * 4;
  off = (off + 127) / 128 * 128;
  (void)grad;
  (void)nthr;
  L.total = off;
  return L;
}

// ---- packed helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 ex2_2(float2 a) { return f2(ex2_approx(a.x), ex2_approx(a.y)); }
__device__ __forceinline__ float2 lg2_2(float2 a) { return f2(lg2_approx(a.x), lg2_approx(a.y)); }
__device__ __forceinline__ float2 rcp_2(float2 a) { return f2(rcp_approx(a.x), rcp_approx(a.y)); }

template <int H, bool VELO, bool GRAD, bool LGINLINE, int NP>
__global__ void __launch_bounds__(max_threads(NP), 1) vcb_stream_kernel(const StreamParams P) {
  constexpr int GPT = 2 * NP;  // genes per thread
  constexpr int K = 2 * H + 1;
  constexpr int TABW = table_width(H);
  constexpr int NMAT = VELO ? 2 : 1;
  constexpr int NQ = VELO ? 3 : 2;
  constexpr int R = kCellsPerStage;
  constexpr int NS = kStages;
  constexpr int P_Z1 = 2 * H, P_Z2 = 4 * H, P_OM = 6 * H, P_CF = 6 * H + 1;  // pair indices in a table row

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  const int tile = blockIdx.x;
  const int split = blockIdx.y;
  const long long g_base = (long long)tile * GPT * nthr;
  const long long rem = P.ld - g_base;
  const int W = (int)(rem < (long long)GPT * nthr ? rem : (long long)GPT * nthr);
  const StreamSmem L = stream_smem_layout(H, VELO, GRAD, nthr, W);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* full = mbar;  // [NS] each; roles documented where they are initialised
  uint64_t* done = mbar + NS;
  float* s_tab = reinterpret_cast<float*>(smem_raw + L.tab_off);
  float* s_cnt = reinterpret_cast<float*>(smem_raw + L.cnt_off);

  // this CTA's cells
  const long long c0 = (P.Nc * split) / P.n_split;
  const long long c1 = (P.Nc * (split + 1)) / P.n_split;
  const int n_cells = (int)(c1 - c0);
  const int n_stages = (n_cells + R - 1) / R;

  // ---- per-gene state in registers (pair p holds genes gj+2p, gj+2p+1) ----------------------------
  const long long gj = g_base + (long long)GPT * tid;
  const bool has_data = GPT * tid < W;  // this thread's 16 bytes exist in the smem rows
  float2 nu[K][NP], nr[NP], gam[NP], invb[NP], nu0c[NP];
  {
    float nus[K][GPT], rs[GPT], gs[GPT], ibs[GPT], lnr[GPT];
#pragma unroll
    for (int j = 0; j < GPT; ++j) {
      const long long g = gj + j;
      if (has_data && g < P.Ng) {
#pragma unroll
        for (int k = 0; k < K; ++k) nus[k][j] = P.nu[g * K + k];
        rs[j] = 1.0f / P.shape_inv[g];
        gs[j] = VELO ? P.gamma[g] : 1.f;
        ibs[j] = VELO ? expf(-P.logbeta[g]) : 1.f;
      } else {  // padding column: eta = -inf makes every contribution exactly zero
#pragma unroll
        for (int k = 0; k < K; ++k) nus[k][j] = 0.f;
        nus[0][j] = -1e30f;
        rs[j] = 1.f;
        gs[j] = 1.f;
        ibs[j] = 1.f;
      }
      lnr[j] = logf(rs[j]);
    }
#pragma unroll
    for (int p = 0; p < NP; ++p) {
#pragma unroll
      for (int k = 0; k < K; ++k) nu[k][p] = f2(nus[k][2 * p], nus[k][2 * p + 1]);
      nr[p] = f2(-rs[2 * p], -rs[2 * p + 1]);
      gam[p] = f2(gs[2 * p], gs[2 * p + 1]);
      invb[p] = f2(ibs[2 * p], ibs[2 * p + 1]);
      nu0c[p] = f2(nus[0][2 * p] - lnr[2 * p], nus[0][2 * p + 1] - lnr[2 * p + 1]);
    }
  }
  int cur_b = -1;

  const float2 zero2 = f2s(0.f);
  float2 accAS[NP], accLS[NP], accAU[NP], accLU[NP], accGU[NP], accW[NP], accPsi[NP];
  float2 accNu[K][NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    accAS[p] = accLS[p] = accAU[p] = accLU[p] = accGU[p] = accW[p] = accPsi[p] = zero2;
#pragma unroll
    for (int k = 0; k < K; ++k) accNu[k][p] = zero2;
  }

  // ---- producer helper (lane 0 of the producer warp) ---------------------------------------------------------------
  auto issue_stage = [&](int st) {
    const int slot = st % NS;
    const long long cs = c0 + (long long)st * R;
    const int nv = (n_cells - st * R) < R ? (n_cells - st * R) : R;
    const uint32_t row_bytes = (uint32_t)W * 4u;
    const uint32_t bytes = (uint32_t)nv * (TABW * 4u + NMAT * row_bytes);
    mbar_expect_tx(&full[slot], bytes);
    bulk_g2s(s_tab + (size_t)slot * R * TABW, P.tab + cs * TABW, (uint32_t)nv * TABW * 4u, &full[slot]);
    float* dstS = s_cnt + (size_t)slot * NMAT * R * W;
    if ((long long)W == P.ld) {  // the tile spans whole rows: the nv rows are one contiguous block
      bulk_g2s(dstS, P.S + cs * P.ld, (uint32_t)nv * row_bytes, &full[slot]);
      if (VELO) bulk_g2s(dstS + (size_t)R * W, P.U + cs * P.ld, (uint32_t)nv * row_bytes, &full[slot]);
    } else {
      for (int rr = 0; rr < nv; ++rr) {
        bulk_g2s(dstS + (size_t)rr * W, P.S + (cs + rr) * P.ld + g_base, row_bytes, &full[slot]);
        if (VELO)
          bulk_g2s(dstS + (size_t)(R + rr) * W, P.U + (cs + rr) * P.ld + g_base, row_bytes, &full[slot]);
      }
    }
  };

  // Synchronisation: no CTA-wide barrier and no warp-to-warp waiting in steady state.
  //   full[s] : TMA bytes of the stage in slot s have landed            (1 arrival + complete_tx)
  //   done[s] : every warp finished reading slot s                      (one arrival per warp)
  // Refills are work-stolen: every warp polls (twice per stage, and while it waits for data) whether the
  // slot of the next stage to issue has been released, and the warp that wins the CAS on s_next issues
  // the copies.  All warps therefore run identical code and no warp is the designated straggler.
  // (A dedicated producer warp would cost a whole 4-warp register allocation unit: 544 threads are
  //  accounted as 640, which caps the consumers at 96 registers.)
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
  int* s_next = reinterpret_cast<int*>(mbar + 2 * NS);
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&done[s], (uint32_t)nwarps);
    }
    mbar_fence_init();
    int st = 0;
    for (; st < NS && st < n_stages; ++st) issue_stage(st);
    *s_next = st;
  }
  __syncthreads();
  auto poll_refill = [&]() {  // warp-uniform; never blocks
    const int v = *reinterpret_cast<volatile int*>(s_next);
    if (v >= n_stages) return;
    const int prev = v - NS;
    int ready = mbar_test_wait(&done[prev % NS], (uint32_t)((prev / NS) & 1)) ? 1 : 0;
    ready = __shfl_sync(0xffffffffu, ready, 0);
    if (!ready) return;
    int won = 0;
    if (lane == 0) won = (atomicCAS(s_next, v, v + 1) == v) ? 1 : 0;
    won = __shfl_sync(0xffffffffu, won, 0);
    if (won && lane == 0) issue_stage(v);
  };
  float* const cellpart_w = P.cellpart + ((long long)tile * nwarps + warp) * P.Nc * NQ;

  const float2 one2 = f2s(1.f), neg1 = f2s(-1.f), l2e = f2s(kLog2e), eps2 = f2s(1e-5f);

  // ---- main loop ----------------------------------------------------------------------------------
  for (int st = 0; st < n_stages; ++st) {
    const int slot = st % NS;
    {
      uint32_t spins = 0;
      while (!__all_sync(0xffffffffu, mbar_try_wait(&full[slot], (uint32_t)((st / NS) & 1)))) {
        poll_refill();
        if (++spins > (1u << 24)) __trap();
      }
    }
    const int nv = (n_cells - st * R) < R ? (n_cells - st * R) : R;
    const float* tabs = s_tab + (size_t)slot * R * TABW;
    const float* cntS = s_cnt + (size_t)slot * NMAT * R * W;
    float part[R * NQ];  // this thread's per-cell partial sums of the stage, [cell][q]
#pragma unroll
    for (int i = 0; i < R * NQ; ++i) part[i] = 0.f;

#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      if (rr >= nv || P.debug_skip_compute) break;
      // table row as broadcast float4 loads: T4[i] = {z_{2i}, z_{2i}, z_{2i+1}, z_{2i+1}}
      const float4* T4 = reinterpret_cast<const float4*>(tabs + rr * TABW);
      auto tpair = [&](int pi) -> float2 {  // pi is a compile-time constant after unrolling
        const float4 v = T4[pi >> 1];
        return (pi & 1) ? f2(v.z, v.w) : f2(v.x, v.y);
      };
      float pcf = 0.f, pphi = 0.f, pom = 0.f;
      if (has_data) {
        if (P.Nb > 0) {
          const int b = __float_as_int(tabs[rr * TABW + 12 * H + 4]);
          if (b != cur_b) {  // CTA-uniform: batch boundary (rare when samples are concatenated)
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              const long long g0 = gj + 2 * p;
              if (GRAD && cur_b >= 0) {
                if (g0 < P.Ng) atomicAdd(&P.d_dnu[(long long)cur_b * P.Ng + g0], accNu[0][p].x);
                if (g0 + 1 < P.Ng) atomicAdd(&P.d_dnu[(long long)cur_b * P.Ng + g0 + 1], accNu[0][p].y);
                accNu[0][p] = zero2;
              }
              const float o0 = (g0 < P.Ng) ? P.dnu[(long long)b * P.Ng + g0] : 0.f;
              const float o1 = (g0 + 1 < P.Ng) ? P.dnu[(long long)b * P.Ng + g0 + 1] : 0.f;
              nu0c[p] = f2(nu[0][p].x - logf(-nr[p].x) + o0, nu[0][p].y - logf(-nr[p].y) + o1);
            }
            cur_b = b;
          }
        }
        const float2 om2 = tpair(P_OM);
        const float2 cf2 = tpair(P_CF);

        // forward contraction: eta' = etaS - ln r, d = nu.zeta', d2 = nu.zeta''
        float2 eta[NP], d[NP], d2[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          eta[p] = add2(nu0c[p], cf2);
          d[p] = zero2;
          d2[p] = zero2;
        }
#pragma unroll
        for (int k = 1; k < K; ++k) {
          const float2 z = tpair(k - 1);
#pragma unroll
          for (int p = 0; p < NP; ++p) eta[p] = fma2(nu[k][p], z, eta[p]);
          if (GRAD || VELO) {
            const float2 z1 = tpair(P_Z1 + k - 1);
#pragma unroll
            for (int p = 0; p < NP; ++p) d[p] = (k == 1) ? mul2(nu[k][p], z1) : fma2(nu[k][p], z1, d[p]);
          }
          if (VELO && GRAD) {
            const float2 z2 = tpair(P_Z2 + k - 1);
#pragma unroll
            for (int p = 0; p < NP; ++p) d2[p] = (k == 1) ? mul2(nu[k][p], z2) : fma2(nu[k][p], z2, d2[p]);
          }
        }

        float2 kS[NP], kU[NP];
        {
          const float* rowS = cntS + (size_t)rr * W + GPT * tid;
          const float* rowU = cntS + (size_t)(R + rr) * W + GPT * tid;
          if (NP == 2) {
            const float4 a4 = *reinterpret_cast<const float4*>(rowS);
            kS[0] = f2(a4.x, a4.y);
            kS[NP - 1] = f2(a4.z, a4.w);
            if (VELO) {
              const float4 b4 = *reinterpret_cast<const float4*>(rowU);
              kU[0] = f2(b4.x, b4.y);
              kU[NP - 1] = f2(b4.z, b4.w);
            }
          } else {
            kS[0] = *reinterpret_cast<const float2*>(rowS);
            if (VELO) kU[0] = *reinterpret_cast<const float2*>(rowU);
          }
          if (!VELO) {
#pragma unroll
            for (int p = 0; p < NP; ++p) kU[p] = zero2;
          }
        }

        float2 gE[NP], gd[NP];
        float2 pcf2 = zero2, pphi2 = zero2, pom2 = zero2;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const float2 y = mul2(eta[p], l2e);
          const float2 u = ex2_2(y);
          const float2 s = add2(u, one2);
          const float2 LS = lg2_2(s);
          accAS[p] = fma2(kS[p], fma2(LS, neg1, y), accAS[p]);
          accLS[p] = add2(accLS[p], LS);
          float2 g = zero2;
          if (GRAD) g = mul2(fma2(nr[p], u, kS[p]), rcp_2(s));
          if (LGINLINE) {
            float2 psi, lg;
            lg.x = lgamma_terms_inline(-nr[p].x, kS[p].x, psi.x);
            lg.y = lgamma_terms_inline(-nr[p].y, kS[p].y, psi.y);
            accAS[p] = fma2(lg, l2e, accAS[p]);
            accPsi[p] = add2(accPsi[p], psi);
          }
          gd[p] = zero2;
          if (VELO) {
            const float2 a = fma2(d[p], om2, gam[p]);
            const float2 m = add2(f2(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f)), eps2);
            const float2 mb = mul2(m, invb[p]);
            const float2 uU = mul2(u, mb);
            const float2 sU = add2(uU, one2);
            const float2 LU = lg2_2(sU);
            const float2 lmb = lg2_2(mb);
            accAU[p] = fma2(kU[p], fma2(LU, neg1, add2(y, lmb)), accAU[p]);
            accLU[p] = add2(accLU[p], LU);
            if (LGINLINE) {
              float2 psi, lg;
              lg.x = lgamma_terms_inline(-nr[p].x, kU[p].x, psi.x);
              lg.y = lgamma_terms_inline(-nr[p].y, kU[p].y, psi.y);
              accAU[p] = fma2(lg, l2e, accAU[p]);
              accPsi[p] = add2(accPsi[p], psi);
            }
            if (GRAD) {
              const float2 w0 = mul2(fma2(nr[p], uU, kU[p]), rcp_2(mul2(sU, m)));
              const float2 gU = mul2(w0, m);
              const float2 w = f2(a.x > 0.f ? w0.x : 0.f, a.y > 0.f ? w0.y : 0.f);
              g = add2(g, gU);
              gd[p] = mul2(w, om2);
              accGU[p] = add2(accGU[p], gU);
              accW[p] = add2(accW[p], w);
              pom2 = fma2(w, d[p], pom2);
              pphi2 = fma2(gd[p], d2[p], pphi2);
            }
          }
          gE[p] = g;
          if (GRAD) {
            pcf2 = add2(pcf2, g);
            pphi2 = fma2(g, d[p], pphi2);
            accNu[0][p] = add2(accNu[0][p], g);
          }
        }
        if (GRAD) {
#pragma unroll
          for (int k = 1; k < K; ++k) {
            const float2 z = tpair(k - 1);
#pragma unroll
            for (int p = 0; p < NP; ++p) accNu[k][p] = fma2(gE[p], z, accNu[k][p]);
            if (VELO) {
              const float2 z1 = tpair(P_Z1 + k - 1);
#pragma unroll
              for (int p = 0; p < NP; ++p) accNu[k][p] = fma2(gd[p], z1, accNu[k][p]);
            }
          }
          pcf = pcf2.x + pcf2.y;
          pphi = pphi2.x + pphi2.y;
          pom = pom2.x + pom2.y;
        }
      }
      if (rr == R / 2 - 1) poll_refill();
      if (GRAD) {
        part[rr * NQ + 0] = pcf;
        part[rr * NQ + 1] = pphi;
        if (VELO) part[rr * NQ + 2] = pom;
      }
    }
    if (GRAD) {
      // reduce-scatter over the stage's R cells: after log2(R) halving steps the lanes whose bits
      // [4 .. 5-log2 R] spell cell c hold that cell's NQ sums over their lane group; finish with an
      // all-reduce inside the group and let the group's first lane store them.
      int n = R * NQ;
      int off = 16;
#pragma unroll
      for (int step = 0; step < 5; ++step, off >>= 1) {
        if ((R >> step) > 1) {
          n >>= 1;
          const bool upper = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < n; ++i) {
            const float keep = upper ? part[i + n] : part[i];
            const float send = upper ? part[i] : part[i + n];
            part[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        } else {
#pragma unroll
          for (int i = 0; i < NQ; ++i) part[i] += __shfl_xor_sync(0xffffffffu, part[i], off);
        }
      }
      constexpr int LPC = 32 / R;  // lanes per cell group
      const int cell = lane / LPC;
      if ((lane % LPC) == 0 && cell < nv) {
        float* dst = cellpart_w + (c0 + (long long)st * R + cell) * NQ;
#pragma unroll
        for (int i = 0; i < NQ; ++i) dst[i] = part[i];
      }
    } else {
      __syncwarp();
    }
    if (lane == 0) mbar_arrive(&done[slot]);  // this warp no longer needs the slot
    poll_refill();
  }

  // ---- flush per-gene partial sums ----------------------------------------------------------------
  if (has_data) {
    if (GRAD && P.Nb > 0 && cur_b >= 0) {
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const long long g0 = gj + 2 * p;
        if (g0 < P.Ng) atomicAdd(&P.d_dnu[(long long)cur_b * P.Ng + g0], accNu[0][p].x);
        if (g0 + 1 < P.Ng) atomicAdd(&P.d_dnu[(long long)cur_b * P.Ng + g0 + 1], accNu[0][p].y);
      }
    }
    constexpr int ROWS = gene_rows(H);
    float* gp = P.genepart + ((long long)split * ROWS) * P.ld + gj;
    auto st4 = [&](int row, const float2* v) {
#pragma unroll
      for (int p = 0; p < NP; ++p) *reinterpret_cast<float2*>(gp + (long long)row * P.ld + 2 * p) = v[p];
    };
    st4(ROW_AS, accAS);
    st4(ROW_LS, accLS);
    if (VELO) {
      st4(ROW_AU, accAU);
      st4(ROW_LU, accLU);
      if (GRAD) {
        st4(ROW_GU, accGU);
        st4(ROW_W, accW);
      }
    }
    if (LGINLINE) st4(ROW_PSI, accPsi);
    if (GRAD) {
#pragma unroll
      for (int k = 0; k < K; ++k) st4(ROW_DNU + k, accNu[k]);
    }
  }
}

}  // namespace vcb
