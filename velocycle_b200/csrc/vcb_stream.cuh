// The streaming kernel: one pass over the [cells][genes] count tiles of S (and U) that produces the
// per-gene log-prob partial sums AND every gradient partial sum.
//
// Mapping (genes are the contiguous axis of the counts, preprocessing.py:193-194):
//   * a CTA owns a gene tile of 4*blockDim genes and a contiguous range of cells;
//   * a thread owns 4 adjacent genes: their Fourier coefficients, dispersion and kinetics stay in
//     registers for the whole kernel, as do the per-gene accumulators (no atomics on the hot path);
//   * cells are streamed through a shared-memory ring filled by the TMA engine with 1-D bulk copies
//     (cp.async.bulk + mbarrier complete_tx); a stage = kCellsPerStage count rows of S and U plus the
//     per-cell table rows (Fourier basis, its derivatives, omega, size factor, batch id) built by
//     vcb_cell_tables_kernel, so a consumer thread reads counts with one conflict-free LDS.128 per
//     matrix and the per-cell constants with broadcast LDS.128;
//   * per-cell sums (d/dphi, d/dcf, d/domega run over genes, i.e. across threads) go through a
//     double-buffered shared-memory transpose: each thread stores its 4-gene partial, and after the
//     one __syncthreads per stage the warps share out the row sums.
//
// Arithmetic per (cell, gene), SURVEY.md Appendix A, in base-2 logs so that ex2/lg2 are single MUFU ops:
//   y = etaS*log2(e); eS = 2^y; tS = r+eS; LS = lg2(tS); qS = 1/tS; gS = r (kS-eS) qS
//   a = d*omega+gamma; m = relu(a)+1e-5; mb = m/beta; eU = eS*mb; tU = r+eU; LU = lg2(tU)
//   w0 = r (kU-eU)/(tU*m); gU = w0*m; w = 1[a>0] w0
//   log-prob pieces: kS*(y-LS), LS, kU*(y+lg2(mb)-LU), LU   (times ln2, plus per-gene terms, in the epilogue)
#pragma once
#include "vcb_common.cuh"

namespace vcb {

constexpr int kCellsPerStage = 4;  // R
constexpr int kStages = 2;         // ring depth
constexpr int kGenesPerThread = 4;
constexpr int kMaxThreads = 512;

// per-gene partial rows written by the streaming kernel
enum GeneRow { ROW_AS = 0, ROW_LS = 1, ROW_AU = 2, ROW_LU = 3, ROW_GU = 4, ROW_W = 5, ROW_PSI = 6, ROW_DNU = 7 };

__host__ __device__ constexpr int table_width(int H) { return ((6 * H + 3) + 3) / 4 * 4; }
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
  float* cellpart;  // [n_tiles][NQ][Nc]
  float* d_dnu;     // [Nb][Ng], zeroed, atomically accumulated at batch boundaries
  long long Nc, Ng, ld;
  int n_split;
  int Nb;
};

struct StreamSmem {
  // byte offsets inside dynamic shared memory
  int tab_off, cnt_off, red_off, total;
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
  L.red_off = off;
  if (grad) off += 2 * kCellsPerStage * (velo ? 3 : 2) * nthr * 4;
  L.total = off;
  return L;
}

template <int H, bool VELO, bool GRAD, bool LGINLINE>
__global__ void __launch_bounds__(kMaxThreads, 1) vcb_stream_kernel(const StreamParams P) {
  constexpr int K = 2 * H + 1;
  constexpr int TABW = table_width(H);
  constexpr int NMAT = VELO ? 2 : 1;
  constexpr int NQ = VELO ? 3 : 2;
  constexpr int R = kCellsPerStage;
  constexpr int NS = kStages;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  const int tile = blockIdx.x;
  const int split = blockIdx.y;
  const long long g_base = (long long)tile * kGenesPerThread * nthr;
  const long long rem = P.ld - g_base;
  const int W = (int)(rem < (long long)kGenesPerThread * nthr ? rem : (long long)kGenesPerThread * nthr);
  const StreamSmem L = stream_smem_layout(H, VELO, GRAD, nthr, W);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);
  float* s_tab = reinterpret_cast<float*>(smem_raw + L.tab_off);
  float* s_cnt = reinterpret_cast<float*>(smem_raw + L.cnt_off);
  float* s_red = reinterpret_cast<float*>(smem_raw + L.red_off);

  // this CTA's cells
  const long long c0 = (P.Nc * split) / P.n_split;
  const long long c1 = (P.Nc * (split + 1)) / P.n_split;
  const int n_cells = (int)(c1 - c0);
  const int n_stages = (n_cells + R - 1) / R;

  // ---- per-gene state in registers ---------------------------------------------------------------
  const long long gj = g_base + (long long)kGenesPerThread * tid;
  const bool has_data = kGenesPerThread * tid < W;  // this thread's 16 bytes exist in the smem rows
  float nu[K][4], r[4], gam[4], invb[4], nu0eff[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const long long g = gj + j;
    if (has_data && g < P.Ng) {
#pragma unroll
      for (int k = 0; k < K; ++k) nu[k][j] = P.nu[g * K + k];
      r[j] = 1.0f / P.shape_inv[g];
      if (VELO) {
        gam[j] = P.gamma[g];
        invb[j] = expf(-P.logbeta[g]);
      } else {
        gam[j] = 1.f;
        invb[j] = 1.f;
      }
    } else {  // padding column: eta = -inf makes every contribution exactly zero
#pragma unroll
      for (int k = 0; k < K; ++k) nu[k][j] = 0.f;
      nu[0][j] = -1e30f;
      r[j] = 1.f;
      gam[j] = 1.f;
      invb[j] = 1.f;
    }
    nu0eff[j] = nu[0][j];
  }
  int cur_b = -1;

  float accAS[4] = {0, 0, 0, 0}, accLS[4] = {0, 0, 0, 0};
  float accAU[4] = {0, 0, 0, 0}, accLU[4] = {0, 0, 0, 0};
  float accGU[4] = {0, 0, 0, 0}, accW[4] = {0, 0, 0, 0}, accPsi[4] = {0, 0, 0, 0};
  float accNu[K][4];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) accNu[k][j] = 0.f;

  // ---- producer helpers (thread 0) ---------------------------------------------------------------
  auto issue_stage = [&](int st) {
    const int slot = st % NS;
    const long long cs = c0 + (long long)st * R;
    const int nv = (n_cells - st * R) < R ? (n_cells - st * R) : R;
    const uint32_t row_bytes = (uint32_t)W * 4u;
    const uint32_t bytes = (uint32_t)nv * (TABW * 4u + NMAT * row_bytes);
    mbar_expect_tx(&mbar[slot], bytes);
    bulk_g2s(s_tab + (size_t)slot * R * TABW, P.tab + cs * TABW, (uint32_t)nv * TABW * 4u, &mbar[slot]);
    float* dstS = s_cnt + (size_t)slot * NMAT * R * W;
    if ((long long)W == P.ld) {  // the tile spans whole rows: the nv rows are one contiguous block
      bulk_g2s(dstS, P.S + cs * P.ld, (uint32_t)nv * row_bytes, &mbar[slot]);
      if (VELO) bulk_g2s(dstS + (size_t)R * W, P.U + cs * P.ld, (uint32_t)nv * row_bytes, &mbar[slot]);
    } else {
      for (int rr = 0; rr < nv; ++rr) {
        bulk_g2s(dstS + (size_t)rr * W, P.S + (cs + rr) * P.ld + g_base, row_bytes, &mbar[slot]);
        if (VELO)
          bulk_g2s(dstS + (size_t)(R + rr) * W, P.U + (cs + rr) * P.ld + g_base, row_bytes, &mbar[slot]);
      }
    }
  };

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(&mbar[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (int st = 0; st < NS && st < n_stages; ++st) issue_stage(st);
  }

  // ---- per-cell row sums across the CTA (after the stage barrier) ---------------------------------
  auto reduce_rows = [&](int st) {
    if (!GRAD) return;
    const int nv = (n_cells - st * R) < R ? (n_cells - st * R) : R;
    const float* red = s_red + (size_t)(st & 1) * R * NQ * nthr;
    const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;
    for (int row = warp; row < nv * NQ; row += nwarps) {
      const float* src = red + (size_t)row * nthr;
      float s = 0.f;
      for (int i = lane; i < nthr; i += 32) s += src[i];
      s = warp_sum(s);
      if (lane == 0) {
        const int rr = row / NQ, q = row - rr * NQ;
        P.cellpart[((long long)tile * NQ + q) * P.Nc + c0 + (long long)st * R + rr] = s;
      }
    }
  };

  // ---- main loop ----------------------------------------------------------------------------------
  for (int st = 0; st < n_stages; ++st) {
    const int slot = st % NS;
    mbar_wait(&mbar[slot], (uint32_t)((st / NS) & 1));
    const int nv = (n_cells - st * R) < R ? (n_cells - st * R) : R;
    const float* tabs = s_tab + (size_t)slot * R * TABW;
    const float* cntS = s_cnt + (size_t)slot * NMAT * R * W;
    float* red = s_red + (size_t)(st & 1) * R * NQ * nthr;

#pragma unroll 1
    for (int rr = 0; rr < nv; ++rr) {
      const float* t = tabs + rr * TABW;
      float pcf = 0.f, pphi = 0.f, pom = 0.f;
      if (has_data) {
        const float omega = t[6 * H];
        const float cf = t[6 * H + 1];
        const int b = __float_as_int(t[6 * H + 2]);
        if (P.Nb > 0 && b != cur_b) {  // CTA-uniform: batch boundary (rare when samples are concatenated)
          if (GRAD && cur_b >= 0 && P.d_dnu != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (gj + j < P.Ng) atomicAdd(&P.d_dnu[(long long)cur_b * P.Ng + gj + j], accNu[0][j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) accNu[0][j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            nu0eff[j] = (gj + j < P.Ng) ? nu[0][j] + P.dnu[(long long)b * P.Ng + gj + j] : nu[0][j];
          cur_b = b;
        }

        // forward contraction: etaS, d = nu.zeta', d2 = nu.zeta''
        float eta[4], d[4], d2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          eta[j] = nu0eff[j] + cf;
          d[j] = 0.f;
          d2[j] = 0.f;
        }
#pragma unroll
        for (int k = 1; k < K; ++k) {
          const float z = t[k - 1];
          const float z1 = t[2 * H + k - 1];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            eta[j] = fmaf(nu[k][j], z, eta[j]);
            d[j] = fmaf(nu[k][j], z1, d[j]);
          }
          if (VELO && GRAD) {
            const float z2 = t[4 * H + k - 1];
#pragma unroll
            for (int j = 0; j < 4; ++j) d2[j] = fmaf(nu[k][j], z2, d2[j]);
          }
        }

        const float4 kS4 = *reinterpret_cast<const float4*>(cntS + (size_t)rr * W + 4 * tid);
        const float kS[4] = {kS4.x, kS4.y, kS4.z, kS4.w};
        float kU[4] = {0.f, 0.f, 0.f, 0.f};
        if (VELO) {
          const float4 kU4 = *reinterpret_cast<const float4*>(cntS + (size_t)(R + rr) * W + 4 * tid);
          kU[0] = kU4.x, kU[1] = kU4.y, kU[2] = kU4.z, kU[3] = kU4.w;
        }

        float gE[4], gd[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float y = eta[j] * kLog2e;
          const float eS = ex2_approx(y);
          const float tS = r[j] + eS;
          const float LS = lg2_approx(tS);
          accAS[j] = fmaf(kS[j], y - LS, accAS[j]);
          accLS[j] += LS;
          float g = 0.f;
          if (GRAD) g = r[j] * (kS[j] - eS) * rcp_approx(tS);
          if (LGINLINE) {
            float psi;
            const float lg = lgamma_terms_inline(r[j], kS[j], psi);
            accAS[j] = fmaf(lg, kLog2e, accAS[j]);
            accPsi[j] += psi;
          }
          gd[j] = 0.f;
          if (VELO) {
            const float a = fmaf(d[j], omega, gam[j]);
            const float m = fmaxf(a, 0.f) + 1e-5f;
            const float mb = m * invb[j];
            const float eU = eS * mb;
            const float tU = r[j] + eU;
            const float LU = lg2_approx(tU);
            const float lmb = lg2_approx(mb);
            accAU[j] = fmaf(kU[j], (y + lmb) - LU, accAU[j]);
            accLU[j] += LU;
            if (LGINLINE) {
              float psi;
              const float lg = lgamma_terms_inline(r[j], kU[j], psi);
              accAU[j] = fmaf(lg, kLog2e, accAU[j]);
              accPsi[j] += psi;
            }
            if (GRAD) {
              const float w0 = r[j] * (kU[j] - eU) * rcp_approx(tU * m);
              const float gU = w0 * m;
              const float w = a > 0.f ? w0 : 0.f;
              g += gU;
              gd[j] = w * omega;
              accGU[j] += gU;
              accW[j] += w;
              pom = fmaf(w, d[j], pom);
              pphi = fmaf(gd[j], d2[j], pphi);
            }
          }
          gE[j] = g;
          if (GRAD) {
            pcf += g;
            pphi = fmaf(g, d[j], pphi);
            accNu[0][j] += g;
          }
        }
        if (GRAD) {
#pragma unroll
          for (int k = 1; k < K; ++k) {
            const float z = t[k - 1];
#pragma unroll
            for (int j = 0; j < 4; ++j) accNu[k][j] = fmaf(gE[j], z, accNu[k][j]);
            if (VELO) {
              const float z1 = t[2 * H + k - 1];
#pragma unroll
              for (int j = 0; j < 4; ++j) accNu[k][j] = fmaf(gd[j], z1, accNu[k][j]);
            }
          }
        }
      }
      if (GRAD) {
        red[(rr * NQ + 0) * nthr + tid] = pcf;
        red[(rr * NQ + 1) * nthr + tid] = pphi;
        if (VELO) red[(rr * NQ + 2) * nthr + tid] = pom;
      }
    }
    __syncthreads();  // everyone is done with this slot and has published its per-cell partials
    if (tid == 0 && st + NS < n_stages) issue_stage(st + NS);
    reduce_rows(st);
  }

  // ---- flush per-gene partial sums ----------------------------------------------------------------
  if (has_data) {
    if (GRAD && P.Nb > 0 && cur_b >= 0 && P.d_dnu != nullptr) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (gj + j < P.Ng) atomicAdd(&P.d_dnu[(long long)cur_b * P.Ng + gj + j], accNu[0][j]);
    }
    constexpr int ROWS = gene_rows(H);
    float* gp = P.genepart + ((long long)split * ROWS) * P.ld + gj;
    auto st4 = [&](int row, const float* v) {
      *reinterpret_cast<float4*>(gp + (long long)row * P.ld) = make_float4(v[0], v[1], v[2], v[3]);
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
