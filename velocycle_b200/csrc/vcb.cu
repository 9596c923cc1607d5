// libvcb: C-ABI entry points (include/vcb.h), the small per-cell / per-gene kernels around the
// streaming kernel (vcb_stream.cuh), the count histogram and the fused ClippedAdam.
#include "vcb.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "vcb_common.cuh"
#include "vcb_stream.cuh"
#include "vcb_stream2.cuh"
#include "vcb_umma.cuh"

#ifndef VCB_DEFAULT_NP
#define VCB_DEFAULT_NP 1
#endif

namespace vcb {

// ======================================================================================================
// Per-cell prologue: the operand table of every 8-cell group (layout: vcb_stream.cuh, TabSection).
// One warp per group: lanes 0..7 evaluate the Fourier basis of their cell (column order [sin, cos] per harmonic
// and sin(fl(n*phi)) follow utils.py:420-435), then every lane gathers the two operand values each section
// expects in its MMA B fragment, splits them into TF32 hi/lo and stores one float4 (coalesced 512 B rows).
// ======================================================================================================
struct CellParams {
  const float* phi;
  const float* cf;
  const int32_t* batch_id;
  const int32_t* cond_id;
  const float* nu_omega;  // [Nx][Kw] or null
  float* tab;
  long long Nc, n_groups;
  int H, Hw, Nx, velo, tabg;
  int v2;  // tables for vcb_stream2.cuh: omega * ln2 in the tail, backward sections as {Zhi, Zhi, Zlo, Zlo} TF32 pairs,
           // slot 16 of the first group of every 16-cell stage = the stage's batch (or -1)
  // The first n_spec_blocks blocks evaluate the parameter-only lgamma / digamma sums over the count spectra (fp64, the slow
  // part of the per-gene epilogue).  They depend on shape_inv alone, so they run here, beside the table blocks and before
  // the streaming kernel, instead of on the serial tail of the step; spec_out = [Ng][4] doubles {lgS, psS, lgU, psU}.
  unsigned n_spec_blocks;
  vcb_spectrum_t spec_S, spec_U;
  const float* shape_inv;
  double* spec_out;
  long long Ng;
};

__device__ void spectrum_block(const CellParams& P, unsigned block);

constexpr int kCellEpiThreads = 256;
constexpr int kTabWarps = 8;
constexpr int kTabSlots = 16;  // 8 * ksteps(VCB_MAX_HARMONICS)

__global__ void __launch_bounds__(kTabWarps * 32) vcb_cell_tables_kernel(const CellParams P) {
  // per warp and cell: [0] forward eta operand, [1] zeta', [2] omega*zeta'', [3] backward zeta, [4] omega*zeta'
  __shared__ float sv[kTabWarps][5][kGroupCells][kTabSlots];
  __shared__ float s_tail[kTabWarps][20];
  if (blockIdx.x < P.n_spec_blocks) {  // the first blocks: they are the long ones (fp64 lgamma / digamma), start them first
    spectrum_block(P, blockIdx.x);
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long group = (long long)(blockIdx.x - P.n_spec_blocks) * kTabWarps + warp;
  if (group >= P.n_groups) return;
  const int H = P.H, K = 2 * H + 1, KS = ksteps(H);
  {  // clear the warp's staging block with all 32 lanes (160 float4)
    float4* z = reinterpret_cast<float4*>(&sv[warp][0][0][0]);
#pragma unroll
    for (int i = 0; i < 5 * kGroupCells * kTabSlots / 4 / 32; ++i) z[i * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncwarp();
  if (lane < kGroupCells) {
    const long long c = group * kGroupCells + lane;
    const bool valid = c < P.Nc;
    const long long cl = valid ? c : P.Nc - 1;  // padding cells copy the batch of the last cell
    const float phi = valid ? P.phi[c] : 0.f;
    // sin / cos of the harmonics: one accurate sincosf, the multiples by the angle-addition recurrence (a few ulp at n = 5)
    float sn[VCB_MAX_HARMONICS + 1], cn[VCB_MAX_HARMONICS + 1];
    sn[0] = 0.f;
    cn[0] = 1.f;
    sincosf(phi, &sn[1], &cn[1]);
#pragma unroll
    for (int n = 2; n <= VCB_MAX_HARMONICS; ++n) {
      sn[n] = fmaf(sn[n - 1], cn[1], cn[n - 1] * sn[1]);
      cn[n] = fmaf(cn[n - 1], cn[1], -sn[n - 1] * sn[1]);
    }
    float omega = 0.f;
    if (P.velo && P.nu_omega != nullptr) {
      const int x = P.cond_id ? P.cond_id[cl] : 0;
      const int Kw = 2 * P.Hw + 1;
      const float* nw = P.nu_omega + (long long)x * Kw;
      omega = nw[0];
#pragma unroll
      for (int n = 1; n <= VCB_MAX_HARMONICS; ++n)
        if (n <= P.Hw) {
          omega = fmaf(nw[2 * n - 1], sn[n], omega);
          omega = fmaf(nw[2 * n], cn[n], omega);
        }
    }
    float(*v)[kGroupCells][kTabSlots] = sv[warp];
    v[0][lane][0] = 1.f;
    v[3][lane][0] = 1.f;
#pragma unroll
    for (int n = 1; n <= VCB_MAX_HARMONICS; ++n) {
      if (n > H) break;
      const float fn = (float)n, s = sn[n], co = cn[n];
      const int ks = 2 * n - 1, kc = 2 * n;
      v[0][lane][ks] = s;
      v[0][lane][kc] = co;
      v[3][lane][ks] = s;
      v[3][lane][kc] = co;
      v[1][lane][ks] = fn * co;
      v[1][lane][kc] = -fn * s;
      v[4][lane][ks] = omega * (fn * co);
      v[4][lane][kc] = omega * (-fn * s);
      v[2][lane][ks] = omega * (-fn * fn * s);
      v[2][lane][kc] = omega * (-fn * fn * co);
    }
    // spare slot K: the size factor rides the forward contraction (A holds 1 there); a padding cell gets
    // eta = -3e4 so that all its terms vanish; backward it collects sum_c w = d/dgamma
    v[0][lane][K] = valid ? (P.cf ? P.cf[c] : 0.f) : -30000.f;  // (inside the FP16 range of the cross-term operand)
    v[4][lane][K] = 1.f;
    s_tail[warp][lane] = P.v2 ? omega * kLn2 : omega;
    s_tail[warp][8 + lane] = __int_as_float(P.batch_id ? P.batch_id[cl] : 0);
  }
  __syncwarp();
  const int grp = lane >> 2, q = lane & 3;
  float4* out = reinterpret_cast<float4*>(P.tab + group * P.tabg);
  // one table entry: {TF32 hi of the main MMA's b0, b1; 16-bit pairs {y0, y1} and {lo y0, lo y1} of the cross-term
  // MMA: FP16 for the forward sections, BF16 for the backward ones (see mma_split_fwd / mma_split_bwd)}
  auto emit = [&](int sec_out, int ks, bool f16, float x0, float x1, float y0, float y1) {
    if (P.v2 && !f16) {  // 3xTF32 backward operand: hi and lo parts of the same two values
      out[(sec_out * KS + ks) * 32 + lane] = make_float4(__uint_as_float(tf32_rna(x0)), __uint_as_float(tf32_rna(x1)),
                                                         tf32_lo(x0), tf32_lo(x1));
      return;
    }
    const uint32_t c0 = f16 ? pack_f16(y0, y1) : pack_bf16(y0, y1);
    const uint32_t c1 = f16 ? pack_f16(tf32_lo(y0), tf32_lo(y1)) : pack_bf16(tf32_lo(y0), tf32_lo(y1));
    out[(sec_out * KS + ks) * 32 + lane] = make_float4(__uint_as_float(tf32_rna(x0)), __uint_as_float(tf32_rna(x1)),
                                                       __uint_as_float(c0), __uint_as_float(c1));
  };
  const int fc = fwd_cell(grp);
  for (int ks = 0; ks < KS; ++ks) {
    // forward: main b0 (k = q, n = grp), b1 (k = q+4); cross k = 2q, 2q+1 (and their lo parts at 2q+8, 2q+9);
    // column n is cell fwd_cell(n)
    auto fwd = [&](int sec_out, int sec) {
      const float* z = sv[warp][sec][fc] + 8 * ks;
      emit(sec_out, ks, true, z[q], z[q + 4], z[2 * q], z[2 * q + 1]);
    };
    // backward: k runs over cells: main b0 (cell q, n = slot grp), b1 (cell q+4); cross pairs the same two cells
    auto bwd = [&](int sec_out, int sec) {
      const float z0 = sv[warp][sec][q][8 * ks + grp], z1 = sv[warp][sec][q + 4][8 * ks + grp];
      emit(sec_out, ks, false, z0, z1, z0, z1);
    };
    fwd(SEC_F0, 0);
    fwd(SEC_F1, 1);
    bwd(SEC_B0, 3);
    if (P.velo) {
      fwd(SEC_F2, 2);
      bwd(SEC_B1, 4);
    }
  }
  if (lane == 0) {  // the group's batch id if its 8 cells agree, else -1 (the streaming kernel's fast-path test)
    int b = __float_as_int(s_tail[warp][8]);
    for (int i = 1; i < kGroupCells; ++i)
      if (__float_as_int(s_tail[warp][8 + i]) != b) b = -1;
    if (P.v2 && P.batch_id != nullptr) {  // the whole 16-cell stage must agree (padding cells copy the last cell)
      const long long c0 = (group / s2::kGPS) * s2::kStageCells;
      for (int i = 0; i < s2::kStageCells; ++i) {
        const long long c = c0 + i < P.Nc ? c0 + i : P.Nc - 1;
        if (P.batch_id[c] != b) b = -1;
      }
    }
    s_tail[warp][16] = __int_as_float(b);
    s_tail[warp][17] = s_tail[warp][18] = s_tail[warp][19] = 0.f;
  }
  __syncwarp();
  if (lane < 20) P.tab[group * P.tabg + table_tail(H, P.velo != 0) + lane] = s_tail[warp][lane];
}

// ------------------------------------------------------------------------------------------------------
// The same tables for vcb_stream2.cuh (one k-step, H <= 3), four groups per warp: every lane owns one of the warp's 32 cells
// for the per-cell part (one sincosf, the higher harmonics by recurrence, omega), then the warp emits the four groups'
// fragments.  The one-group kernel above ran its per-cell part on 8 of 32 lanes; this one needs less than half the
// instructions per group.
// ------------------------------------------------------------------------------------------------------
constexpr int kTab4Groups = 4;
constexpr int kTab4Pitch = 9;  // 8 slots + 1: the per-cell writes of a warp (stride = one cell) hit 32 different banks

__global__ void __launch_bounds__(kTabWarps * 32) vcb_cell_tables4_kernel(const CellParams P) {
  // sections [0] zeta (forward eta operand; the backward zeta is the same with the spare slot cleared), [1] zeta',
  // [2] omega*zeta'', [3] omega*zeta'
  __shared__ float sv[kTabWarps][kTab4Groups][4][kGroupCells][kTab4Pitch];
  __shared__ float s_tail[kTabWarps][kTab4Groups][20];
  if (blockIdx.x < P.n_spec_blocks) {
    spectrum_block(P, blockIdx.x);
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long group0 = ((long long)(blockIdx.x - P.n_spec_blocks) * kTabWarps + warp) * kTab4Groups;
  if (group0 >= P.n_groups) return;
  const int H = P.H, K = 2 * H + 1;
  {  // ---- per cell: lane = (group gi, cell ci) -------------------------------------------------------------------------
    const int gi = lane >> 3, ci = lane & 7;
    const long long c = group0 * kGroupCells + lane;
    const bool valid = c < P.Nc;
    const long long cl = valid ? c : P.Nc - 1;  // padding cells copy the batch of the last cell
    const float phi = valid ? P.phi[c] : 0.f;
    float sn[4], cn[4];  // harmonics 1..3 (slot 0 unused)
    sincosf(phi, &sn[1], &cn[1]);
#pragma unroll
    for (int n = 2; n <= 3; ++n) {
      sn[n] = fmaf(sn[n - 1], cn[1], cn[n - 1] * sn[1]);
      cn[n] = fmaf(cn[n - 1], cn[1], -sn[n - 1] * sn[1]);
    }
    float omega = 0.f;
    if (P.velo && P.nu_omega != nullptr) {
      const int x = P.cond_id ? P.cond_id[cl] : 0;
      const float* nw = P.nu_omega + (long long)x * (2 * P.Hw + 1);
      omega = nw[0];
#pragma unroll
      for (int n = 1; n <= 3; ++n)
        if (n <= P.Hw) {
          omega = fmaf(nw[2 * n - 1], sn[n], omega);
          omega = fmaf(nw[2 * n], cn[n], omega);
        }
      for (int n = 4; n <= P.Hw; ++n) {
        float s, co;
        sincosf((float)n * phi, &s, &co);
        omega = fmaf(nw[2 * n - 1], s, omega);
        omega = fmaf(nw[2 * n], co, omega);
      }
    }
    float v[4][8];  // all 8 slots of every section are written
#pragma unroll
    for (int sec = 0; sec < 4; ++sec)
#pragma unroll
      for (int k = 0; k < 8; ++k) v[sec][k] = 0.f;
    v[0][0] = 1.f;
#pragma unroll
    for (int n = 1; n <= 3; ++n)
      if (n <= H) {
        const float fn = (float)n, s = sn[n], co = cn[n];
        const int ks = 2 * n - 1, kc = 2 * n;
        v[0][ks] = s;
        v[0][kc] = co;
        v[1][ks] = fn * co;
        v[1][kc] = -fn * s;
        v[3][ks] = omega * (fn * co);
        v[3][kc] = omega * (-fn * s);
        v[2][ks] = omega * (-fn * fn * s);
        v[2][kc] = omega * (-fn * fn * co);
      }
    // spare slot K: the size factor rides the forward contraction; a padding cell gets eta = -3e4; backward: sum_c w = d/dgamma
    const float spare0 = valid ? (P.cf ? P.cf[c] : 0.f) : -30000.f;
#pragma unroll
    for (int k = 1; k < 8; k += 2)  // K is odd
      if (k == K) {
        v[0][k] = spare0;
        v[3][k] = 1.f;
      }
#pragma unroll
    for (int sec = 0; sec < 4; ++sec)
#pragma unroll
      for (int k = 0; k < 8; ++k) sv[warp][gi][sec][ci][k] = v[sec][k];
    const int bid = P.batch_id ? P.batch_id[cl] : 0;
    s_tail[warp][gi][ci] = omega * kLn2;
    s_tail[warp][gi][8 + ci] = __int_as_float(bid);
    // the stage's batch if all of its cells agree, else -1 (slot 16 of every group of the stage)
    constexpr int SC = s2::kStageCells;
    const int first = __shfl_sync(0xffffffffu, bid, lane & ~(SC - 1));
    const unsigned stage_mask = SC >= 32 ? 0xffffffffu : (((1u << (SC & 31)) - 1u) << (lane & ~(SC - 1)));
    const unsigned agree = __ballot_sync(0xffffffffu, bid == first);
    if (ci == 0) {
      s_tail[warp][gi][16] = __int_as_float((agree & stage_mask) == stage_mask ? first : -1);
      s_tail[warp][gi][17] = s_tail[warp][gi][18] = s_tail[warp][gi][19] = 0.f;
    }
  }
  __syncwarp();
  // ---- fragments: lane = (grp, q) of the MMA layouts ----------------------------------------------------------------------
  const int grp = lane >> 2, q = lane & 3;
  const int fc = fwd_cell(grp);
  const int tail = table_tail(H, P.velo != 0);
  for (int g = 0; g < kTab4Groups; ++g) {
    const long long group = group0 + g;
    if (group >= P.n_groups) break;
    float4* out = reinterpret_cast<float4*>(P.tab + group * P.tabg);
    const float(*z)[kGroupCells][kTab4Pitch] = sv[warp][g];
    // forward sections: {TF32 hi of b0 (k = q, n = grp), b1 (k = q + 4); FP16 pairs {y(2q), y(2q+1)} and their lo parts}; column n
    // is cell fwd_cell(n)
    auto fwd = [&](int sec_out, int sec) {
      const float* zc = z[sec][fc];
      const float x0 = zc[q], x1 = zc[q + 4], y0 = zc[2 * q], y1 = zc[2 * q + 1];
      out[sec_out * 32 + lane] = make_float4(__uint_as_float(tf32_rna(x0)), __uint_as_float(tf32_rna(x1)),
                                             __uint_as_float(pack_f16(y0, y1)),
                                             __uint_as_float(pack_f16(tf32_lo(y0), tf32_lo(y1))));
    };
    // backward sections (3xTF32): k runs over cells: b0 (cell q, n = slot grp), b1 (cell q + 4): {hi, hi, lo, lo}
    auto bwd = [&](int sec_out, int sec, bool clear_spare) {
      float z0 = z[sec][q][grp], z1 = z[sec][q + 4][grp];
      if (clear_spare && grp == K) z0 = z1 = 0.f;
      out[sec_out * 32 + lane] =
          make_float4(__uint_as_float(tf32_rna(z0)), __uint_as_float(tf32_rna(z1)), tf32_lo(z0), tf32_lo(z1));
    };
    fwd(SEC_F0, 0);
    fwd(SEC_F1, 1);
    bwd(SEC_B0, 0, true);
    if (P.velo) {
      fwd(SEC_F2, 2);
      bwd(SEC_B1, 3, false);
    }
    if (lane < 20) P.tab[group * P.tabg + tail + lane] = s_tail[warp][g][lane];
  }
}

// ======================================================================================================
// Per-cell epilogue: sum the gene-tile partials, add the omega(phi) path to d/dphi, reduce d/dnu_omega.
// ======================================================================================================
struct CellEpiParams {
  const float* cellpart;  // [n_part][Ncp][NQ], n_part = gene tiles, Ncp = cells padded to whole groups
  const float* phi;
  const int32_t* cond_id;
  const float* nu_omega;
  float* d_phi;
  float* d_cf;
  float* d_omega;
  double* dnw_part;  // [gridDim.x][Nx*Kw] per-block partial sums of d/dnu_omega (summed in order by the gene epilogue)
  long long Nc, Ncp;
  int n_part, NQ, Hw, Nx;
  float scale;  // vcb_stream2 accumulates d/dphi and d/domega against log2(e)-scaled contractions: ln 2 here, else 1
};

__global__ void vcb_cell_epilogue_kernel(const CellEpiParams P) {
  extern __shared__ double s_acc[];  // [warps][Nx*Kw]
  const bool velo = P.NQ == 3;
  const int Kw = 2 * P.Hw + 1;
  const int nacc = velo ? P.Nx * Kw : 0;
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  float pom = 0.f, phi = 0.f;
  int x = -1;  // condition of this thread's cell; -1 = no cell
  if (c < P.Nc) {
    float q[3] = {0.f, 0.f, 0.f};
    for (int t = 0; t < P.n_part; ++t)
      for (int i = 0; i < P.NQ; ++i) q[i] += P.cellpart[((long long)t * P.Ncp + c) * P.NQ + i];
    q[1] *= P.scale;
    q[2] *= P.scale;
    float dphi = q[1];
    if (velo) {
      phi = P.phi[c];
      x = P.cond_id ? P.cond_id[c] : 0;
      const float* nw = P.nu_omega + (long long)x * Kw;
      pom = q[2];
      float domega_dphi = 0.f;
      for (int n = 1; n <= P.Hw; ++n) {
        float s, co;
        const float fn = (float)n;
        sincosf(fn * phi, &s, &co);
        domega_dphi = fmaf(nw[2 * n - 1], fn * co, domega_dphi);
        domega_dphi = fmaf(nw[2 * n], -fn * s, domega_dphi);
      }
      dphi = fmaf(pom, domega_dphi, dphi);
      if (P.d_omega) P.d_omega[c] = pom;
    }
    if (P.d_cf) P.d_cf[c] = q[0];
    if (P.d_phi) P.d_phi[c] = dphi;
  }
  // d/dnu_omega[x][h] = sum over the cells of condition x of d/domega * zeta_omega[h]: fp64 warp sums per condition
  // (shuffles), parked per warp and added in warp order -- no atomics, so the result is reproducible bit for bit
  if (velo) {
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int xx = 0; xx < P.Nx; ++xx) {
      const bool any = __any_sync(full, x == xx);
      const double w = (x == xx) ? (double)pom : 0.0;
      for (int h = 0; h < Kw; ++h) {
        double v = 0.0;
        if (any) {
          v = w;
          if (h > 0) {
            float s, co;
            sincosf((float)((h + 1) / 2) * phi, &s, &co);
            v *= (double)((h & 1) ? s : co);
          }
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(full, v, o);
        }
        if (lane == 0) s_acc[(size_t)warp * nacc + xx * Kw + h] = v;
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nacc; i += blockDim.x) {
      double t = 0.0;
      for (int w = 0; w < nwarps; ++w) t += s_acc[(size_t)w * nacc + i];
      P.dnw_part[(size_t)blockIdx.x * nacc + i] = t;
    }
  }
}

// ======================================================================================================
// Per-gene epilogue: sum the cell-split partials in fp64, add the parameter-only terms
// (n r log r, the lgamma / digamma sums over the count spectrum), apply the chain-rule factors.
// Block = 8 genes x 32 lanes (the fp64 lgamma/digamma sums over the spectrum want many lanes per gene).
// ======================================================================================================
struct GeneEpiParams {
  const float* genepart;  // [n_split][ROWS][ld]
  const float* shape_inv;
  const float* dnu_acc;  // [Nb][Ng] or null
  const double* dnw_part;  // [n_cell_blocks][Nx*Kw]
  int n_cell_blocks;
  vcb_spectrum_t spec_S, spec_U;
  float *lp_S, *lp_U, *d_nu, *d_dnu, *d_shape_inv, *d_logbeta, *d_gamma, *d_nu_omega;
  long long Nc, Ng, ld;
  int n_split, H, Nb, Nx, Hw;
  int velo, grad, lginline;
  int dnu_rows;  // single batch, d/ddnu = the constant column of d/dnu (tcgen05 path): no atomics buffer
  int v2;        // partial rows written by vcb_stream2: ROW_AS / ROW_AU already hold the -r L terms, ROW_LS = sum(LS + LU);
                 // dnu_acc = per-split batch sums [n_split][Nb][ld] (the constant column of ROW_DNU is then zero)
  const double* spec_pre;  // [Ng][4] spectrum sums evaluated beside the table kernel (see CellParams), or null: evaluate here
};

constexpr int kEpiGenes = 8;
constexpr int kEpiLanes = 32;
constexpr int kEpiMaxRows = ROW_DNU + 2 * VCB_MAX_HARMONICS + 1;

__device__ __forceinline__ void spectrum_sums(const vcb_spectrum_t& sp, long long g, double r, int lane, int nlanes,
                                              double& lg_sum, double& psi_sum) {
  lg_sum = 0.0;
  psi_sum = 0.0;
  if (sp.off == nullptr) return;
  const int e0 = sp.off[g], e1 = sp.off[g + 1];
  if (e1 <= e0) return;
  double lgr, psr;
  lgamma_digamma_d(r, lgr, psr);
  for (int e = e0 + lane; e < e1; e += nlanes) {
    const double k = (double)sp.val[e], m = (double)sp.mult[e];
    double lgk, psk;
    lgamma_digamma_d(r + k, lgk, psk);
    lg_sum += m * (lgk - lgr);
    psi_sum += m * (psk - psr);
  }
}

// the spectrum role of the table kernel's extra blocks: block b serves genes [8 b, 8 b + 8), 32 lanes per gene
__device__ void spectrum_block(const CellParams& P, unsigned block) {
  __shared__ double s_sp[4][kEpiLanes][kEpiGenes];
  const int gx = threadIdx.x % kEpiGenes, ly = threadIdx.x / kEpiGenes;
  const long long g = (long long)block * kEpiGenes + gx;
  const bool valid = g < P.Ng;
  double a = 0, b = 0, c = 0, d = 0;
  if (valid) {
    const double r = 1.0 / (double)P.shape_inv[g];
    spectrum_sums(P.spec_S, g, r, ly, kEpiLanes, a, b);
    if (P.spec_U.off != nullptr) spectrum_sums(P.spec_U, g, r, ly, kEpiLanes, c, d);
  }
  s_sp[0][ly][gx] = a;
  s_sp[1][ly][gx] = b;
  s_sp[2][ly][gx] = c;
  s_sp[3][ly][gx] = d;
  __syncthreads();
  if (ly < 4 && valid) {  // fixed summation order: deterministic
    double t = 0.0;
    for (int l = 0; l < kEpiLanes; ++l) t += s_sp[ly][l][gx];
    P.spec_out[g * 4 + ly] = t;
  }
}

__global__ void __launch_bounds__(kEpiGenes* kEpiLanes) vcb_gene_epilogue_kernel(const GeneEpiParams P) {
  __shared__ double s_rows[kEpiMaxRows][kEpiGenes];
  const int gx = threadIdx.x % kEpiGenes, ly = threadIdx.x / kEpiGenes;
  const long long g = (long long)blockIdx.x * kEpiGenes + gx;
  const int K = 2 * P.H + 1;
  const int ROWS = ROW_DNU + K;
  const bool valid = g < P.Ng;

  // per-split partial rows -> per-gene sums: every thread first sums its share of the splits for ALL rows (independent loads, no
  // barrier in between: the row-by-row version with two barriers per row was a 14-deep latency chain), then thread (row, gene)
  // adds the kEpiLanes partials of its row in a fixed order (deterministic)
  __shared__ double s_rowpart[kEpiMaxRows][kEpiLanes][kEpiGenes];
  double(*s_part)[kEpiGenes] = s_rowpart[0];               // reused by the batch loop below
  double(*s_spec)[kEpiLanes][kEpiGenes] = &s_rowpart[1];   // and rows 1..4 by the spectrum sums (after the barrier)
  static_assert(kEpiMaxRows >= 5, "s_spec aliases rows 1..4");
  {
    unsigned used = 0;  // bit per row
    for (int row = 0; row < ROWS; ++row)
      if ((row == ROW_AS || row == ROW_LS) || (P.velo && (row == ROW_AU || (row == ROW_LU && !P.v2))) ||  // (v2 folds LU into ROW_LS)
          (P.velo && P.grad && (row == ROW_GU || row == ROW_W)) || (P.lginline && row == ROW_PSI) || (P.grad && row >= ROW_DNU))
        used |= 1u << row;
    if (!valid) used = 0;
    double acc[kEpiMaxRows];
#pragma unroll
    for (int row = 0; row < kEpiMaxRows; ++row) acc[row] = 0.0;
    const long long stride = (long long)ROWS * P.ld;
    // all rows of one split are loaded before any is added: ~14 independent loads in flight per thread instead of one
    for (int sidx = ly; sidx < P.n_split; sidx += kEpiLanes) {
      const float* src = P.genepart + sidx * stride + g;
      float v[kEpiMaxRows];
#pragma unroll
      for (int row = 0; row < kEpiMaxRows; ++row) v[row] = (used >> row) & 1u ? __ldg(src + (long long)row * P.ld) : 0.f;
#pragma unroll
      for (int row = 0; row < kEpiMaxRows; ++row) acc[row] += (double)v[row];
    }
#pragma unroll
    for (int row = 0; row < kEpiMaxRows; ++row)
      if (row < ROWS) s_rowpart[row][ly][gx] = acc[row];
  }
  __syncthreads();
  for (int row = ly; row < ROWS; row += kEpiLanes) {
    double t = 0.0;
    for (int l = 0; l < kEpiLanes; ++l) t += s_rowpart[row][l][gx];
    s_rows[row][gx] = t;
  }
  __syncthreads();
  // vcb_stream2: d/dDelta-nu[b][g] = fixed-order sum of the per-split batch sums; their total is d/dnu_0
  double dnu0_v2 = 0.0;
  if (P.v2 && P.grad && P.Nb > 0 && P.dnu_acc != nullptr) {
    for (int b = 0; b < P.Nb; ++b) {
      double s = 0.0;
      if (valid)
        for (int sidx = ly; sidx < P.n_split; sidx += kEpiLanes) s += (double)P.dnu_acc[((long long)sidx * P.Nb + b) * P.ld + g];
      s_part[ly][gx] = s;
      __syncthreads();
      if (ly == 0 && valid) {
        double t = 0.0;
        for (int l = 0; l < kEpiLanes; ++l) t += s_part[l][gx];
        dnu0_v2 += t;
        if (P.d_dnu) P.d_dnu[(long long)b * P.Ng + g] = (float)t;
      }
      __syncthreads();
    }
  }
  double r = 1.0;
  if (valid) r = 1.0 / (double)P.shape_inv[g];
  {
    double a = 0, b = 0, c = 0, d = 0;
    if (valid && !P.lginline) {
      if (P.spec_pre != nullptr) {
        if (ly == 0) {
          a = P.spec_pre[g * 4 + 0];
          b = P.spec_pre[g * 4 + 1];
          c = P.spec_pre[g * 4 + 2];
          d = P.spec_pre[g * 4 + 3];
        }
      } else {
        spectrum_sums(P.spec_S, g, r, ly, kEpiLanes, a, b);
        if (P.velo) spectrum_sums(P.spec_U, g, r, ly, kEpiLanes, c, d);
      }
    }
    s_spec[0][ly][gx] = a;
    s_spec[1][ly][gx] = b;
    s_spec[2][ly][gx] = c;
    s_spec[3][ly][gx] = d;
  }
  __syncthreads();

  if (ly == 0 && valid) {
    double lgS = 0, psS = 0, lgU = 0, psU = 0;
    for (int l = 0; l < kEpiLanes; ++l) {
      lgS += s_spec[0][l][gx];
      psS += s_spec[1][l][gx];
      lgU += s_spec[2][l][gx];
      psU += s_spec[3][l][gx];
    }
    if (!P.lginline) {
      if (P.spec_S.lgk1) lgS -= P.spec_S.lgk1[g];
      if (P.velo && P.spec_U.lgk1) lgU -= P.spec_U.lgk1[g];
    }
    // the streaming kernel works in units of mu/r: LS = sum lg2(1 + mu/r), so n r log r has already cancelled
    const double AS = s_rows[ROW_AS][gx] * kLn2d, LS = s_rows[ROW_LS][gx] * kLn2d;
    const double lpS = P.v2 ? AS + lgS : AS - r * LS + lgS;
    P.lp_S[g] = (float)lpS;
    double LU = 0.0;
    if (P.velo) {
      const double AU = s_rows[ROW_AU][gx] * kLn2d;
      if (!P.v2) LU = s_rows[ROW_LU][gx] * kLn2d;
      P.lp_U[g] = (float)(P.v2 ? AU + lgU : AU - r * LU + lgU);
    }
    if (P.grad) {
      double dnu0 = s_rows[ROW_DNU][gx];
      if (P.dnu_rows) {
        if (P.d_dnu) P.d_dnu[g] = (float)dnu0;
      } else if (P.v2 && P.Nb > 0 && P.dnu_acc != nullptr) {
        dnu0 = dnu0_v2;
      } else if (P.Nb > 0 && P.dnu_acc != nullptr) {
        dnu0 = 0.0;
        for (int b = 0; b < P.Nb; ++b) {
          const float v = P.dnu_acc[(long long)b * P.Ng + g];
          dnu0 += (double)v;
          if (P.d_dnu) P.d_dnu[(long long)b * P.Ng + g] = v;
        }
      }
      if (P.d_nu) {
        P.d_nu[g * K] = (float)dnu0;
        for (int k = 1; k < K; ++k) P.d_nu[g * K + k] = (float)s_rows[ROW_DNU + k][gx];
      }
      const double psi = P.lginline ? s_rows[ROW_PSI][gx] : (psS + psU);
      const double dr = psi - (LS + LU) - dnu0 / r;
      if (P.d_shape_inv) P.d_shape_inv[g] = (float)(-r * r * dr);
      if (P.velo) {
        if (P.d_logbeta) P.d_logbeta[g] = (float)(-s_rows[ROW_GU][gx]);
        if (P.d_gamma) P.d_gamma[g] = (float)s_rows[ROW_W][gx];
      }
    }
  }
  if (blockIdx.x == 0 && P.velo && P.grad && P.d_nu_omega != nullptr) {
    const int n = P.Nx * (2 * P.Hw + 1);
    // fixed-order sum of the cell epilogue's per-block partials, four values at a time: thread t takes blocks t, t+T, ... (the
    // four loads of a block are contiguous), a shuffle tree adds the lanes, then the warps' sums in warp order (deterministic)
    __shared__ double s_red[kEpiGenes * kEpiLanes / 32][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int i0 = 0; i0 < n; i0 += 4) {
      double t[4] = {0.0, 0.0, 0.0, 0.0};
      for (int b = threadIdx.x; b < P.n_cell_blocks; b += blockDim.x) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (i0 + j < n) t[j] += P.dnw_part[(size_t)b * n + i0 + j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        for (int o = 16; o > 0; o >>= 1) t[j] += __shfl_xor_sync(0xffffffffu, t[j], o);
        if (lane == 0) s_red[warp][j] = t[j];
      }
      __syncthreads();
      if (threadIdx.x < 4 && i0 + threadIdx.x < n) {
        double tot = 0.0;
        for (int w = 0; w < nwarps; ++w) tot += s_red[w][threadIdx.x];
        P.d_nu_omega[i0 + threadIdx.x] = (float)tot;
      }
      __syncthreads();
    }
  }
}

// ======================================================================================================
// One-off dataset statistic: dense per-gene histogram of a count matrix.
// A thread walks one gene column of a cell chunk; zeros are counted in a register.
// ======================================================================================================
__global__ void vcb_count_histogram_kernel(const float* __restrict__ M, long long Nc, long long Ng, long long ld, int B,
                                           unsigned int* hist, int* status, int cells_per_block) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Ng) return;
  const long long c0 = (long long)blockIdx.y * cells_per_block;
  long long c1 = c0 + cells_per_block;
  if (c1 > Nc) c1 = Nc;
  unsigned int zeros = 0;
  bool bad = false;
  for (long long c = c0; c < c1; ++c) {
    const float v = M[c * ld + g];
    if (v == 0.f) {
      ++zeros;
    } else {
      if (!(v > 0.f) || v != floorf(v) || v >= (float)(B - 1)) {
        bad = true;
        if (!(v > 0.f)) continue;
      }
      const int k = v >= (float)(B - 1) ? B - 1 : (int)v;
      atomicAdd(&hist[g * B + k], 1u);
    }
  }
  if (zeros) atomicAdd(&hist[g * B], zeros);
  if (bad) atomicExch(status, 1);
}

// ======================================================================================================
// Count staging formats -> float32 (vcb_expand_counts).  Pure streaming: 16 entries per thread and iteration,
// 128-bit loads and stores, grid = a few CTAs per SM.
// ======================================================================================================
template <typename T>
__device__ __forceinline__ float count_to_float(T v) {
  return (float)v;
}

template <typename T>
__global__ void __launch_bounds__(256) vcb_expand_counts_kernel(const T* __restrict__ src, float* __restrict__ dst,
                                                               long long n) {
  constexpr int V = 16 / sizeof(T);  // entries per 16-byte load
  const long long nvec = n / V;
  const uint4* src4 = reinterpret_cast<const uint4*>(src);
  float4* dst4 = reinterpret_cast<float4*>(dst);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 raw = __ldcs(src4 + i);
    T v[V];
    *reinterpret_cast<uint4*>(v) = raw;
#pragma unroll
    for (int j = 0; j < V / 4; ++j)
      __stcs(dst4 + i * (V / 4) + j, make_float4(count_to_float(v[4 * j]), count_to_float(v[4 * j + 1]),
                                                 count_to_float(v[4 * j + 2]), count_to_float(v[4 * j + 3])));
  }
  for (long long i = nvec * V + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = count_to_float(src[i]);
}

// Sub-byte staging -> float32 (format: include/vcb.h).  One thread per 32-bit code word, one CTA per 256-word block; the
// escapes of a block are located with a block-wide exclusive scan of the per-word escape counts.
template <int BITS>
__global__ void __launch_bounds__(VCB_PACKED_BLOCK_WORDS) vcb_expand_packed_kernel(const uint32_t* __restrict__ codes,
                                                                                  const uint8_t* __restrict__ side,
                                                                                  const long long* __restrict__ block_off,
                                                                                  long long n, long long n_blocks,
                                                                                  float* __restrict__ dst) {
  constexpr int PER = 32 / BITS;
  constexpr uint32_t E = (1u << BITS) - 1u;
  __shared__ int s_warp[VCB_PACKED_BLOCK_WORDS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const long long w = blk * VCB_PACKED_BLOCK_WORDS + threadIdx.x;
    const uint32_t word = __ldcs(codes + w);
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) cnt += (((word >> (k * BITS)) & E) == E) ? 1 : 0;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int before = 0;
    for (int i = 0; i < warp; ++i) before += s_warp[i];
    long long pos = block_off[blk] + before + (incl - cnt);
    float v[PER];
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const uint32_t c = (word >> (k * BITS)) & E;
      v[k] = (float)c;
      if (c == E) v[k] = (float)side[pos++];
    }
    const long long e0 = w * PER;
    if (e0 + PER <= n) {
      float4* d4 = reinterpret_cast<float4*>(dst + e0);
#pragma unroll
      for (int j = 0; j < PER / 4; ++j) __stcs(d4 + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
    } else {
      for (int k = 0; k < PER; ++k)
        if (e0 + k < n) dst[e0 + k] = v[k];
    }
    __syncthreads();
  }
}

// block-wide exclusive scan of one int per thread (256 threads); s_warp: 8 ints of shared scratch
__device__ __forceinline__ int block_exclusive_scan_256(int v, int* s_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int before = 0;
  for (int i = 0; i < warp; ++i) before += s_warp[i];
  __syncthreads();
  return before + incl - v;
}

// Two-level 2-bit staging -> float32 (format: include/vcb.h, vcb_expand_counts_twolevel).
__global__ void __launch_bounds__(VCB_PACKED_BLOCK_WORDS) vcb_expand_packed2_kernel(
    const uint32_t* __restrict__ codes, const uint32_t* __restrict__ nibbles, const uint8_t* __restrict__ side,
    const long long* __restrict__ block_off1, const long long* __restrict__ block_off2, long long n, long long n_blocks,
    float* __restrict__ dst) {
  __shared__ int s_warp[VCB_PACKED_BLOCK_WORDS / 32];
  for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    const long long w = blk * VCB_PACKED_BLOCK_WORDS + threadIdx.x;
    const uint32_t word = __ldcs(codes + w);
    int cnt1 = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) cnt1 += (((word >> (2 * k)) & 3u) == 3u) ? 1 : 0;
    long long p1 = block_off1[blk] + block_exclusive_scan_256(cnt1, s_warp);
    // this thread's nibbles, in entry order, and how many of them escape again
    unsigned long long nib = 0ull;
    int cnt2 = 0;
    for (int j = 0; j < cnt1; ++j) {
      const long long i = p1 + j;
      const unsigned v = (nibbles[i >> 3] >> (4 * (int)(i & 7))) & 15u;
      nib |= (unsigned long long)v << (4 * j);
      cnt2 += (v == 15u) ? 1 : 0;
    }
    long long p2 = block_off2[blk] + block_exclusive_scan_256(cnt2, s_warp);
    float v[16];
    int j = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const uint32_t c = (word >> (2 * k)) & 3u;
      float x = (float)c;
      if (c == 3u) {
        const unsigned nv = (unsigned)(nib >> (4 * j)) & 15u;
        ++j;
        x = (nv == 15u) ? (float)side[p2++] : (float)(3u + nv);
      }
      v[k] = x;
    }
    const long long e0 = w * 16;
    if (e0 + 16 <= n) {
      float4* d4 = reinterpret_cast<float4*>(dst + e0);
#pragma unroll
      for (int q = 0; q < 4; ++q) __stcs(d4 + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
    } else {
      for (int k = 0; k < 16; ++k)
        if (e0 + k < n) dst[e0 + k] = v[k];
    }
  }
}

__global__ void vcb_scatter_overflow_kernel(const long long* __restrict__ idx, const float* __restrict__ val,
                                            long long n, float* __restrict__ dst) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[idx[i]] = val[i];
}

// ======================================================================================================
// Fused multi-tensor ClippedAdam (pyro/optim/clipped_adam.py semantics, lr decayed before use).
// ======================================================================================================
// CSR (cells x genes) -> float32 cell-major counts: one warp per cell row, lanes stride over the row's entries.
// atomicAdd on a zeroed row: duplicates sum like scipy's toarray(), and integer-valued sums are exact and order-independent.
template <typename T>
__global__ void __launch_bounds__(256) vcb_csr_to_counts_kernel(const long long* __restrict__ indptr,
                                                                const int* __restrict__ indices, const T* __restrict__ data,
                                                                long long Nc, long long Ng, long long ld,
                                                                float* __restrict__ dst, int* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  bool bad = false;
  for (long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < Nc; row += warps) {
    const long long j0 = indptr[row], j1 = indptr[row + 1];
    float* out = dst + row * ld;
    for (long long j = j0 + lane; j < j1; j += 32) {
      const int g = indices[j];
      const double v = (double)data[j];
      if (g < 0 || g >= Ng || !(v >= 0.0) || v >= 16777216.0 || v != floor(v)) {
        bad = true;
        continue;
      }
      if (v != 0.0) atomicAdd(out + g, (float)v);
    }
  }
  if (bad && status != nullptr) *status = 1;
}

__global__ void vcb_adam_tick_kernel(long long* step) { *step += 1; }

__global__ void vcb_clipped_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                        float* __restrict__ v, long long n, const long long* step_dev, float lr0,
                                        float lrd, float b1, float b2, float eps, float clip) {
  const double step = (double)(*step_dev);
  const double lr = (double)lr0 * pow((double)lrd, step);
  const double bc1 = 1.0 - pow((double)b1, step), bc2 = 1.0 - pow((double)b2, step);
  const float step_size = (float)(lr * sqrt(bc2) / bc1);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i];
    gi = (gi != gi) ? gi : fminf(fmaxf(gi, -clip), clip);  // torch.clamp_ propagates NaN (fminf / fmaxf would not)
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - step_size * (mi / (sqrtf(vi) + eps));
  }
}

// ======================================================================================================
// Host side
// ======================================================================================================
struct Plan {
  int np, nthr, tile_g, n_tiles, n_split, n_ring;
  int tabg, rows, NQ;
  long long n_groups, Ncp;
  size_t off_tab, off_genepart, off_cellpart, off_dnuacc, off_dnwpart, total;
  int n_cell_blocks;
  int smem;
};

// SM count of the CURRENT device (asked every call: a process may drive several GPUs); 148 on a B200
static int sm_count() {
  int dev = 0, v = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
      v > 0)
    return v;
  return 148;
}

// VCB_OK when the current device can run the sm_100a code in this library
static int check_device() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return VCB_ERR_DEVICE;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return VCB_ERR_DEVICE;
  return major == 10 ? VCB_OK : VCB_ERR_DEVICE;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static int pairs_per_warp() {
  // NPAIR = 1: 512 threads, a warp owns 32 genes (<=128 regs); NPAIR = 2: 256 threads, 64 genes per warp (<=255 regs).
  // A build-time choice (-DVCB_DEFAULT_NP=2): the library reads no environment variables.
  return VCB_DEFAULT_NP;
}

constexpr int kSmemBudget = 227 * 1024;  // opt-in dynamic shared memory per CTA on sm_100

static Plan make_plan(const vcb_problem_t* p, bool velo) {
  Plan pl{};
  pl.np = pairs_per_warp();
  const int gpw = 32 * pl.np;  // genes per warp
  const int tmax = max_threads(pl.np);
  // threads per CTA: the widest tile that wastes the fewest lanes
  double best = -1.0;
  for (int t = tmax; t >= 32; t >>= 1) {
    const long long tg = (long long)gpw * (t / 32);
    const long long nt = (p->ld + tg - 1) / tg;
    const double eff = (double)p->ld / (double)(nt * tg);
    if (eff > best + 0.02) {
      best = eff;
      pl.nthr = t;
    }
  }
  pl.tile_g = gpw * (pl.nthr / 32);
  pl.n_tiles = (int)((p->ld + pl.tile_g - 1) / pl.tile_g);
  if (pl.n_tiles < 1) pl.n_tiles = 1;
  pl.n_groups = (p->Nc + kGroupCells - 1) / kGroupCells;
  pl.Ncp = pl.n_groups * kGroupCells;
  const int ctas_per_sm = tmax / pl.nthr;
  long long ns = ((long long)sm_count() * ctas_per_sm) / pl.n_tiles;
  if (ns > pl.n_groups) ns = pl.n_groups;
  if (ns < 1) ns = 1;
  pl.n_split = (int)ns;
  pl.tabg = table_group_floats(p->H, velo);
  pl.rows = gene_rows(p->H);
  pl.NQ = velo ? 3 : 2;
  // ring depth: as deep as the CTA's share of shared memory allows
  pl.n_ring = 2;
  for (int r = kMaxStages; r >= 2; --r)
    if (stream_smem_layout(p->H, velo, pl.nthr / 32, pl.np, r).total <= kSmemBudget / ctas_per_sm - 1024) {
      pl.n_ring = r;
      break;
    }
  pl.smem = stream_smem_layout(p->H, velo, pl.nthr / 32, pl.np, pl.n_ring).total;
  size_t off = 0;
  pl.off_tab = off;
  off = align_up(off + (size_t)pl.n_groups * pl.tabg * 4, 256);
  pl.off_genepart = off;
  off = align_up(off + (size_t)pl.n_split * pl.rows * p->ld * 4, 256);
  pl.off_cellpart = off;
  off = align_up(off + (size_t)pl.n_tiles * pl.NQ * pl.Ncp * 4, 256);
  pl.off_dnuacc = off;
  off = align_up(off + (size_t)(p->Nb > 0 ? p->Nb : 0) * p->Ng * 4, 256);
  pl.off_dnwpart = off;
  pl.n_cell_blocks = (int)((p->Nc + kCellEpiThreads - 1) / kCellEpiThreads);
  off = align_up(off + (size_t)(pl.n_cell_blocks > 0 ? pl.n_cell_blocks : 1) * (p->Nx > 0 ? p->Nx : 1) * (2 * p->Hw + 1) * 8, 256);
  pl.total = off;
  return pl;
}

// ---- round-2 streaming kernel (vcb_stream2.cuh): every call without inline lgamma and with H <= 3 ----------------------
struct Plan2 {
  int n_tiles, n_split, n_ring, tabg, rows, NQ, smem, n_cell_blocks;
  long long n_stages, n_groups, Ncp;
  size_t off_tab, off_genepart, off_cellpart, off_dnupart, off_dnwpart, off_spec, total;
};

static bool stream2_applies(const vcb_problem_t* p) {
  return !(p->flags & (VCB_FLAG_LGAMMA_INLINE | VCB_FLAG_LEGACY_STREAM)) && ksteps(p->H) == 1 && p->Nc > 0;
}

static Plan2 make_plan2(const vcb_problem_t* p, bool velo) {
  Plan2 pl{};
  constexpr int tile_g = 32 * (s2::kThreads / 32);
  pl.n_tiles = (int)((p->ld + tile_g - 1) / tile_g);
  if (pl.n_tiles < 1) pl.n_tiles = 1;
  pl.n_stages = (p->Nc + s2::kStageCells - 1) / s2::kStageCells;
  pl.n_groups = pl.n_stages * s2::kGPS;
  pl.Ncp = pl.n_stages * s2::kStageCells;
  long long ns = (long long)sm_count() / pl.n_tiles;
  if (ns > pl.n_stages) ns = pl.n_stages;
  if (ns < 1) ns = 1;
  pl.n_split = (int)ns;
  pl.tabg = table_group_floats(p->H, velo);
  pl.rows = gene_rows(p->H);
  pl.NQ = velo ? 3 : 2;
  pl.n_ring = s2::ring_depth(p->H, velo);  // the kernel computes the same depth from its template arguments
  pl.smem = s2::smem_layout(p->H, velo, s2::kThreads / 32, pl.n_ring).total;
  size_t off = 0;
  pl.off_tab = off;
  off = align_up(off + (size_t)pl.n_groups * pl.tabg * 4, 256);
  pl.off_genepart = off;
  off = align_up(off + (size_t)pl.n_split * pl.rows * p->ld * 4, 256);
  pl.off_cellpart = off;
  off = align_up(off + (size_t)pl.n_tiles * pl.NQ * pl.Ncp * 4, 256);
  pl.off_dnupart = off;
  off = align_up(off + (size_t)pl.n_split * (p->Nb > 0 ? p->Nb : 0) * p->ld * 4, 256);
  pl.off_dnwpart = off;
  pl.n_cell_blocks = (int)((p->Nc + kCellEpiThreads - 1) / kCellEpiThreads);
  off = align_up(off + (size_t)(pl.n_cell_blocks > 0 ? pl.n_cell_blocks : 1) * (p->Nx > 0 ? p->Nx : 1) * (2 * p->Hw + 1) * 8, 256);
  pl.off_spec = off;
  off = align_up(off + (size_t)p->Ng * 4 * 8, 256);
  pl.total = off;
  return pl;
}

// ---- tcgen05 path (vcb_umma.cuh): velocity model with gradients, H <= 3, at most one batch -----------------------------
struct UmmaPlan {
  int n_tiles, n_split, rows;
  long long n_chunks, Ncp;
  size_t off_tabF, off_tabB, off_omega, off_zero, off_genepart, off_cellpart, off_dnwpart, total;
  int n_cell_blocks;
};

static bool umma_applies(const vcb_problem_t* p, bool velo) {
  return velo && (p->flags & VCB_FLAG_GRAD) && !(p->flags & VCB_FLAG_LGAMMA_INLINE) && p->H <= 3 && p->Nb <= 1 && p->Nc > 0;
}

static UmmaPlan make_plan_umma(const vcb_problem_t* p) {
  UmmaPlan pl{};
  pl.n_tiles = (int)((p->ld + umma::GT - 1) / umma::GT);
  pl.n_chunks = (p->Nc + umma::NC - 1) / umma::NC;
  pl.Ncp = pl.n_chunks * umma::NC;
  long long ns = sm_count() / pl.n_tiles;
  if (ns > pl.n_chunks) ns = pl.n_chunks;
  if (ns < 1) ns = 1;
  pl.n_split = (int)ns;
  pl.rows = gene_rows(p->H);
  size_t off = 0;
  pl.off_tabF = off;
  off = align_up(off + (size_t)pl.n_chunks * umma::TABF_BYTES, 256);
  pl.off_tabB = off;
  off = align_up(off + (size_t)pl.n_chunks * umma::TABB_BYTES, 256);
  pl.off_omega = off;
  off = align_up(off + (size_t)pl.Ncp * 4, 256);
  pl.off_zero = off;
  off = align_up(off + (size_t)umma::GT * 4, 256);
  pl.off_genepart = off;
  off = align_up(off + (size_t)(2 * pl.n_split) * pl.rows * p->ld * 4, 256);
  pl.off_cellpart = off;
  off = align_up(off + (size_t)pl.n_tiles * 3 * pl.Ncp * 4, 256);
  pl.off_dnwpart = off;
  pl.n_cell_blocks = (int)((p->Nc + kCellEpiThreads - 1) / kCellEpiThreads);
  off = align_up(off + (size_t)(pl.n_cell_blocks > 0 ? pl.n_cell_blocks : 1) * (p->Nx > 0 ? p->Nx : 1) * (2 * p->Hw + 1) * 8, 256);
  pl.total = off;
  return pl;
}

cudaError_t vcb_launch_umma_tables(const umma::TableParams& tp, cudaStream_t st);
cudaError_t vcb_launch_umma_stream(const umma::Params& sp, int n_tiles, int n_split, cudaStream_t st);

static int validate(const vcb_problem_t* p, bool velo) {
  if (p == nullptr) return VCB_ERR_NULL;
  if (p->Nc < 0 || p->Ng <= 0 || p->ld < p->Ng || p->Nc > (1LL << 40) || p->Ng > (1LL << 24)) return VCB_ERR_SIZE;
  if (p->H < 0 || p->H > VCB_MAX_HARMONICS) return VCB_ERR_HARMONICS;
  if (p->Nb < 0) return VCB_ERR_SIZE;
  if (p->ld % 4 != 0) return VCB_ERR_ALIGN;
  if (!p->S || !p->phi || !p->nu || !p->shape_inv || !p->lp_S) return VCB_ERR_NULL;
  if (((uintptr_t)p->S & 15) != 0) return VCB_ERR_ALIGN;
  if (p->Nb > 0 && (!p->dnu || !p->batch_id)) return VCB_ERR_NULL;
  if (velo) {
    if (p->Hw < 0 || p->Hw > VCB_MAX_HARMONICS) return VCB_ERR_HARMONICS;
    if (p->Nx < 1) return VCB_ERR_SIZE;
    if (!p->U || !p->logbeta || !p->gamma || !p->nu_omega || !p->lp_U) return VCB_ERR_NULL;
    if (((uintptr_t)p->U & 15) != 0) return VCB_ERR_ALIGN;
    if (p->Nx > 1 && !p->cond_id) return VCB_ERR_NULL;
  }
  if (!(p->flags & VCB_FLAG_LGAMMA_INLINE)) {
    if (!p->spec_S.off || !p->spec_S.val || !p->spec_S.mult || !p->spec_S.lgk1) return VCB_ERR_SPECTRUM;
    if (velo && (!p->spec_U.off || !p->spec_U.val || !p->spec_U.mult || !p->spec_U.lgk1)) return VCB_ERR_SPECTRUM;
  }
  return VCB_OK;
}

// one translation unit per H (vcb_stream_inst.cu, compiled with -DVCB_INST_H=h) so the build parallelises
#define VCB_DECL_LAUNCH(h)                                                                                     \
  cudaError_t vcb_launch_stream_h##h(bool velo, bool grad, bool lgi, int np, const StreamParams& sp, dim3 grid, \
                                     int nthr, int smem, cudaStream_t st);
VCB_DECL_LAUNCH(0)
VCB_DECL_LAUNCH(1)
VCB_DECL_LAUNCH(2)
VCB_DECL_LAUNCH(3)
VCB_DECL_LAUNCH(4)
VCB_DECL_LAUNCH(5)
#undef VCB_DECL_LAUNCH

#define VCB_DECL_LAUNCH2(h) \
  cudaError_t vcb_launch_stream2_h##h(bool velo, bool grad, const s2::Params& sp, dim3 grid, int nthr, int smem, cudaStream_t st);
VCB_DECL_LAUNCH2(0)
VCB_DECL_LAUNCH2(1)
VCB_DECL_LAUNCH2(2)
VCB_DECL_LAUNCH2(3)
#undef VCB_DECL_LAUNCH2

static cudaError_t launch_stream2(int H, bool velo, bool grad, const s2::Params& sp, dim3 grid, int smem, cudaStream_t st) {
  switch (H) {
    case 0: return vcb_launch_stream2_h0(velo, grad, sp, grid, s2::kThreads, smem, st);
    case 1: return vcb_launch_stream2_h1(velo, grad, sp, grid, s2::kThreads, smem, st);
    case 2: return vcb_launch_stream2_h2(velo, grad, sp, grid, s2::kThreads, smem, st);
    case 3: return vcb_launch_stream2_h3(velo, grad, sp, grid, s2::kThreads, smem, st);
    default: return cudaErrorInvalidValue;
  }
}

static cudaError_t launch_stream(int H, bool velo, bool grad, bool lgi, int np, const StreamParams& sp, dim3 grid,
                                 int nthr, int smem, cudaStream_t st) {
  switch (H) {
    case 0: return vcb_launch_stream_h0(velo, grad, lgi, np, sp, grid, nthr, smem, st);
    case 1: return vcb_launch_stream_h1(velo, grad, lgi, np, sp, grid, nthr, smem, st);
    case 2: return vcb_launch_stream_h2(velo, grad, lgi, np, sp, grid, nthr, smem, st);
    case 3: return vcb_launch_stream_h3(velo, grad, lgi, np, sp, grid, nthr, smem, st);
    case 4: return vcb_launch_stream_h4(velo, grad, lgi, np, sp, grid, nthr, smem, st);
    case 5: return vcb_launch_stream_h5(velo, grad, lgi, np, sp, grid, nthr, smem, st);
    default: return cudaErrorInvalidValue;
  }
}

// The tcgen05 path: table kernel -> streaming kernel -> the same two epilogues (they only see partial-sum buffers).
static int run_umma(const vcb_problem_t* p, void* workspace, size_t ws_bytes, void* stream) {
  const UmmaPlan pl = make_plan_umma(p);
  if (ws_bytes < pl.total) return VCB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* ws = (unsigned char*)workspace;
  float* tabF = (float*)(ws + pl.off_tabF);
  float* tabB = (float*)(ws + pl.off_tabB);
  float* omega = (float*)(ws + pl.off_omega);
  float* zero = (float*)(ws + pl.off_zero);
  float* genepart = (float*)(ws + pl.off_genepart);
  float* cellpart = (float*)(ws + pl.off_cellpart);
  double* dnw_part = (double*)(ws + pl.off_dnwpart);
  cudaError_t e;
  umma::TableParams tp{p->phi, p->cf, p->cond_id, p->nu_omega, tabF, tabB, omega, zero, p->Nc, pl.n_chunks, p->H, p->Hw};
  e = vcb_launch_umma_tables(tp, st);
  if (e != cudaSuccess) return (int)e;
  {
    umma::Params sp{p->S, p->U, tabF, tabB, omega, zero, p->nu, p->Nb > 0 ? p->dnu : nullptr, p->shape_inv, p->logbeta, p->gamma,
                    genepart, cellpart, p->Nc, p->Ng, p->ld, pl.Ncp, pl.n_split, p->H, pl.rows, nullptr,
                    0};
#ifdef VCB_UMMA_INSTRUMENT
    if (getenv("VCB_UMMA_DEBUG")) sp.debug = atoi(getenv("VCB_UMMA_DEBUG"));
#endif
#ifdef VCB_UMMA_INSTRUMENT
    static long long* trace_buf = nullptr;  // instrumented builds only (VCB_UMMA_TRACE): synchronises and prints
    if (getenv("VCB_UMMA_TRACE") != nullptr) {
      if (trace_buf == nullptr) cudaMalloc(&trace_buf, 64 * 8 * sizeof(long long));
      cudaMemset(trace_buf, 0, 64 * 8 * sizeof(long long));
      sp.trace = trace_buf;
    }
#endif
    unsigned ev_flags = cudaEventRecordDefault;
    if (p->ev_stream_begin || p->ev_stream_end) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive)
        ev_flags = cudaEventRecordExternal;
    }
    if (p->ev_stream_begin) cudaEventRecordWithFlags((cudaEvent_t)p->ev_stream_begin, st, ev_flags);
    e = vcb_launch_umma_stream(sp, pl.n_tiles, pl.n_split, st);
    if (e != cudaSuccess) return (int)e;
    if (p->ev_stream_end) cudaEventRecordWithFlags((cudaEvent_t)p->ev_stream_end, st, ev_flags);
#ifdef VCB_UMMA_INSTRUMENT
    if (sp.trace != nullptr) {
      static int printed = 0;
      long long h[64 * 8];
      cudaStreamSynchronize(st);
      cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
      if (printed++ == 2)
        for (int c = 0; c < 40; ++c) {
          printf("chunk %3d:", 64 + c);
          for (int k = 0; k < 8; ++k) printf(" %7lld", h[c * 8 + k] ? h[c * 8 + k] - h[0] : -1LL);
          printf("\n");
        }
    }
#endif
  }
  {
    CellEpiParams ce{cellpart, p->phi, p->cond_id, p->nu_omega, p->d_phi, p->d_cf, p->d_omega,
                     dnw_part, p->Nc,  pl.Ncp, pl.n_tiles, 3, p->Hw, p->Nx, 1.f};
    const int bs = kCellEpiThreads;
    const size_t sm = (size_t)(bs / 32) * p->Nx * (2 * p->Hw + 1) * 8;
    vcb_cell_epilogue_kernel<<<(unsigned)pl.n_cell_blocks, bs, sm, st>>>(ce);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  {
    GeneEpiParams ge{};
    ge.genepart = genepart;
    ge.shape_inv = p->shape_inv;
    ge.dnu_acc = nullptr;
    ge.dnw_part = dnw_part;
    ge.n_cell_blocks = pl.n_cell_blocks;
    ge.spec_S = p->spec_S;
    ge.spec_U = p->spec_U;
    ge.lp_S = p->lp_S;
    ge.lp_U = p->lp_U;
    ge.d_nu = p->d_nu;
    ge.d_dnu = p->d_dnu;
    ge.d_shape_inv = p->d_shape_inv;
    ge.d_logbeta = p->d_logbeta;
    ge.d_gamma = p->d_gamma;
    ge.d_nu_omega = p->d_nu_omega;
    ge.Nc = p->Nc;
    ge.Ng = p->Ng;
    ge.ld = p->ld;
    ge.n_split = 2 * pl.n_split;
    ge.H = p->H;
    ge.Nb = p->Nb;
    ge.Nx = p->Nx;
    ge.Hw = p->Hw;
    ge.velo = 1;
    ge.grad = 1;
    ge.lginline = 0;
    ge.dnu_rows = p->Nb > 0 ? 1 : 0;
    vcb_gene_epilogue_kernel<<<(unsigned)((p->Ng + kEpiGenes - 1) / kEpiGenes), kEpiGenes * kEpiLanes, 0, st>>>(ge);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  return VCB_OK;
}

static unsigned event_flags(cudaStream_t st) {  // timing events inside a stream capture must become event-record NODES
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive) return cudaEventRecordExternal;
  return cudaEventRecordDefault;
}

// table kernel -> vcb_stream2 -> cell epilogue -> gene epilogue
static int run2(const vcb_problem_t* p, bool velo, void* workspace, size_t ws_bytes, void* stream) {
  const Plan2 pl = make_plan2(p, velo);
  if (ws_bytes < pl.total) return VCB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const bool grad = (p->flags & VCB_FLAG_GRAD) != 0;
  unsigned char* ws = (unsigned char*)workspace;
  float* tab = (float*)(ws + pl.off_tab);
  float* genepart = (float*)(ws + pl.off_genepart);
  float* cellpart = (float*)(ws + pl.off_cellpart);
  float* dnupart = (grad && p->Nb > 0) ? (float*)(ws + pl.off_dnupart) : nullptr;
  double* dnw_part = (double*)(ws + pl.off_dnwpart);
  double* spec_pre = (double*)(ws + pl.off_spec);
  cudaError_t e;
  if (dnupart) {
    e = cudaMemsetAsync(dnupart, 0, (size_t)pl.n_split * p->Nb * p->ld * 4, st);
    if (e != cudaSuccess) return (int)e;
  }
  {
    CellParams cp{p->phi, p->cf, p->Nb > 0 ? p->batch_id : nullptr, p->cond_id, velo ? p->nu_omega : nullptr,
                  tab,    p->Nc, pl.n_groups, p->H, p->Hw, p->Nx, velo ? 1 : 0, pl.tabg, 1};
    const unsigned n_table_blocks = (unsigned)((pl.n_groups + kTabWarps * kTab4Groups - 1) / (kTabWarps * kTab4Groups));
    const unsigned n_spec_blocks = (unsigned)((p->Ng + kEpiGenes - 1) / kEpiGenes);
    cp.n_spec_blocks = n_spec_blocks;
    cp.spec_S = p->spec_S;
    if (velo) cp.spec_U = p->spec_U;
    cp.shape_inv = p->shape_inv;
    cp.spec_out = spec_pre;
    cp.Ng = p->Ng;
    vcb_cell_tables4_kernel<<<n_table_blocks + n_spec_blocks, kTabWarps * 32, 0, st>>>(cp);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  {
    s2::Params sp{p->S,    velo ? p->U : nullptr, tab,      p->nu,    p->Nb > 0 ? p->dnu : nullptr, p->shape_inv, p->logbeta,
                  p->gamma, genepart,              cellpart, dnupart,  p->Nc,                        p->Ng,        p->ld,
                  pl.Ncp,  pl.n_stages,            pl.n_split, p->Nb, pl.n_ring};
    dim3 grid((unsigned)pl.n_tiles, (unsigned)pl.n_split);
    const unsigned ev_flags = (p->ev_stream_begin || p->ev_stream_end) ? event_flags(st) : 0u;
    if (p->ev_stream_begin) cudaEventRecordWithFlags((cudaEvent_t)p->ev_stream_begin, st, ev_flags);
    e = launch_stream2(p->H, velo, grad, sp, grid, pl.smem, st);
    if (e != cudaSuccess) return (int)e;
    if (p->ev_stream_end) cudaEventRecordWithFlags((cudaEvent_t)p->ev_stream_end, st, ev_flags);
  }
  if (grad) {
    CellEpiParams ce{cellpart, p->phi, p->cond_id, p->nu_omega, p->d_phi, p->d_cf, p->d_omega,
                     dnw_part, p->Nc,  pl.Ncp, pl.n_tiles, pl.NQ, p->Hw, velo ? p->Nx : 0, kLn2};
    const int bs = kCellEpiThreads;
    const size_t sm = velo ? (size_t)(bs / 32) * p->Nx * (2 * p->Hw + 1) * 8 : 0;
    vcb_cell_epilogue_kernel<<<(unsigned)pl.n_cell_blocks, bs, sm, st>>>(ce);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  {
    GeneEpiParams ge{};
    ge.genepart = genepart;
    ge.shape_inv = p->shape_inv;
    ge.dnu_acc = dnupart;
    ge.dnw_part = dnw_part;
    ge.n_cell_blocks = grad ? pl.n_cell_blocks : 0;
    ge.spec_S = p->spec_S;
    ge.spec_U = p->spec_U;
    ge.lp_S = p->lp_S;
    ge.lp_U = p->lp_U;
    ge.d_nu = p->d_nu;
    ge.d_dnu = p->d_dnu;
    ge.d_shape_inv = p->d_shape_inv;
    ge.d_logbeta = p->d_logbeta;
    ge.d_gamma = p->d_gamma;
    ge.d_nu_omega = p->d_nu_omega;
    ge.Nc = p->Nc;
    ge.Ng = p->Ng;
    ge.ld = p->ld;
    ge.n_split = pl.n_split;
    ge.H = p->H;
    ge.Nb = p->Nb;
    ge.Nx = p->Nx;
    ge.Hw = p->Hw;
    ge.velo = velo;
    ge.grad = grad;
    ge.lginline = 0;
    ge.v2 = 1;
    ge.spec_pre = spec_pre;
    vcb_gene_epilogue_kernel<<<(unsigned)((p->Ng + kEpiGenes - 1) / kEpiGenes), kEpiGenes * kEpiLanes, 0, st>>>(ge);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  return VCB_OK;
}

static int run(const vcb_problem_t* p, bool velo, void* workspace, size_t ws_bytes, void* stream) {
  int rc = validate(p, velo);
  if (rc != VCB_OK) return rc;
  if (workspace == nullptr) return VCB_ERR_NULL;
  if (((uintptr_t)workspace & 15) != 0) return VCB_ERR_ALIGN;
  rc = check_device();
  if (rc != VCB_OK) return rc;
  if (!(p->flags & VCB_FLAG_TCGEN05) && stream2_applies(p)) return run2(p, velo, workspace, ws_bytes, stream);
  if ((p->flags & VCB_FLAG_TCGEN05) && umma_applies(p, velo)) return run_umma(p, workspace, ws_bytes, stream);
  const Plan pl = make_plan(p, velo);
  if (ws_bytes < pl.total) return VCB_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const bool grad = (p->flags & VCB_FLAG_GRAD) != 0;
  const bool lgi = (p->flags & VCB_FLAG_LGAMMA_INLINE) != 0;
  unsigned char* ws = (unsigned char*)workspace;
  float* tab = (float*)(ws + pl.off_tab);
  float* genepart = (float*)(ws + pl.off_genepart);
  float* cellpart = (float*)(ws + pl.off_cellpart);
  float* dnu_acc = p->Nb > 0 ? (float*)(ws + pl.off_dnuacc) : nullptr;
  double* dnw_part = (double*)(ws + pl.off_dnwpart);
  cudaError_t e;

  if (grad) {
    if (dnu_acc) {
      e = cudaMemsetAsync(dnu_acc, 0, (size_t)p->Nb * p->Ng * 4, st);
      if (e != cudaSuccess) return (int)e;
    }
  }
  if (p->Nc > 0) {
    CellParams cp{p->phi, p->cf, p->Nb > 0 ? p->batch_id : nullptr, p->cond_id, velo ? p->nu_omega : nullptr,
                  tab,    p->Nc, pl.n_groups, p->H, p->Hw, p->Nx, velo ? 1 : 0, pl.tabg, 0};
    vcb_cell_tables_kernel<<<(unsigned)((pl.n_groups + kTabWarps - 1) / kTabWarps), kTabWarps * 32, 0, st>>>(cp);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  {
    StreamParams sp{p->S,     velo ? p->U : nullptr, tab,   p->nu,  p->Nb > 0 ? p->dnu : nullptr,
                    p->shape_inv, p->logbeta,        p->gamma, genepart, cellpart,
                    dnu_acc,  p->Nc,                 p->Ng, p->ld,  pl.Ncp, pl.n_split,
                    p->Nb, pl.n_ring};
    dim3 grid((unsigned)pl.n_tiles, (unsigned)pl.n_split);
    // timing events: inside a stream capture they must become event-record NODES (external flag)
    unsigned ev_flags = cudaEventRecordDefault;
    if (p->ev_stream_begin || p->ev_stream_end) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive)
        ev_flags = cudaEventRecordExternal;
    }
    if (p->ev_stream_begin) cudaEventRecordWithFlags((cudaEvent_t)p->ev_stream_begin, st, ev_flags);
    e = launch_stream(p->H, velo, grad, lgi, pl.np, sp, grid, pl.nthr, pl.smem, st);
    if (e != cudaSuccess) return (int)e;
    if (p->ev_stream_end) cudaEventRecordWithFlags((cudaEvent_t)p->ev_stream_end, st, ev_flags);
  }
  if (grad && p->Nc > 0) {
    CellEpiParams ce{cellpart, p->phi, p->cond_id, p->nu_omega, p->d_phi, p->d_cf, p->d_omega,
                     dnw_part, p->Nc,  pl.Ncp, pl.n_tiles, pl.NQ, p->Hw, velo ? p->Nx : 0, 1.f};
    const int bs = kCellEpiThreads;
    const size_t sm = velo ? (size_t)(bs / 32) * p->Nx * (2 * p->Hw + 1) * 8 : 0;
    vcb_cell_epilogue_kernel<<<(unsigned)pl.n_cell_blocks, bs, sm, st>>>(ce);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  {
    GeneEpiParams ge{};
    ge.genepart = genepart;
    ge.shape_inv = p->shape_inv;
    ge.dnu_acc = grad ? dnu_acc : nullptr;
    ge.dnw_part = dnw_part;
    ge.n_cell_blocks = (grad && p->Nc > 0) ? pl.n_cell_blocks : 0;
    ge.spec_S = p->spec_S;
    ge.spec_U = p->spec_U;
    ge.lp_S = p->lp_S;
    ge.lp_U = p->lp_U;
    ge.d_nu = p->d_nu;
    ge.d_dnu = p->d_dnu;
    ge.d_shape_inv = p->d_shape_inv;
    ge.d_logbeta = p->d_logbeta;
    ge.d_gamma = p->d_gamma;
    ge.d_nu_omega = p->d_nu_omega;
    ge.Nc = p->Nc;
    ge.Ng = p->Ng;
    ge.ld = p->ld;
    ge.n_split = pl.n_split;
    ge.H = p->H;
    ge.Nb = p->Nb;
    ge.Nx = p->Nx;
    ge.Hw = p->Hw;
    ge.velo = velo;
    ge.grad = grad;
    ge.lginline = lgi;
    vcb_gene_epilogue_kernel<<<(unsigned)((p->Ng + kEpiGenes - 1) / kEpiGenes), kEpiGenes * kEpiLanes, 0, st>>>(ge);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  return VCB_OK;
}

}  // namespace vcb

// ======================================================================================================
// C ABI
// ======================================================================================================
extern "C" {

int vcb_version(void) { return VCB_VERSION; }

const char* vcb_strerror(int code) {
  switch (code) {
    case VCB_OK: return "ok";
    case VCB_ERR_NULL: return "vcb: a required pointer is NULL";
    case VCB_ERR_SIZE: return "vcb: a size argument is out of range";
    case VCB_ERR_ALIGN: return "vcb: ld must be a multiple of 4 and S/U/workspace 16-byte aligned";
    case VCB_ERR_HARMONICS: return "vcb: number of harmonics above VCB_MAX_HARMONICS";
    case VCB_ERR_WORKSPACE: return "vcb: workspace smaller than vcb_workspace_bytes()";
    case VCB_ERR_SPECTRUM: return "vcb: count spectrum missing (or pass VCB_FLAG_LGAMMA_INLINE)";
    case VCB_ERR_DEVICE: return "vcb: no sm_100 CUDA device";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "vcb: unknown error";
}

size_t vcb_workspace_bytes(const vcb_problem_t* p) {
  if (p == nullptr || p->Ng <= 0 || p->ld < p->Ng || p->Nc < 0 || p->H < 0 || p->H > VCB_MAX_HARMONICS || p->Hw < 0 ||
      p->Hw > VCB_MAX_HARMONICS)
    return 0;
  size_t n = vcb::make_plan(p, p->U != nullptr).total;
  if (vcb::stream2_applies(p)) {
    const size_t m2 = vcb::make_plan2(p, p->U != nullptr).total;
    if (m2 > n) n = m2;
  }
  if (p->U != nullptr && p->H <= 3 && p->Nb <= 1 && p->Nc > 0) {  // either streaming kernel may serve the call
    const size_t m = vcb::make_plan_umma(p).total;
    if (m > n) n = m;
  }
  return n;
}

int vcb_phase_fwd_bwd(const vcb_problem_t* p, void* workspace, size_t workspace_bytes, void* stream) {
  vcb::DeviceGuard guard(p ? p->S : nullptr);
  return vcb::run(p, false, workspace, workspace_bytes, stream);
}

int vcb_velocity_fwd_bwd(const vcb_problem_t* p, void* workspace, size_t workspace_bytes, void* stream) {
  vcb::DeviceGuard guard(p ? p->S : nullptr);
  return vcb::run(p, true, workspace, workspace_bytes, stream);
}

int vcb_count_histogram(const float* M, int64_t Nc, int64_t Ng, int64_t ld, int32_t B, uint32_t* hist, int32_t* status,
                        void* stream) {
  vcb::DeviceGuard guard(M);
  if (!M || !hist || !status) return VCB_ERR_NULL;
  if (Nc < 0 || Ng <= 0 || ld < Ng || B < 2) return VCB_ERR_SIZE;
  if (Nc == 0) return VCB_OK;
  const int bs = 128;
  const int cells_per_block = 512;
  dim3 grid((unsigned)((Ng + bs - 1) / bs), (unsigned)((Nc + cells_per_block - 1) / cells_per_block));
  vcb::vcb_count_histogram_kernel<<<grid, bs, 0, (cudaStream_t)stream>>>(M, Nc, Ng, ld, B, hist, status,
                                                                        cells_per_block);
  return (int)cudaGetLastError();
}

int vcb_expand_counts(const void* src, int32_t src_dtype, int64_t n, float* dst, const int64_t* over_idx,
                      const float* over_val, int64_t n_over, void* stream) {
  vcb::DeviceGuard guard(dst);
  if (!src || !dst) return VCB_ERR_NULL;
  if (n < 0 || n_over < 0) return VCB_ERR_SIZE;
  if (n_over > 0 && (!over_idx || !over_val || src_dtype != VCB_COUNTS_U8)) return VCB_ERR_NULL;
  if ((((uintptr_t)src) & 15) != 0 || (((uintptr_t)dst) & 15) != 0) return VCB_ERR_ALIGN;
  if (n == 0) return VCB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int bs = 256;
  const unsigned blocks = (unsigned)vcb::sm_count() * 8;
  switch (src_dtype) {
    case VCB_COUNTS_U8: vcb::vcb_expand_counts_kernel<uint8_t><<<blocks, bs, 0, st>>>((const uint8_t*)src, dst, n); break;
    case VCB_COUNTS_U16: vcb::vcb_expand_counts_kernel<uint16_t><<<blocks, bs, 0, st>>>((const uint16_t*)src, dst, n); break;
    case VCB_COUNTS_I32: vcb::vcb_expand_counts_kernel<int32_t><<<blocks, bs, 0, st>>>((const int32_t*)src, dst, n); break;
    default: return VCB_ERR_SIZE;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if (n_over > 0) {
    long long b = (n_over + bs - 1) / bs;
    if (b > blocks) b = blocks;
    vcb::vcb_scatter_overflow_kernel<<<(unsigned)b, bs, 0, st>>>((const long long*)over_idx, over_val, n_over, dst);
    e = cudaGetLastError();
  }
  return (int)e;
}

int vcb_expand_counts_packed(const uint32_t* codes, int32_t bits, const uint8_t* side, const int64_t* block_off, int64_t n,
                             float* dst, const int64_t* over_idx, const float* over_val, int64_t n_over, void* stream) {
  vcb::DeviceGuard guard(dst);
  if (!codes || !block_off || !dst) return VCB_ERR_NULL;
  if (n < 0 || n_over < 0 || (bits != 2 && bits != 4)) return VCB_ERR_SIZE;
  if (n_over > 0 && (!over_idx || !over_val)) return VCB_ERR_NULL;
  if ((((uintptr_t)codes) & 3) != 0 || (((uintptr_t)dst) & 15) != 0) return VCB_ERR_ALIGN;
  if (n == 0) return VCB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const long long per_block = (long long)VCB_PACKED_BLOCK_WORDS * (32 / bits);
  const long long n_blocks = (n + per_block - 1) / per_block;
  long long grid = n_blocks < (long long)vcb::sm_count() * 16 ? n_blocks : (long long)vcb::sm_count() * 16;
  if (bits == 2)
    vcb::vcb_expand_packed_kernel<2><<<(unsigned)grid, VCB_PACKED_BLOCK_WORDS, 0, st>>>(codes, side, (const long long*)block_off, n, n_blocks, dst);
  else
    vcb::vcb_expand_packed_kernel<4><<<(unsigned)grid, VCB_PACKED_BLOCK_WORDS, 0, st>>>(codes, side, (const long long*)block_off, n, n_blocks, dst);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if (n_over > 0) {
    const int bs = 256;
    long long b = (n_over + bs - 1) / bs;
    if (b > vcb::sm_count() * 8) b = vcb::sm_count() * 8;
    vcb::vcb_scatter_overflow_kernel<<<(unsigned)b, bs, 0, st>>>((const long long*)over_idx, over_val, n_over, dst);
    e = cudaGetLastError();
  }
  return (int)e;
}

int vcb_expand_counts_twolevel(const uint32_t* codes, const uint32_t* nibbles, const uint8_t* side, const int64_t* block_off1,
                              const int64_t* block_off2, int64_t n, float* dst, const int64_t* over_idx, const float* over_val,
                              int64_t n_over, void* stream) {
  vcb::DeviceGuard guard(dst);
  if (!codes || !nibbles || !side || !block_off1 || !block_off2 || !dst) return VCB_ERR_NULL;
  if (n < 0 || n_over < 0) return VCB_ERR_SIZE;
  if (n_over > 0 && (!over_idx || !over_val)) return VCB_ERR_NULL;
  if ((((uintptr_t)codes) & 3) != 0 || (((uintptr_t)nibbles) & 3) != 0 || (((uintptr_t)dst) & 15) != 0) return VCB_ERR_ALIGN;
  if (n == 0) return VCB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const long long per_block = (long long)VCB_PACKED_BLOCK_WORDS * 16;
  const long long n_blocks = (n + per_block - 1) / per_block;
  const long long grid = n_blocks < (long long)vcb::sm_count() * 16 ? n_blocks : (long long)vcb::sm_count() * 16;
  vcb::vcb_expand_packed2_kernel<<<(unsigned)grid, VCB_PACKED_BLOCK_WORDS, 0, st>>>(
      codes, nibbles, side, (const long long*)block_off1, (const long long*)block_off2, n, n_blocks, dst);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if (n_over > 0) {
    const int bs = 256;
    long long b = (n_over + bs - 1) / bs;
    if (b > vcb::sm_count() * 8) b = vcb::sm_count() * 8;
    vcb::vcb_scatter_overflow_kernel<<<(unsigned)b, bs, 0, st>>>((const long long*)over_idx, over_val, n_over, dst);
    e = cudaGetLastError();
  }
  return (int)e;
}

int vcb_csr_to_counts(const int64_t* indptr, const int32_t* indices, const void* data, int32_t data_dtype, int64_t Nc,
                      int64_t Ng, int64_t ld, float* dst, int32_t* status, void* stream) {
  vcb::DeviceGuard guard(dst);
  if (!indptr || !dst) return VCB_ERR_NULL;
  if (Nc < 0 || Ng <= 0 || ld < Ng) return VCB_ERR_SIZE;
  if (ld % 4 != 0 || (((uintptr_t)dst) & 15) != 0) return VCB_ERR_ALIGN;
  if (Nc == 0) return VCB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(dst, 0, (size_t)Nc * (size_t)ld * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  if (status) {
    e = cudaMemsetAsync(status, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return (int)e;
  }
  if (!indices || !data) return VCB_OK;  // an all-zero matrix (nnz = 0) may come without index / value arrays
  const int bs = 256;
  long long blocks = (Nc + (bs / 32) - 1) / (bs / 32);
  if (blocks > vcb::sm_count() * 16) blocks = vcb::sm_count() * 16;
  const long long* ip = (const long long*)indptr;
  switch (data_dtype) {
    case VCB_CSR_F32: vcb::vcb_csr_to_counts_kernel<float><<<(unsigned)blocks, bs, 0, st>>>(ip, indices, (const float*)data, Nc, Ng, ld, dst, status); break;
    case VCB_CSR_I32: vcb::vcb_csr_to_counts_kernel<int><<<(unsigned)blocks, bs, 0, st>>>(ip, indices, (const int*)data, Nc, Ng, ld, dst, status); break;
    case VCB_CSR_F64: vcb::vcb_csr_to_counts_kernel<double><<<(unsigned)blocks, bs, 0, st>>>(ip, indices, (const double*)data, Nc, Ng, ld, dst, status); break;
    case VCB_CSR_I64: vcb::vcb_csr_to_counts_kernel<long long><<<(unsigned)blocks, bs, 0, st>>>(ip, indices, (const long long*)data, Nc, Ng, ld, dst, status); break;
    default: return VCB_ERR_SIZE;
  }
  return (int)cudaGetLastError();
}

int vcb_clipped_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int64_t* step_dev,
                     float lr0, float lrd, float beta1, float beta2, float eps, float clip, void* stream) {
  vcb::DeviceGuard guard(param);
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step_dev) return VCB_ERR_NULL;
  if (n < 0) return VCB_ERR_SIZE;
  cudaStream_t st = (cudaStream_t)stream;
  vcb::vcb_adam_tick_kernel<<<1, 1, 0, st>>>((long long*)step_dev);
  if (n > 0) {
    const int bs = 256;
    long long blocks = (n + bs - 1) / bs;
    if (blocks > vcb::sm_count() * 8) blocks = vcb::sm_count() * 8;
    vcb::vcb_clipped_adam_kernel<<<(unsigned)blocks, bs, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n,
                                                                 (const long long*)step_dev, lr0, lrd, beta1, beta2,
                                                                 eps, clip);
  }
  return (int)cudaGetLastError();
}

}  // extern "C"
