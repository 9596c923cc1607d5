// libvcb: the non-likelihood part of one Trace_ELBO step of the package's own model / guide pairs, fused.
//
// pyro.infer.SVI.step traces guide(mp) and model(mp) through effect handlers: per sample site a reparameterised draw, two
// log_prob evaluations (prior and guide) and their autograd backward -- ~460 tiny launches per step for the velocity model
// (profiles/r01_launch_list_step.csv), 1.6 ms of a 7.1 ms step at 1M cells and 3x the likelihood at 100k cells.  The sites,
// priors and guide families are fixed by the reference (phase_inference_model.py:361-366,391, phase_inference_guide.py:36-56,
// velocity_inference_model.py:323-353,383 / :390-471, velocity_inference_guide.py:25-63 / :65-141), so the same arithmetic
// is done here by four kernels around the likelihood call:
//
//   vcb_svi_sample    cell kernel : phixy = phixy_locs + eps, phi = atan2(y, x), block partials of log p - log q
//                     gene kernel : nu, Delta-nu, shape_inv, (log gamma, gamma, log beta, nu_omega) from the guide parameters
//                                   and the given standard-normal draws; block partials of log p - log q
//   ... vcb_phase_fwd_bwd / vcb_velocity_fwd_bwd (and the all-reduce under cell sharding) ...
//   vcb_svi_backward  cell kernel : d loss / d phixy_locs from d/dphi and the prior
//                     gene kernel : d loss / d (every gene-level / global guide parameter) from the likelihood gradients,
//                                   priors and guide entropy terms, written into the flat gradient buffer
//                     finalize    : loss = -(sum lp_S + sum lp_U + log p(z) - log q(z)), fixed summation order
//
// The noise is an INPUT (drawn by the caller with torch's generator in the guide's draw order), so a step consumes the RNG
// stream exactly like the traced guide and like the reference under Pyro.  Parameters live in one flat buffer in
// unconstrained form (positive ones as logs, transform_to(positive) = exp), gradients are those of the loss (= -ELBO) with
// respect to the unconstrained values, like the .grad that Trace_ELBO leaves for the optimizer.
#include "vcb.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "vcb_common.cuh"

namespace vcb {

constexpr int kSviThreads = 256;
constexpr double kHalfLog2Pi = 0.91893853320467274178;

__device__ __forceinline__ double normal_lp(float x, float mu, float sd) {
  const double z = ((double)x - (double)mu) / (double)sd;
  return -0.5 * z * z - log((double)sd) - kHalfLog2Pi;
}

// fixed-order block sum of a double (deterministic); valid in thread 0
__device__ double block_sum(double v) {
  __shared__ double s[kSviThreads / 32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) s[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s[w];
  return t;
}

// ---- cells ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSviThreads) vcb_svi_cell_sample_kernel(const vcb_svi_t P) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double lp = 0.0;
  if (c < P.Nc) {
    const float2 loc = reinterpret_cast<const float2*>(P.param + P.o_phixy_locs)[c];
    const float2 e = reinterpret_cast<const float2*>(P.eps_phixy)[c];
    const float2 pr = reinterpret_cast<const float2*>(P.phixy_prior)[c];
    float x = loc.x + e.x, y = loc.y + e.y;  // Normal(locs, 1).rsample()
    if (P.cond_phixy != nullptr) {
      const float2 v = reinterpret_cast<const float2*>(P.cond_phixy)[c];
      x = v.x;
      y = v.y;
    }
    reinterpret_cast<float2*>(P.phixy)[c] = make_float2(x, y);
    P.phi[P.cell_row ? P.cell_row[c] : c] = atan2f(y, x);  // pack_direction, utils.py:488-506
    const double dx = (double)x - pr.x, dy = (double)y - pr.y;
    // log p - log q: Normal(prior, 1) against Normal(locs, 1) at locs + eps; the 2 x 1/2 log 2pi cancel.  Conditioned: the
    // prior term alone, constants included
    lp = P.cond_phixy != nullptr ? -0.5 * (dx * dx + dy * dy) - 2.0 * kHalfLog2Pi
                                 : -0.5 * (dx * dx + dy * dy) + 0.5 * ((double)e.x * e.x + (double)e.y * e.y);
  }
  const double t = block_sum(lp);
  if (threadIdx.x == 0) P.cell_partials[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kSviThreads) vcb_svi_cell_backward_kernel(const vcb_svi_t P) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P.Nc) return;
  const float2 z = reinterpret_cast<const float2*>(P.phixy)[c];
  const float2 pr = reinterpret_cast<const float2*>(P.phixy_prior)[c];
  const float dphi = P.d_phi[P.cell_row ? P.cell_row[c] : c];
  const float r2 = z.x * z.x + z.y * z.y;
  // phi = atan2(y, x): dphi/dx = -y / r2, dphi/dy = x / r2; the guide's log q does not depend on locs (pathwise)
  const float ex = dphi * (-z.y / r2) - (z.x - pr.x);
  const float ey = dphi * (z.x / r2) - (z.y - pr.y);
  reinterpret_cast<float2*>(P.grad + P.o_phixy_locs)[c] = P.cond_phixy != nullptr ? make_float2(0.f, 0.f) : make_float2(-ex, -ey);
}

// ---- genes ---------------------------------------------------------------------------------------------------------
// model: 0 = phase (mean field), 1 = velocity mean field, 2 = velocity LRMN
__global__ void __launch_bounds__(kSviThreads) vcb_svi_gene_sample_kernel(const vcb_svi_t P, int n_cell_blocks) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int K = 2 * P.H + 1, Kw = 2 * P.Hw + 1;
  const long long Ng = P.Ng;
  double lp = 0.0;
  if (g < Ng) {
    for (int k = 0; k < K; ++k) {
      const long long i = g * K + k;
      const float s = expf(P.param[P.o_nu_scales + i]), e = P.eps_nu[i];
      const float v = P.cond_nu != nullptr ? P.cond_nu[i] : fmaf(s, e, P.param[P.o_nu_locs + i]);
      P.nu[i] = v;
      lp += normal_lp(v, P.mu_nu[i], P.sd_nu[i]);
      if (P.cond_nu == nullptr) lp -= -0.5 * (double)e * e - log((double)s) - kHalfLog2Pi;
    }
    if (P.o_dnu_locs >= 0)
      for (int b = 0; b < P.Nb; ++b) {
        const long long i = (long long)b * Ng + g;
        const float v = P.cond_dnu != nullptr ? P.cond_dnu[i] : P.param[P.o_dnu_locs + i];  // Delta guide
        P.dnu[i] = v;
        lp += normal_lp(v, 0.f, P.sd_dnu);
      }
    {
      const float x = P.cond_shape_inv != nullptr ? P.cond_shape_inv[g]
                                                  : expf(P.param[P.o_shape_inv_locs + g]);  // Delta guide, positive parameter
      P.shape_inv[g] = x;
      const double a = P.gamma_alpha, b = P.gamma_beta;
      lp += a * log(b) + (a - 1.0) * log((double)x) - b * (double)x - lgamma(a);
    }
    if (P.model == 1) {
      const float sg = expf(P.param[P.o_loggamma_scales + g]), eg = P.eps_loggamma[g];
      const float lg = fmaf(sg, eg, P.param[P.o_loggamma_locs + g]);
      const float sb = expf(P.param[P.o_logbeta_scales + g]), eb = P.eps_logbeta[g];
      const float lb = fmaf(sb, eb, P.param[P.o_logbeta_locs + g]);
      P.loggamma[g] = lg;
      P.gamma[g] = expf(lg);
      P.logbeta[g] = lb;
      lp += normal_lp(lg, P.mu_loggamma[g], P.sd_loggamma[g]) - (-0.5 * (double)eg * eg - log((double)sg) - kHalfLog2Pi);
      lp += normal_lp(lb, P.mu_logbeta[g], P.sd_logbeta[g]) - (-0.5 * (double)eb * eb - log((double)sb) - kHalfLog2Pi);
    } else if (P.model == 2) {
      // joint = loc + W eps_W + sqrt(D) eps_D (LowRankMultivariateNormal.rsample); log gamma = joint[g] through a Delta site
      float n = 0.f, w2 = 0.f;
      for (int r = 0; r < P.rank; ++r) {
        const float w = expf(P.param[P.o_cov_factor + g * P.rank + r]);
        n = fmaf(w, P.eps_W[r], n);
        w2 = fmaf(w, w, w2);
      }
      const float D = expf(P.param[P.o_cov_diag + g]);
      n = fmaf(sqrtf(D), P.eps_D[g], n);
      const float lg = P.param[P.o_loc + g] + n;
      const float gsd = sqrtf(w2 + D);
      const float rr = P.param[P.o_rho_real_loc + g];  // Delta guide
      const float rho = 1.998f / (1.f + expf(-rr / P.rho_scale)) - 0.999f;
      const float sb = expf(P.param[P.o_logbeta_scales + g]), eb = P.eps_logbeta[g];
      const float t = n / gsd;
      const float cm = P.param[P.o_logbeta_locs + g] + rho * sb * t;
      const float cs = sb * sqrtf(1.f - rho * rho);
      const float lb = fmaf(cs, eb, cm);
      P.loggamma[g] = lg;
      P.gamma[g] = expf(lg);
      P.logbeta[g] = lb;
      lp += normal_lp(lg, P.mu_loggamma[g], P.sd_loggamma[g]);
      lp += normal_lp(rr, P.rho_mean, P.rho_std);
      lp += normal_lp(lb, P.mu_logbeta[g], P.sd_logbeta[g]) - (-0.5 * (double)eb * eb - log((double)cs) - kHalfLog2Pi);
    }
  }
  // nu_omega (Nx x Kw numbers) and the sum of the cell kernel's block partials: block 0
  if (blockIdx.x == 0) {
    if (P.model != 0) {
      const int nw = P.Nx * Kw;
      for (int j = threadIdx.x; j < nw; j += blockDim.x) {
        float v;
        if (P.model == 1) {
          const float s = expf(P.param[P.o_nuw_scales + j]), e = P.eps_nuw[j];
          v = fmaf(s, e, P.param[P.o_nuw_locs + j]);
          lp += -(-0.5 * (double)e * e - log((double)s) - kHalfLog2Pi);
        } else {
          const long long i = Ng + j;  // the tail of the joint draw, through a Delta site
          float n = 0.f;
          for (int r = 0; r < P.rank; ++r) n = fmaf(expf(P.param[P.o_cov_factor + i * P.rank + r]), P.eps_W[r], n);
          n = fmaf(sqrtf(expf(P.param[P.o_cov_diag + i])), P.eps_D[i], n);
          v = P.param[P.o_loc + i] + n;
        }
        P.nu_omega[j] = v;
        lp += normal_lp(v, P.mu_nuw[j], P.sd_nuw[j]);
      }
    }
  }
  const double t = block_sum(lp);
  if (threadIdx.x == 0) P.gene_partials[blockIdx.x] = t;
  if (blockIdx.x == 0) {
    // fixed-order sum of the cell partials (thread t takes blocks t, t + T, ...; then the T terms in order)
    double c = 0.0;
    for (int b = threadIdx.x; b < n_cell_blocks; b += blockDim.x) c += P.cell_partials[b];
    const double tc = block_sum(c);
    if (threadIdx.x == 0) *P.cell_lp = (float)tc;  // a slot of the buffer the all-reduce moves: global under cell sharding
  }
}

__global__ void __launch_bounds__(kSviThreads) vcb_svi_gene_backward_kernel(const vcb_svi_t P) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int K = 2 * P.H + 1, Kw = 2 * P.Hw + 1;
  const long long Ng = P.Ng;
  double lp = 0.0;  // likelihood log-prob of this gene
  if (g < Ng) {
    lp = (double)P.lp_S[g] + (P.lp_U ? (double)P.lp_U[g] : 0.0);
    for (int k = 0; k < K; ++k) {
      const long long i = g * K + k;
      const float s = expf(P.param[P.o_nu_scales + i]), e = P.eps_nu[i];
      const float sd = P.sd_nu[i];
      const float A = P.d_nu[i] - (P.nu[i] - P.mu_nu[i]) / (sd * sd);  // d ELBO / d nu
      const bool c = P.cond_nu != nullptr;                             // conditioned: the guide's parameters see nothing
      P.grad[P.o_nu_locs + i] = c ? 0.f : -A;
      P.grad[P.o_nu_scales + i] = c ? 0.f : -(A * e * s + 1.f);  // + log s from the entropy; d/du = d/ds * s
    }
    if (P.o_dnu_locs >= 0)
      for (int b = 0; b < P.Nb; ++b) {
        const long long i = (long long)b * Ng + g;
        P.grad[P.o_dnu_locs + i] = P.cond_dnu != nullptr ? 0.f : -(P.d_dnu[i] - P.dnu[i] / (P.sd_dnu * P.sd_dnu));
      }
    {
      const float x = P.shape_inv[g];
      P.grad[P.o_shape_inv_locs + g] =
          P.cond_shape_inv != nullptr ? 0.f : -(P.d_shape_inv[g] + (P.gamma_alpha - 1.f) / x - P.gamma_beta) * x;
    }
    if (P.model == 1) {
      const float sg = expf(P.param[P.o_loggamma_scales + g]), eg = P.eps_loggamma[g];
      const float sb = expf(P.param[P.o_logbeta_scales + g]), eb = P.eps_logbeta[g];
      const float sdg = P.sd_loggamma[g], sdb = P.sd_logbeta[g];
      const float Ag = P.d_gamma[g] * P.gamma[g] - (P.loggamma[g] - P.mu_loggamma[g]) / (sdg * sdg);
      const float Ab = P.d_logbeta[g] - (P.logbeta[g] - P.mu_logbeta[g]) / (sdb * sdb);
      P.grad[P.o_loggamma_locs + g] = -Ag;
      P.grad[P.o_loggamma_scales + g] = -(Ag * eg * sg + 1.f);
      P.grad[P.o_logbeta_locs + g] = -Ab;
      P.grad[P.o_logbeta_scales + g] = -(Ab * eb * sb + 1.f);
    } else if (P.model == 2) {
      // recompute the guide's intermediates (cheaper than saving them)
      float n = 0.f, w2 = 0.f;
      for (int r = 0; r < P.rank; ++r) {
        const float w = expf(P.param[P.o_cov_factor + g * P.rank + r]);
        n = fmaf(w, P.eps_W[r], n);
        w2 = fmaf(w, w, w2);
      }
      const float D = expf(P.param[P.o_cov_diag + g]), sqD = sqrtf(D);
      n = fmaf(sqD, P.eps_D[g], n);
      const float gsd = sqrtf(w2 + D);
      const float rr = P.param[P.o_rho_real_loc + g];
      const float sig = 1.f / (1.f + expf(-rr / P.rho_scale));
      const float rho = 1.998f * sig - 0.999f;
      const float om = 1.f - rho * rho, sq = sqrtf(om);
      const float sb = expf(P.param[P.o_logbeta_scales + g]), eb = P.eps_logbeta[g];
      const float t = n / gsd;
      const float sdg = P.sd_loggamma[g], sdb = P.sd_logbeta[g];
      const float B = P.d_logbeta[g] - (P.logbeta[g] - P.mu_logbeta[g]) / (sdb * sdb);                // d ELBO / d log beta
      const float Ag = P.d_gamma[g] * P.gamma[g] - (P.loggamma[g] - P.mu_loggamma[g]) / (sdg * sdg);  // ... / d log gamma, direct
      // log beta = bl + rho sb t + sb sqrt(1 - rho^2) eb;  entropy term + log(sb sqrt(1 - rho^2))
      P.grad[P.o_logbeta_locs + g] = -B;
      P.grad[P.o_logbeta_scales + g] = -((B * (rho * t + sq * eb) + 1.f / sb) * sb);
      const float dE_drho = B * (sb * t - sb * rho / sq * eb) - rho / om;
      const float drho_drr = 1.998f * sig * (1.f - sig) / P.rho_scale;
      P.grad[P.o_rho_real_loc + g] = -(dE_drho * drho_drr - (rr - P.rho_mean) / (P.rho_std * P.rho_std));
      const float Bt = B * rho * sb;           // d ELBO / d t
      const float Cn = Ag + Bt / gsd;          // ... / d n  (log gamma = loc + n; t = n / gsd)
      const float dgsd = -Bt * t / gsd;        // ... / d gsd
      P.grad[P.o_loc + g] = -Ag;
      for (int r = 0; r < P.rank; ++r) {
        const long long i = g * P.rank + r;
        const float w = expf(P.param[P.o_cov_factor + i]);
        P.grad[P.o_cov_factor + i] = -((Cn * P.eps_W[r] + dgsd * w / gsd) * w);
      }
      P.grad[P.o_cov_diag + g] = -((Cn * P.eps_D[g] / (2.f * sqD) + dgsd / (2.f * gsd)) * D);
    }
  }
  if (blockIdx.x == 0 && P.model != 0) {
    const int nw = P.Nx * Kw;
    for (int j = threadIdx.x; j < nw; j += blockDim.x) {
      const float sd = P.sd_nuw[j];
      const float A = P.d_nu_omega[j] - (P.nu_omega[j] - P.mu_nuw[j]) / (sd * sd);
      if (P.model == 1) {
        const float s = expf(P.param[P.o_nuw_scales + j]), e = P.eps_nuw[j];
        P.grad[P.o_nuw_locs + j] = -A;
        P.grad[P.o_nuw_scales + j] = -(A * e * s + 1.f);
      } else {
        const long long i = Ng + j;
        P.grad[P.o_loc + i] = -A;
        for (int r = 0; r < P.rank; ++r) {
          const float w = expf(P.param[P.o_cov_factor + i * P.rank + r]);
          P.grad[P.o_cov_factor + i * P.rank + r] = -(A * P.eps_W[r] * w);
        }
        const float D = expf(P.param[P.o_cov_diag + i]);
        P.grad[P.o_cov_diag + i] = -(A * P.eps_D[i] / (2.f * sqrtf(D)) * D);
      }
    }
  }
  const double t = block_sum(lp);
  if (threadIdx.x == 0) P.lik_partials[blockIdx.x] = t;
}

__global__ void __launch_bounds__(kSviThreads) vcb_svi_finalize_kernel(const vcb_svi_t P, int n_gene_blocks) {
  double a = 0.0;
  for (int b = threadIdx.x; b < n_gene_blocks; b += blockDim.x) a += P.gene_partials[b] + P.lik_partials[b];
  const double t = block_sum(a);
  if (threadIdx.x == 0) *P.loss = (float)(-(t + (double)*P.cell_lp));
}

static int svi_validate(const vcb_svi_t* p) {
  if (p == nullptr) return VCB_ERR_NULL;
  if (p->Nc < 0 || p->Ng <= 0 || p->Nc > (1LL << 40) || p->Ng > (1LL << 24)) return VCB_ERR_SIZE;
  if (p->H < 0 || p->H > VCB_MAX_HARMONICS || p->Hw < 0 || p->Hw > VCB_MAX_HARMONICS) return VCB_ERR_HARMONICS;
  if (p->model < 0 || p->model > 2 || p->Nb < 0) return VCB_ERR_SIZE;
  if (!p->param || !p->grad || !p->eps_nu || !p->eps_phixy || !p->mu_nu || !p->sd_nu || !p->phixy_prior) return VCB_ERR_NULL;
  if (!p->nu || !p->shape_inv || !p->phi || !p->phixy || !p->cell_partials || !p->gene_partials || !p->lik_partials ||
      !p->cell_lp || !p->loss)
    return VCB_ERR_NULL;
  if (p->o_nu_locs < 0 || p->o_nu_scales < 0 || p->o_phixy_locs < 0 || p->o_shape_inv_locs < 0) return VCB_ERR_SIZE;
  if ((p->o_phixy_locs & 1) != 0) return VCB_ERR_ALIGN;  // float2 access
  if (p->o_dnu_locs >= 0 && (p->Nb < 1 || !p->dnu)) return VCB_ERR_NULL;
  if (p->model != 0) {
    if (p->Nx < 1) return VCB_ERR_SIZE;
    if (!p->eps_logbeta || !p->mu_loggamma || !p->sd_loggamma || !p->mu_logbeta || !p->sd_logbeta || !p->mu_nuw || !p->sd_nuw ||
        !p->loggamma || !p->gamma || !p->logbeta || !p->nu_omega)
      return VCB_ERR_NULL;
    if (p->o_logbeta_locs < 0 || p->o_logbeta_scales < 0) return VCB_ERR_SIZE;
    if (p->model == 1) {
      if (!p->eps_loggamma || !p->eps_nuw) return VCB_ERR_NULL;
      if (p->o_loggamma_locs < 0 || p->o_loggamma_scales < 0 || p->o_nuw_locs < 0 || p->o_nuw_scales < 0) return VCB_ERR_SIZE;
    } else {
      if (!p->eps_W || !p->eps_D) return VCB_ERR_NULL;
      if (p->o_loc < 0 || p->o_cov_factor < 0 || p->o_cov_diag < 0 || p->o_rho_real_loc < 0 || p->rank < 1) return VCB_ERR_SIZE;
    }
  }
  return VCB_OK;
}

}  // namespace vcb

extern "C" {

int vcb_svi_partials(int64_t Nc, int64_t Ng, int64_t* n_cell_blocks, int64_t* n_gene_blocks) {
  if (Nc < 0 || Ng <= 0 || !n_cell_blocks || !n_gene_blocks) return VCB_ERR_SIZE;
  *n_cell_blocks = (Nc + vcb::kSviThreads - 1) / vcb::kSviThreads;
  if (*n_cell_blocks < 1) *n_cell_blocks = 1;
  *n_gene_blocks = (Ng + vcb::kSviThreads - 1) / vcb::kSviThreads;
  return VCB_OK;
}

int vcb_svi_sample(const vcb_svi_t* p, void* stream) {
  vcb::DeviceGuard guard(p ? p->param : nullptr);
  int rc = vcb::svi_validate(p);
  if (rc != VCB_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int ncb = (int)((p->Nc + vcb::kSviThreads - 1) / vcb::kSviThreads);
  const int ngb = (int)((p->Ng + vcb::kSviThreads - 1) / vcb::kSviThreads);
  if (ncb > 0) {
    vcb::vcb_svi_cell_sample_kernel<<<ncb, vcb::kSviThreads, 0, st>>>(*p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  vcb::vcb_svi_gene_sample_kernel<<<ngb, vcb::kSviThreads, 0, st>>>(*p, ncb);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? VCB_OK : (int)e;
}

int vcb_svi_backward(const vcb_svi_t* p, void* stream) {
  vcb::DeviceGuard guard(p ? p->param : nullptr);
  int rc = vcb::svi_validate(p);
  if (rc != VCB_OK) return rc;
  if (!p->lp_S || !p->d_nu || !p->d_shape_inv || !p->d_phi) return VCB_ERR_NULL;
  if (p->o_dnu_locs >= 0 && !p->d_dnu) return VCB_ERR_NULL;
  if (p->model != 0 && (!p->lp_U || !p->d_logbeta || !p->d_gamma || !p->d_nu_omega)) return VCB_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  const int ncb = (int)((p->Nc + vcb::kSviThreads - 1) / vcb::kSviThreads);
  const int ngb = (int)((p->Ng + vcb::kSviThreads - 1) / vcb::kSviThreads);
  cudaError_t e;
  if (ncb > 0) {
    vcb::vcb_svi_cell_backward_kernel<<<ncb, vcb::kSviThreads, 0, st>>>(*p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  vcb::vcb_svi_gene_backward_kernel<<<ngb, vcb::kSviThreads, 0, st>>>(*p);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  vcb::vcb_svi_finalize_kernel<<<1, vcb::kSviThreads, 0, st>>>(*p, ngb);
  e = cudaGetLastError();
  return e == cudaSuccess ? VCB_OK : (int)e;
}

}  // extern "C"
