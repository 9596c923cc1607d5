/*
 * vcb.h -- C ABI of libvcb.so: the B200 (sm_100a) fused ELBO + gradient path of VeloCycle.
 *
 * What this boundary replaces in the reference (lamanno-epfl/velocycle, working-tree line numbers):
 *
 *   vcb_phase_fwd_bwd     <->  velocycle/phase_inference_model.py:368-393
 *                              (pack_direction -> torch_fourier_basis -> ElogS einsums ->
 *                               pyro.sample("S", GammaPoisson(...), obs=mp.S)) plus the autograd
 *                              backward of that block driven by pyro.infer.Trace_ELBO at :162/:169.
 *   vcb_velocity_fwd_bwd  <->  velocycle/velocity_inference_model.py:344-386 (and the identical
 *                              LRMN block :443-469): zeta, zeta', zeta_omega, ElogS, omega, ElogU and
 *                              the two GammaPoisson sites "S" and "U", plus their backward (:111/:120).
 *   vcb_count_histogram   <->  no reference counterpart; one-off data statistic that lets the
 *                              parameter-only lgamma/digamma terms of GammaPoisson.log_prob
 *                              (pyro/distributions/conjugate.py) leave the per-element loop.
 *   vcb_expand_counts     <->  the device side of preprocessing.py:142-143 / 193-194 (248-249 / 308-309):
 *                              `torch.tensor(S).to(device)` ... `.T.float()`.  The reference ships the dense
 *                              count matrix to the device as 8-byte integers and widens it there; this entry
 *                              point widens a compact host staging format (u8 with an overflow list, u16, i32)
 *                              into the float32 cell-major layout the kernels stream, so a count matrix crosses
 *                              PCIe at 1-2 bytes per entry instead of 4-8.
 *   vcb_csr_to_counts     <->  preprocessing.py:138-143 / 243-249: the sparse anndata layers are scattered into the
 *                              device layout directly (no dense int64 host copy: 8 B x Nc x Ng of host memory and PCIe).
 *   vcb_clipped_adam      <->  pyro.optim.ClippedAdam.step (pyro/optim/clipped_adam.py), configured in
 *                              tutorials/Tutorial_Capolupo_HumanFibroblasts_OneSample.ipynb cell 27.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer owned by the caller (torch tensors' data_ptr()); the library
 *     never allocates or frees device memory and keeps no global state.  Scratch space is the
 *     caller-provided workspace (size from vcb_workspace_bytes).
 *   - All work is enqueued on the cudaStream_t passed as `void* stream`; nothing synchronises the
 *     device, so every entry point is CUDA-graph capturable.
 *   - Return value: 0 = OK; negative = argument error detected on the host (nothing launched);
 *     positive = cudaError_t of a failed launch.  vcb_strerror() names both.
 *   - Counts S/U: float32, cell-major ("[Nc][ld]", genes contiguous) -- the physical layout of the
 *     reference's mp.S = S.T.float() (preprocessing.py:193-194, 308-309).  ld is the row pitch in
 *     floats: a multiple of 4 with 16-byte aligned base pointers; columns Ng..ld-1 must be zero.
 *   - Outputs are overwritten, never accumulated.
 */
#ifndef VCB_H
#define VCB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VCB_VERSION 210 /* 0.2.1 */

#define VCB_MAX_HARMONICS 5 /* gene / angular-speed harmonics compiled in: H in 0..5 */

/* flags */
#define VCB_FLAG_GRAD 1u          /* also emit every gradient (otherwise log-prob sums only) */
#define VCB_FLAG_LGAMMA_INLINE 2u /* evaluate lgamma/digamma terms per element instead of via the histogram */
#define VCB_FLAG_TCGEN05 4u       /* stream with the tcgen05 kernel (vcb_umma.cuh) where it applies: velocity model, VCB_FLAG_GRAD,
                                     no VCB_FLAG_LGAMMA_INLINE, H <= 3, Nb <= 1; ignored otherwise.  Same results to fp32 rounding.
                                     (The library reads no environment variables.) */

#define VCB_FLAG_LEGACY_STREAM 8u  /* stream with the round-1 kernel (vcb_stream.cuh) even where the round-2 kernel applies
                                     (kept for A/B measurements and as the H > 3 / inline-lgamma path) */

/* error codes (negative) */
#define VCB_OK 0
#define VCB_ERR_NULL -1        /* a required pointer is NULL */
#define VCB_ERR_SIZE -2        /* a size is out of range */
#define VCB_ERR_ALIGN -3       /* ld not a multiple of 4 or S/U/workspace not 16-byte aligned */
#define VCB_ERR_HARMONICS -4   /* H or Hw above VCB_MAX_HARMONICS */
#define VCB_ERR_WORKSPACE -5   /* workspace too small */
#define VCB_ERR_SPECTRUM -6    /* histogram required (no VCB_FLAG_LGAMMA_INLINE) but not provided */
#define VCB_ERR_DEVICE -7      /* not an sm_100 device / no device */

/* Sparse per-gene histogram of one count matrix ("count spectrum"), built once per dataset:
 * for gene g the distinct values k > 0 that occur are val[off[g] .. off[g+1]) with multiplicities mult[].
 * lgk1[g] = sum_c lgamma(k_gc + 1) (parameter-free part of the log-pmf). */
typedef struct vcb_spectrum {
  const int32_t* off;  /* [Ng+1] */
  const float* val;    /* [nnz]  */
  const float* mult;   /* [nnz]  */
  const double* lgk1;  /* [Ng]   */
} vcb_spectrum_t;

typedef struct vcb_problem {
  int64_t Nc; /* cells held by this rank */
  int64_t Ng; /* genes */
  int64_t ld; /* row pitch of S and U in floats */
  int32_t H;  /* gene harmonics, K = 2H+1                      (mp.num_harmonics_S / kwargs zeta)      */
  int32_t Hw; /* angular-speed harmonics, Kw = 2Hw+1             (kwargs zeta_omega; velocity only)     */
  int32_t Nb; /* rows of dnu; 0 = no batch offset term          (mp.Nb, mp.with_delta_nu)              */
  int32_t Nx; /* rows of nu_omega                               (mp.Nx; velocity only)                 */
  uint32_t flags;
  uint32_t reserved;

  /* observed counts */
  const float* S; /* [Nc][ld] */
  const float* U; /* [Nc][ld], velocity only */

  /* per-cell inputs */
  const float* phi;        /* [Nc] phase angle (atan2 of the phixy sample stays in torch) */
  const float* cf;         /* [Nc] mp.count_factor                                          */
  const int32_t* batch_id; /* [Nc] argmax of the one-hot mp.Db column, NULL = all 0         */
  const int32_t* cond_id;  /* [Nc] argmax of the one-hot mp.D column,  NULL = all 0         */

  /* per-gene inputs */
  const float* nu;        /* [Ng][K]  Fourier coefficients, order [1, sin, cos, sin2, cos2, ...] */
  const float* dnu;       /* [Nb][Ng] batch offsets (Delta-nu), NULL when Nb == 0               */
  const float* shape_inv; /* [Ng]     NB dispersion, r = 1/shape_inv                            */
  const float* logbeta;   /* [Ng]     velocity only                                             */
  const float* gamma;     /* [Ng]     exp(log gamma), velocity only                             */
  const float* nu_omega;  /* [Nx][Kw] velocity only                                             */

  /* count spectra (ignored under VCB_FLAG_LGAMMA_INLINE) */
  vcb_spectrum_t spec_S;
  vcb_spectrum_t spec_U;

  /* outputs: per-gene log-prob sums over this rank's cells */
  float* lp_S; /* [Ng] */
  float* lp_U; /* [Ng] velocity only */

  /* outputs: gradients of sum(lp_S)+sum(lp_U); only written under VCB_FLAG_GRAD; any may be NULL */
  float* d_nu;        /* [Ng][K]  */
  float* d_dnu;       /* [Nb][Ng] */
  float* d_shape_inv; /* [Ng]     */
  float* d_logbeta;   /* [Ng]     */
  float* d_gamma;     /* [Ng]     */
  float* d_nu_omega;  /* [Nx][Kw] */
  float* d_phi;       /* [Nc] total derivative (includes the path through omega(phi)) */
  float* d_cf;        /* [Nc]     */
  float* d_omega;     /* [Nc] partial w.r.t. the per-cell angular speed (informational) */

  /* optional caller-owned cudaEvent_t handles recorded on `stream` right before / after the streaming kernel
   * (the dominant launch), so that a benchmark can time it without a profiler; NULL = not recorded */
  void* ev_stream_begin;
  void* ev_stream_end;
} vcb_problem_t;

int vcb_version(void);
const char* vcb_strerror(int code);

/* Bytes of device scratch the two fwd_bwd entry points need for this problem (16-byte aligned base). */
size_t vcb_workspace_bytes(const vcb_problem_t* p);

/* Phase (manifold-learning) model: spliced counts only. */
int vcb_phase_fwd_bwd(const vcb_problem_t* p, void* workspace, size_t workspace_bytes, void* stream);

/* Velocity model: spliced + unspliced counts in the same pass. */
int vcb_velocity_fwd_bwd(const vcb_problem_t* p, void* workspace, size_t workspace_bytes, void* stream);

/* Dense per-gene histogram of one count matrix: hist[g*B + min(k, B-1)] += 1 (uint32, caller zeroes it).
 * status[0] is set non-zero if a value is negative, non-integer or >= B-1 (caller zeroes it too). */
int vcb_count_histogram(const float* M, int64_t Nc, int64_t Ng, int64_t ld, int32_t B, uint32_t* hist,
                        int32_t* status, void* stream);

/* Staging formats of vcb_expand_counts (src_dtype). */
#define VCB_COUNTS_U8 1  /* one byte per entry; 255 = "see the overflow list" */
#define VCB_COUNTS_U16 2 /* two bytes per entry */
#define VCB_COUNTS_I32 4 /* four bytes per entry (the reference's anndata layers after astype(int)) */

/* Widen `n` staged count entries (a whole number of [ld]-pitched rows, device memory, 16-byte aligned, n % 16 == 0
 * or the tail is handled element-wise) to float32: dst[i] = (float)src[i].  Under VCB_COUNTS_U8 the `n_over` entries
 * whose count is >= 255 are listed in (over_idx[j], over_val[j]) -- flat indices into dst, any order, no duplicates --
 * and are written after the bulk pass on the same stream.  Exact for every count < 2^24. */
int vcb_expand_counts(const void* src, int32_t src_dtype, int64_t n, float* dst, const int64_t* over_idx,
                      const float* over_val, int64_t n_over, void* stream);

/* Sub-byte staging (vcb_expand_counts_packed): `bits` = 2 or 4 bits per entry, 32/bits entries per little-endian 32-bit
 * word (entry i of a word in bits [i*bits, (i+1)*bits)).  The all-ones code is an escape: the entry's value is the next byte
 * of `side` in entry order; a side byte of 255 is itself an escape into the (over_idx, over_val) list, exactly as in
 * vcb_expand_counts.  `block_off[j]` = number of escapes before word 256*j (one entry per VCB_PACKED_BLOCK_WORDS words), so
 * that blocks decode independently.  `codes` holds a whole number of blocks (zero padded); n = Nc*ld entries are written.
 * Typical scRNA-seq counts cost 0.3-0.55 bytes per entry instead of 1 (u8) or 8 (the reference's int64 upload). */
#define VCB_PACKED_BLOCK_WORDS 256
int vcb_expand_counts_packed(const uint32_t* codes, int32_t bits, const uint8_t* side, const int64_t* block_off, int64_t n,
                             float* dst, const int64_t* over_idx, const float* over_val, int64_t n_over, void* stream);

/* Two-level variant of the 2-bit format: code 3 -> the next NIBBLE of `nibbles` (8 per little-endian 32-bit word, entry
 * order) holding value-3 in 0..14; nibble 15 -> the next byte of `side` (value, 255 = overflow list).  block_off1[j] / block_off2[j]
 * = nibbles / side bytes consumed before word 256*j.  0.30-0.45 bytes per entry for scRNA-seq counts. */
int vcb_expand_counts_twolevel(const uint32_t* codes, const uint32_t* nibbles, const uint8_t* side, const int64_t* block_off1,
                              const int64_t* block_off2, int64_t n, float* dst, const int64_t* over_idx, const float* over_val,
                              int64_t n_over, void* stream);

/* Value types of vcb_csr_to_counts (data_dtype). */
#define VCB_CSR_F32 0
#define VCB_CSR_I32 1
#define VCB_CSR_F64 2
#define VCB_CSR_I64 3

/* Build the float32 cell-major count matrix dst[Nc][ld] (ld % 4 == 0, ld >= Ng, padding columns zero) on the device from a
 * CSR matrix of shape (Nc cells, Ng genes) -- the layout of the anndata layers `spliced` / `unspliced` the reference
 * densifies on the host (preprocessing.py:138-143, 243-249: `.A` / np.array(...) then `torch.tensor(S).to(device)`).
 * indptr: Nc+1 int64 offsets; indices: int32 gene ids (any order inside a row); data: nnz values of `data_dtype`.
 * Duplicate (row, gene) entries are summed like scipy's toarray(); entries with a gene id outside [0, Ng) or a negative /
 * non-integer / >= 2^24 value are skipped and make *status non-zero (status may be NULL).  dst is overwritten. */
int vcb_csr_to_counts(const int64_t* indptr, const int32_t* indices, const void* data, int32_t data_dtype, int64_t Nc,
                      int64_t Ng, int64_t ld, float* dst, int32_t* status, void* stream);

/* Multi-tensor ClippedAdam over one flat fp32 buffer of n elements:
 *   lr_t = lr0 * lrd^step (step counts from 1, read from device memory so that graphs replay),
 *   g = clamp(g, -clip, clip); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
 *   p -= lr_t * sqrt(1-b2^step)/(1-b1^step) * m / (sqrt(v) + eps).
 * `step_dev` points at an int64 step counter that this call increments first. */
int vcb_clipped_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                     int64_t* step_dev, float lr0, float lrd, float beta1, float beta2, float eps,
                     float clip, void* stream);


/* ---- the non-likelihood part of one Trace_ELBO step, fused (csrc/vcb_svi.cu) ----------------------------------------------
 * Replaces, for the package's own (model, guide) pairs, what pyro.infer.SVI.step does around the likelihood: the guide's
 * reparameterised draws (phase_inference_guide.py:47-56, velocity_inference_guide.py:45-63 and :90-141), pack_direction
 * (utils.py:488-506), the log_prob of every prior and guide site (phase_inference_model.py:361-366,391,
 * velocity_inference_model.py:323-353,383 and :408-440) and their autograd backward.  Call order of a step:
 *   vcb_svi_sample -> vcb_phase_fwd_bwd | vcb_velocity_fwd_bwd (on the sampled values) [-> all-reduce] -> vcb_svi_backward
 *   -> vcb_clipped_adam.
 * `param` / `grad`: ONE flat fp32 buffer each holding every guide parameter in unconstrained form (positive ones as logs:
 * transform_to(positive) = exp); the o_* fields are offsets in floats, -1 = the model has no such parameter.  Gradients are
 * those of loss = -ELBO with respect to the unconstrained values and are OVERWRITTEN.  The eps_* arrays are standard-normal
 * draws supplied by the caller in the guide's draw order (so a step consumes the RNG stream like the reference does).
 * model: 0 = phase, 1 = velocity with the mean-field guide, 2 = velocity with the LRMN guide.                               */
typedef struct vcb_svi_t {
  int64_t Nc, Ng;
  int32_t H, Hw, Nb, Nx, rank, model;
  float* param;
  float* grad;
  int64_t o_nu_locs, o_nu_scales;            /* [Ng][K] */
  int64_t o_dnu_locs;                        /* [Nb][Ng] */
  int64_t o_phixy_locs;                      /* [Nc][2], even offset */
  int64_t o_shape_inv_locs;                  /* [Ng] */
  int64_t o_logbeta_locs, o_logbeta_scales;  /* [Ng] */
  int64_t o_loggamma_locs, o_loggamma_scales; /* [Ng]            (model 1) */
  int64_t o_nuw_locs, o_nuw_scales;          /* [Nx][Kw]        (model 1) */
  int64_t o_loc, o_cov_factor, o_cov_diag;   /* [Ng + Nx Kw], [Ng + Nx Kw][rank], [Ng + Nx Kw]   (model 2) */
  int64_t o_rho_real_loc;                    /* [Ng]            (model 2) */
  const float *eps_nu, *eps_loggamma, *eps_logbeta, *eps_nuw, *eps_phixy, *eps_W, *eps_D;
  const float *mu_nu, *sd_nu;                /* priors, expanded: [Ng][K] */
  const float *mu_loggamma, *sd_loggamma, *mu_logbeta, *sd_logbeta; /* [Ng] */
  const float *mu_nuw, *sd_nuw;              /* [Nx][Kw] */
  const float* phixy_prior;                  /* [Nc][2] */
  const int64_t* cell_row;                   /* [Nc] or NULL: row of cell c in the (batch-sorted) count matrices; phi and d_phi
                                                are indexed by row, everything else here by cell */
  /* conditioned sites (poutine.condition on the model + poutine.block on the guide, the tutorial's velocity stage,
     phase_inference_model.py:110-115): the site takes the given value, keeps its prior log-prob, loses its guide term and
     its parameters get a zero gradient; NULL = sampled from the guide.  (The draws are made all the same: RNG parity.) */
  const float *cond_nu, *cond_dnu, *cond_shape_inv, *cond_phixy; /* [Ng][K], [Nb][Ng], [Ng], [Nc][2] */
  float sd_dnu, gamma_alpha, gamma_beta, rho_mean, rho_std, rho_scale;
  /* sampled values: written by vcb_svi_sample, inputs of the likelihood call and of vcb_svi_backward */
  float *nu, *dnu, *shape_inv, *loggamma, *gamma, *logbeta, *nu_omega, *phixy, *phi;
  /* outputs of the likelihood call (after the all-reduce under cell sharding): inputs of vcb_svi_backward */
  const float *lp_S, *lp_U, *d_nu, *d_dnu, *d_shape_inv, *d_logbeta, *d_gamma, *d_nu_omega, *d_phi;
  /* scratch: fp64 block partials, sizes from vcb_svi_partials() */
  double *cell_partials, *gene_partials, *lik_partials;
  float* cell_lp; /* sum over this rank's cells of log p(phixy) - log q(phixy); lives in the all-reduced buffer under sharding */
  float* loss;    /* the step's ELBO loss, written by vcb_svi_backward */
} vcb_svi_t;

int vcb_svi_partials(int64_t Nc, int64_t Ng, int64_t* n_cell_blocks, int64_t* n_gene_blocks);
int vcb_svi_sample(const vcb_svi_t* p, void* stream);
int vcb_svi_backward(const vcb_svi_t* p, void* stream);


/* ---- one-shot SUM all-reduce over NVLink peer memory (csrc/vcb_comm.cu) -----------------------------------------------------
 * The step's single exchange under cell sharding (north_star: "gene-level parameter gradients are summed with a single
 * allreduce over NVLink per step").  slots[r] / flags[r] point at rank r's receive buffer (2 x world x slot_floats floats)
 * and flag words (2 x world uint32, zero-initialised), mapped into this process with CUDA IPC; `epoch` is a uint32 in this
 * rank's device memory, zero-initialised, advanced by every call.  All ranks must call with the same n (a multiple of 4,
 * <= slot_floats) in the same order.  The sum is taken in rank order: reproducible and identical on every rank.            */
#define VCB_MAX_RANKS 16
typedef struct vcb_comm_t {
  int32_t rank, world;
  int64_t slot_floats;
  float* slots[VCB_MAX_RANKS];
  uint32_t* flags[VCB_MAX_RANKS];
  uint32_t* epoch;
} vcb_comm_t;

int vcb_allreduce_sum(const vcb_comm_t* c, float* data, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VCB_H */
