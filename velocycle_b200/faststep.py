"""One Trace_ELBO step of the package's own (model, guide) pairs with everything around the likelihood fused.

``pyro.infer.SVI.step`` (``phase_inference_model.py:162-169``, ``velocity_inference_model.py:111-120``) traces the guide
and the model through effect handlers: per sample site a reparameterised draw, a prior and a guide ``log_prob`` and
their autograd backward -- about 460 small launches per step for the velocity model.  When the model and guide are the
ones this package ships (the reference's, ``phase_inference_guide.py:10-56`` / ``velocity_inference_guide.py:9-141``)
and nothing is conditioned, the same step is

    torch normal_() draws in the guide's order  ->  vcb_svi_sample  ->  vcb_{phase,velocity}_fwd_bwd
    [-> all-reduce under cell sharding]  ->  vcb_svi_backward  ->  vcb_clipped_adam

i.e. a dozen launches (``csrc/vcb_svi.cu``).  The draws come from torch's generator with the shapes and in the order
of the traced guide, so a seed gives the same noise either way; the traced path stays the reference for the tests
(``tests/test_faststep_gpu.py``) and the fall-back for conditioned or user-supplied models.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from .fused import VCB_FLAG_GRAD, VcbProblem, _ptr
from .likelihood import packed_counts_for
from .sharding import PeerComm, allreduce_flat_

__all__ = ["FusedStep"]

_PARAMS = {
    0: ["ν_locs", "ν_scales", "ϕxy_locs", "shape_inv_locs"],
    1: ["ν_locs", "ν_scales", "ϕxy_locs", "shape_inv_locs", "logβg_locs", "logβg_scales", "logγg_locs", "logγg_scales",
        "νω_locs", "νω_scales"],
    2: ["ν_locs", "ν_scales", "ϕxy_locs", "shape_inv_locs", "logβg_locs", "logβg_scales", "loc", "cov_factor", "cov_diag",
        "rho_real_loc"],
}
_FIELD = {"ν_locs": "o_nu_locs", "ν_scales": "o_nu_scales", "Δν_locs": "o_dnu_locs", "ϕxy_locs": "o_phixy_locs",
          "shape_inv_locs": "o_shape_inv_locs", "logβg_locs": "o_logbeta_locs", "logβg_scales": "o_logbeta_scales",
          "logγg_locs": "o_loggamma_locs", "logγg_scales": "o_loggamma_scales", "νω_locs": "o_nuw_locs",
          "νω_scales": "o_nuw_scales", "loc": "o_loc", "cov_factor": "o_cov_factor", "cov_diag": "o_cov_diag",
          "rho_real_loc": "o_rho_real_loc"}


_CONDITIONABLE = {"ν": "cond_nu", "Δν": "cond_dnu", "shape_inv": "cond_shape_inv", "ϕxy": "cond_phixy"}


def model_code(model, guide, mp):
    """``(code, conditioned)`` when (model, guide) is one of the package's standard pairs -- code 0 / 1 / 2 as in
    include/vcb.h (vcb_svi_t.model) -- possibly wrapped the way the fit drivers wrap them for ``condition_on``
    (``poutine.condition(model, data)`` + ``poutine.block(guide, hide=the same sites)``, phase_inference_model.py:110-115) with
    sites among ν, Δν, shape_inv, ϕxy (the tutorial's velocity stage conditions on exactly these); otherwise None."""
    from . import phase_inference_guide as pg, phase_inference_model as pm
    from . import velocity_inference_guide as vg, velocity_inference_model as vm
    from .ppl import poutine as shim_poutine

    if getattr(mp, "noisemodel", None) != "NegativeBinomial":
        return None
    cond = {}
    if isinstance(model, shim_poutine.ConditionMessenger) and isinstance(guide, shim_poutine.BlockMessenger):
        if not guide.plain_hide or guide.hide != frozenset(model.data) or not set(model.data) <= set(_CONDITIONABLE):
            return None
        cond, model, guide = dict(model.data), model.fn, guide.fn
        if "Δν" in cond and not mp.with_delta_nu:
            return None
    if model is pm.phase_latent_variable_model and guide is pg.phase_latent_variable_guide:
        return 0, cond
    if model is vm.velocity_latent_variable_model and guide is vg.velocity_latent_variable_guide:
        return 1, cond
    if model is vm.velocity_latent_variable_model_LRMN and guide is vg.velocity_latent_variable_guide_LRMN:
        return 2, cond
    return None


class FusedStep:
    def __init__(self, gsvi, code: int, conditioned=None):
        mp = gsvi.mp
        self.g, self.mp, self.code = gsvi, mp, code
        self.lib = _lib.load()
        dev = gsvi.device
        self.dev = dev
        velocity = code != 0
        self.counts = packed_counts_for(mp, need_U=velocity)
        self.shard = getattr(self.counts, "shard", None) or getattr(mp, "shard", None)
        Nc, Ng = int(mp.Nc), int(mp.Ng)
        K = int(mp.μνg.shape[-1])
        Nb = int(mp.Nb) if mp.with_delta_nu else 0
        Nx = int(mp.Nx) if velocity else 0
        Kw = int(mp.Nhω) if velocity else 1
        f32 = dict(dtype=torch.float32, device=dev)
        names = list(_PARAMS[code]) + (["Δν_locs"] if mp.with_delta_nu else [])
        if set(names) != set(gsvi.param_slices):
            raise _lib.VcbError(f"unexpected parameter set {sorted(gsvi.param_slices)} for the fused step")
        p = _lib.VcbSvi()
        p.Nc, p.Ng, p.H, p.Hw, p.Nb, p.Nx = Nc, Ng, (K - 1) // 2, (Kw - 1) // 2, Nb, Nx
        p.model, p.rank = code, int(mp.rho_rank) if code == 2 else 0
        p.param, p.grad = gsvi.flat_param.data_ptr(), gsvi.flat_grad.data_ptr()
        for f in _FIELD.values():
            setattr(p, f, -1)
        for n in names:
            setattr(p, _FIELD[n], gsvi.param_slices[n][0])
        keep = []  # tensors the struct points at

        def buf(t: torch.Tensor) -> int:
            t = t.detach().to(**f32).contiguous()
            keep.append(t)
            return t.data_ptr()

        def new(n: int) -> torch.Tensor:
            t = torch.zeros(max(int(n), 1), **f32)
            keep.append(t)
            return t

        p.mu_nu, p.sd_nu = buf(mp.μνg.reshape(Ng, K)), buf(mp.σνg.reshape(Ng, K))
        p.phixy_prior = buf(mp.φxy_prior.reshape(Nc, 2))
        p.gamma_alpha, p.gamma_beta = float(mp.gamma_alpha), float(mp.gamma_beta)
        # Delta-nu prior: Normal(0, sigma_dnu) in the phase model, Normal(0, 0.01) in the velocity models
        p.sd_dnu = float(mp.σΔν) if code == 0 else 0.01
        if velocity:
            rep = lambda t: t.to(dev).reshape(-1).expand(Ng) if t.numel() == 1 else t.reshape(Ng)
            p.mu_loggamma, p.sd_loggamma = buf(rep(mp.μγ)), buf(rep(mp.σγ))
            p.mu_logbeta, p.sd_logbeta = buf(rep(mp.μβ)), buf(rep(mp.σβ))
            p.mu_nuw, p.sd_nuw = buf(mp.μνω.reshape(Nx, Kw)), buf(mp.σνω.reshape(Nx, Kw))
        if code == 2:
            p.rho_mean, p.rho_std, p.rho_scale = float(mp.rho_mean), float(mp.rho_std), float(mp.rho_scale)
        # sampled values
        self.nu, self.shape_inv, self.phi = new(Ng * K), new(Ng), new(Nc)
        self.phixy = new(2 * Nc)
        p.nu, p.shape_inv, p.phi, p.phixy = (t.data_ptr() for t in (self.nu, self.shape_inv, self.phi, self.phixy))
        self.dnu = new(Nb * Ng) if Nb else None
        p.dnu = _ptr(self.dnu)
        if velocity:
            self.loggamma, self.gamma, self.logbeta, self.nu_omega = new(Ng), new(Ng), new(Ng), new(Nx * Kw)
            p.loggamma, p.gamma, p.logbeta, p.nu_omega = (t.data_ptr() for t in (self.loggamma, self.gamma, self.logbeta,
                                                                                self.nu_omega))
        # scratch
        ncb, ngb = C.c_int64(0), C.c_int64(0)
        _lib.check(self.lib.vcb_svi_partials(Nc, Ng, C.byref(ncb), C.byref(ngb)), "vcb_svi_partials")
        self.partials = torch.zeros(ncb.value + 2 * ngb.value, dtype=torch.float64, device=dev)
        p.cell_partials = self.partials.data_ptr()
        p.gene_partials = self.partials.data_ptr() + 8 * ncb.value
        p.lik_partials = self.partials.data_ptr() + 8 * (ncb.value + ngb.value)
        p.loss = gsvi.loss_buf.data_ptr()
        # conditioned sites: fixed values in the layouts the kernels index ([Ng][K], [Nb][Ng], [Ng], [Nc][2])
        shapes = {"ν": (Ng * K,), "Δν": (max(Nb, 1) * Ng,), "shape_inv": (Ng,), "ϕxy": (2 * Nc,)}
        for site, value in (conditioned or {}).items():
            v = torch.as_tensor(value)
            if v.numel() != shapes[site][0]:
                raise _lib.VcbError(f"conditioned site {site!r} has {v.numel()} values, expected {shapes[site][0]}")
            setattr(p, _CONDITIONABLE[site], buf(v.reshape(-1)))
        self.conditioned = sorted(conditioned or ())
        self.p, self._keep = p, keep
        self.K, self.Nb, self.Nx, self.Kw, self.Nc, self.Ng = K, Nb, Nx, Kw, Nc, Ng
        self._prepare_likelihood()

    # ------------------------------------------------------------------------------------------------------
    def _prepare_likelihood(self) -> None:
        """The likelihood call on the sampled values: persistent outputs, one flat gene-level buffer (the all-reduce payload,
        with one extra slot for the cells' log p - log q), persistent workspace."""
        counts, dev = self.counts, self.dev
        velocity = self.code != 0
        Ng, K, Nb, Nx, Kw, Nc = self.Ng, self.K, self.Nb, self.Nx, self.Kw, self.Nc
        q = VcbProblem()
        q.Nc, q.Ng, q.ld = counts.Nc, counts.Ng, counts.ld
        q.H, q.Nb = (K - 1) // 2, Nb
        q.flags = VCB_FLAG_GRAD
        q.S = counts.S.data_ptr()
        cf = self.mp.count_factor.detach().reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
        if counts.perm is not None:  # rows sorted by batch (PackedCounts): phi / d_phi live in row order, cf is permuted once
            cf = cf[counts.perm].contiguous()
            self.p.cell_row = counts.inv_perm.data_ptr()
        self._keep.append(cf)
        q.phi, q.cf = self.phi.data_ptr(), cf.data_ptr()
        if Nb > 0 and counts.batch_id is None:
            if Nb != 1:
                raise _lib.VcbError("Δν with more than one batch needs batch ids")
            counts.batch_id = torch.zeros(counts.Nc, dtype=torch.int32, device=dev)
        q.batch_id = _ptr(counts.batch_id) if Nb > 0 else None
        q.nu, q.dnu, q.shape_inv = self.nu.data_ptr(), _ptr(self.dnu), self.shape_inv.data_ptr()
        if velocity:
            if counts.U is None:
                raise _lib.VcbError("velocity model needs unspliced counts")
            if Nx > 1 and counts.cond_id is None:
                raise _lib.VcbError("more than one condition needs condition ids")
            q.Hw, q.Nx = (Kw - 1) // 2, Nx
            q.U = counts.U.data_ptr()
            q.cond_id = _ptr(counts.cond_id)
            q.logbeta, q.gamma, q.nu_omega = self.logbeta.data_ptr(), self.gamma.data_ptr(), self.nu_omega.data_ptr()
        sizes = [("lp_S", Ng), ("lp_U", Ng if velocity else 0), ("d_shape_inv", Ng), ("d_logbeta", Ng if velocity else 0),
                 ("d_gamma", Ng if velocity else 0), ("d_nu", Ng * K), ("d_dnu", Nb * Ng), ("d_nu_omega", Nx * Kw if velocity else 0)]
        n_flat = sum(n for _, n in sizes) + 1
        self.gene_flat = torch.zeros((n_flat + 3) // 4 * 4, dtype=torch.float32, device=dev)  # (128-bit exchange)
        # the step's exchange under cell sharding: one kernel over NVLink peer memory, NCCL when the mapping is not possible
        self.comm = PeerComm.create(self.shard, self.gene_flat.numel(), dev) if self.g.peer_allreduce else None
        off = 0
        base = self.gene_flat.data_ptr()
        for name, n in sizes:
            if n:
                setattr(q, name, base + 4 * off)
                setattr(self.p, name, base + 4 * off)
            off += n
        self.p.cell_lp = base + 4 * off
        self.d_phi = torch.zeros(Nc, dtype=torch.float32, device=dev)
        self.d_cf = torch.zeros(Nc, dtype=torch.float32, device=dev)
        q.d_phi, q.d_cf = self.d_phi.data_ptr(), self.d_cf.data_ptr()
        self.p.d_phi = self.d_phi.data_ptr()
        counts.build_spectra()
        q.spec_S = counts.spec_S.as_struct()
        if velocity:
            q.spec_U = counts.spec_U.as_struct()
        ev = getattr(counts, "profile_events", None)  # (begin, end) raw cudaEvent_t handles, set by bench.py
        if ev is not None:
            q.ev_stream_begin, q.ev_stream_end = ev
        self.ws_bytes = self.lib.vcb_workspace_bytes(C.byref(q))
        self.ws = torch.empty(max(self.ws_bytes, 16), dtype=torch.uint8, device=dev)
        self.q = q
        self.like_fn = self.lib.vcb_velocity_fwd_bwd if velocity else self.lib.vcb_phase_fwd_bwd

    # ------------------------------------------------------------------------------------------------------
    def _draw(self) -> None:
        """Standard-normal draws with the shapes and in the order of the traced guide (torch's generator)."""
        dev, p = self.dev, self.p
        Ng, K, Nc = self.Ng, self.K, self.Nc
        n = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev).normal_()
        e: Dict[str, torch.Tensor] = {}
        if self.code == 2:  # velocity_inference_guide.py:104-106 (eps_W, eps_D), then the sample sites nu, log beta
            # the guide re-evaluates cov_factor's random initialiser on every call (:91-94): same RNG consumption here
            n(Ng + self.Nx * self.Kw, int(self.mp.rho_rank))
            e["eps_W"] = n(int(self.mp.rho_rank))
            e["eps_D"] = n(Ng + self.Nx * self.Kw)
            e["eps_nu"] = n(Ng, 1, K)
            e["eps_logbeta"] = n(Ng, 1)
        elif self.code == 1:  # velocity_inference_guide.py:45-63: log gamma, log beta, nu, nu_omega
            e["eps_loggamma"] = n(Ng, 1)
            e["eps_logbeta"] = n(Ng, 1)
            e["eps_nu"] = n(Ng, 1, K)
            e["eps_nuw"] = n(self.Nx, self.Kw, 1, 1)
        else:  # phase_inference_guide.py:47-56
            e["eps_nu"] = n(Ng, 1, K)
        shard = self.shard
        if shard is not None and shard.world > 1:  # the rank's rows of the draw a single process would make (ShardedNormal)
            full = n(shard.Nc_global, 2)
            e["eps_phixy"] = full[shard.cell_offset: shard.cell_offset + Nc]
        else:
            e["eps_phixy"] = n(Nc, 2)
        for k, t in e.items():
            setattr(p, k, t.data_ptr())
        self.eps = e  # alive until the step's kernels have run (and the graph's private pool keeps the addresses)

    def body(self, mark=None) -> None:
        """Enqueue one step on the current stream (graph capturable).  ``mark(name)`` is called at the phase boundaries
        (bench.py records CUDA events there: draws / sample / likelihood / allreduce / backward)."""
        mark = mark or (lambda name: None)
        st = torch.cuda.current_stream(self.dev).cuda_stream
        self._draw()
        mark("draws")
        _lib.check(self.lib.vcb_svi_sample(C.byref(self.p), st), "vcb_svi_sample")
        mark("sample")
        _lib.check(self.like_fn(C.byref(self.q), self.ws.data_ptr(), self.ws_bytes, st), "likelihood")
        mark("likelihood")
        if self.comm is not None:
            self.comm.allreduce_(self.gene_flat)
        else:
            allreduce_flat_(self.gene_flat, self.shard)
        mark("allreduce")
        _lib.check(self.lib.vcb_svi_backward(C.byref(self.p), st), "vcb_svi_backward")
        mark("backward")
