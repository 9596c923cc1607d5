"""Guides of the velocity model: mean-field (``model_type="normal"``) and the low-rank multivariate-normal
variant (``model_type="lrmn"``, the reference's default).

Parameter names / shapes / constraints and the RNG draw order follow
``velocycle/velocity_inference_guide.py:9-63`` (mean-field: log gamma, log beta, nu, [Delta sites], nu_omega,
phixy) and ``:65-141`` (LRMN: one joint LowRankMultivariateNormal draw for (log gamma, nu_omega) taken outside
``pyro.sample`` and injected through Delta sites; nu; log beta | log gamma conditional Normal; phixy).
"""
from __future__ import annotations

import torch

from ._const import const
from .phase_inference_guide import _phixy_guide_dist
from .ppl import backend

__all__ = ["velocity_latent_variable_guide", "velocity_latent_variable_guide_LRMN"]


def _plates(pyro, mp, device=None):
    kw = {} if device is None else {"device": device}
    return (
        pyro.plate("cells", mp.Nc, dim=-1, **kw),
        pyro.plate("genes", mp.Ng, dim=-2, **kw),
        pyro.plate("harmonics", mp.Nhω, dim=-3, **kw),
        pyro.plate("conditions", mp.Nx, dim=-4, **kw),
        pyro.plate("batches", mp.Nb, dim=-5, **kw),
    )


def velocity_latent_variable_guide(mp):
    pyro, dist, _, _, _ = backend.get()
    dev = mp.device
    positive = dist.constraints.positive
    cells, genes, harmonics, conditions, batches = _plates(pyro, mp, dev)
    init = lambda t: t.detach().clone().to(dev)

    loggamma_locs = pyro.param("logγg_locs", init(mp.μγ))
    logbeta_locs = pyro.param("logβg_locs", init(mp.μβ))
    loggamma_scales = pyro.param("logγg_scales", init(mp.σγ), constraint=positive)
    logbeta_scales = pyro.param("logβg_scales", init(mp.σβ), constraint=positive)
    nu_locs = pyro.param("ν_locs", init(mp.μνg))
    nu_scales = pyro.param("ν_scales", init(mp.σνg), constraint=positive)
    if mp.with_delta_nu:
        dnu_locs = pyro.param("Δν_locs", torch.ones((mp.Nb, 1, 1, mp.Ng, 1), device=dev) * mp.μΔν.to(dev))
    phixy_locs = pyro.param("ϕxy_locs", init(mp.φxy_prior))
    nuw_locs = pyro.param("νω_locs", init(mp.μνω))
    nuw_scales = pyro.param("νω_scales", init(mp.σνω), constraint=positive)
    nb = mp.noisemodel == "NegativeBinomial"
    if nb:
        shape_inv_locs = pyro.param(
            "shape_inv_locs", (torch.ones((mp.Ng, 1), device=dev) * mp.gamma_alpha / mp.gamma_beta).to(dev),
            constraint=positive,
        )

    with genes:
        pyro.sample("logγg", dist.Normal(loggamma_locs, loggamma_scales))
        pyro.sample("logβg", dist.Normal(logbeta_locs, logbeta_scales))
        pyro.sample("ν", dist.Normal(nu_locs, nu_scales).to_event(1))
        if mp.with_delta_nu:
            with batches:
                pyro.sample("Δν", dist.Delta(dnu_locs))
        if nb:
            pyro.sample("shape_inv", dist.Delta(shape_inv_locs))
    with harmonics, conditions:
        pyro.sample("νω", dist.Normal(nuw_locs, nuw_scales))
    with cells:
        pyro.sample("ϕxy", _phixy_guide_dist(dist, mp, phixy_locs, const(1.0, dev)))


def velocity_latent_variable_guide_LRMN(mp):
    pyro, dist, _, _, _ = backend.get()
    dev = mp.device
    positive = dist.constraints.positive
    cells, genes, harmonics, conditions, batches = _plates(pyro, mp)
    Ng = mp.Ng

    nu_locs = pyro.param("ν_locs", mp.μνg.detach().clone())
    nu_scales = pyro.param("ν_scales", mp.σνg.detach().clone(), constraint=positive)
    if mp.with_delta_nu:
        dnu_locs = pyro.param("Δν_locs", torch.ones((mp.Nb, 1, 1, Ng, 1), device=dev) * mp.μΔν.to(dev))
    phixy_locs = pyro.param("ϕxy_locs", mp.φxy_prior.detach().clone())
    logbeta_locs = pyro.param("logβg_locs", mp.μβ.detach().clone())
    logbeta_scales = pyro.param("logβg_scales", mp.σβ.detach().clone(), constraint=positive)

    # joint low-rank Gaussian over (log gamma_g for every gene, every nu_omega coefficient)
    n_joint = Ng + mp.Nhω * mp.Nx
    rank = int(mp.rho_rank)
    loc = pyro.param("loc", torch.hstack([mp.μγ.squeeze().detach().clone(), mp.μνω.squeeze().detach().clone().flatten()]))
    cov_factor = pyro.param(
        "cov_factor",
        # == torch.clip(torch.normal(zeros, ones * 0.02), min=0): that overload draws N(0,1) and scales it; this
        # spelling consumes the RNG stream identically (the reference re-evaluates the initialiser on every
        # call) but has no host-side check of ``std``, so it can be captured into a CUDA graph
        torch.clip(torch.empty((n_joint, rank), device=dev).normal_() * 0.02, min=0, max=None),
        constraint=positive,
    )
    cov_diag = pyro.param(
        "cov_diag",
        (torch.hstack([mp.σγ.squeeze().detach().clone(), mp.σνω.squeeze().detach().clone().flatten()]) ** 2).to(dev),
        constraint=positive,
    )
    # = LowRankMultivariateNormal(loc, cov_factor, cov_diag).rsample(): same two standard-normal draws in the
    # same order (eps_W then eps_D), without the constructor's Cholesky of the capacitance matrix, which the
    # reference pays every step although only rsample is used (and which would force a host sync on CUDA)
    from torch.distributions.utils import _standard_normal

    eps_W = _standard_normal(cov_factor.shape[-1:], dtype=loc.dtype, device=loc.device)
    eps_D = _standard_normal(loc.shape, dtype=loc.dtype, device=loc.device)
    joint = loc + torch.matmul(cov_factor, eps_W.unsqueeze(-1)).squeeze(-1) + cov_diag.sqrt() * eps_D
    rho_real_loc = pyro.param("rho_real_loc", torch.ones(Ng, device=dev) * mp.rho_mean)
    nb = mp.noisemodel == "NegativeBinomial"
    if nb:
        shape_inv_locs = pyro.param(
            "shape_inv_locs", torch.ones((Ng, 1), device=dev) * mp.gamma_alpha / mp.gamma_beta, constraint=positive
        )

    with genes:
        loggamma = pyro.sample("logγg", dist.Delta(joint[:Ng].unsqueeze(-1)))
        pyro.sample("ν", dist.Normal(nu_locs, nu_scales).to_event(1))
        rho_real = pyro.sample("rho_real", dist.Delta(rho_real_loc.unsqueeze(-1)))
        rho = torch.sigmoid(rho_real / mp.rho_scale) * 1.998 - 0.999
        if mp.with_delta_nu:
            with batches:
                pyro.sample("Δν", dist.Delta(dnu_locs.to(dev)))
        if nb:
            pyro.sample("shape_inv", dist.Delta(shape_inv_locs))

    # log beta | log gamma: bivariate-normal conditional with per-gene correlation rho
    gamma_sd = torch.sqrt(torch.diag(cov_factor @ cov_factor.T + torch.diag(cov_diag))[:Ng])
    cond_mean = logbeta_locs.squeeze() + rho.squeeze() * logbeta_scales.squeeze() * (loggamma.squeeze() - loc[:Ng]) / gamma_sd
    cond_sd = logbeta_scales.squeeze() * torch.sqrt(1 - rho.squeeze() ** 2)
    with genes:
        pyro.sample("logβg", dist.Normal(cond_mean.unsqueeze(-1), cond_sd.unsqueeze(-1)))

    with harmonics, conditions:
        tail = joint[Ng:]
        if mp.Nx > 1:
            tail = tail.reshape((mp.Nx, mp.Nhω))
        pyro.sample("νω", dist.Delta(tail.unsqueeze(-1).unsqueeze(-1)))
    with cells:
        pyro.sample("ϕxy", _phixy_guide_dist(dist, mp, phixy_locs, const(1.0, dev)))
