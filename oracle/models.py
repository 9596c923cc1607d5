"""Oracle: the reference's model functions restated UNFUSED (TEST INFRASTRUCTURE).

Same sites, priors and (Ng,Nc) op chain as ``velocycle/phase_inference_model.py:343-395`` and
``velocycle/velocity_inference_model.py:304-388`` / ``:390-471`` -- einsums, relu + 1e-5, two GammaPoisson sites
inside the cells x genes plates -- written against the ``velocycle_b200.ppl`` runtime.  They run on any device,
so a GPU test can drive them with the same RNG stream as the fused models and compare whole SVI trajectories;
on the CPU they are the timed baseline of ``bench.py --impl reference``.
"""
from __future__ import annotations

import torch

from velocycle_b200 import ppl as pyro
from velocycle_b200.ppl import distributions as dist

from .likelihood import fourier_basis, pack_direction


def phase_model_unfused(mp):
    dev = mp.device
    cells = pyro.plate("cells", mp.Nc, dim=-1)
    genes = pyro.plate("genes", mp.Ng, dim=-2)
    batches = pyro.plate("batches", mp.Nb, dim=-3)
    with genes:
        nu = pyro.sample("ν", dist.Normal(mp.μνg, mp.σνg).to_event(1))
        with batches:
            if mp.with_delta_nu:
                dnu = pyro.sample("Δν", dist.Normal(0, mp.σΔν))
    with cells:
        phixy = pyro.sample("ϕxy", dist.Normal(mp.φxy_prior, torch.tensor(1.0, device=dev)).to_event(1))
    phi = pack_direction(phixy)
    zeta = fourier_basis(phi.squeeze(), mp.num_harmonics_S, der=0).to(dev)
    cf = mp.count_factor
    if mp.with_delta_nu:
        ElogS = torch.einsum("...gch,ch->gc", nu, zeta) + torch.einsum("bgc,bgc->gc", mp.Db, dnu) + cf
    else:
        ElogS = torch.einsum("...gch,ch->gc", nu, zeta) + cf
    with genes:
        shape_inv = pyro.sample("shape_inv", dist.Gamma(mp.gamma_alpha, mp.gamma_beta))
    with cells, genes:
        pyro.sample("S", dist.GammaPoisson(1.0 / shape_inv, 1.0 / (shape_inv * torch.exp(ElogS))), obs=mp.S)


def _velocity_unfused(mp, lrmn):
    dev = mp.device
    cells = pyro.plate("cells", mp.Nc, dim=-1)
    genes = pyro.plate("genes", mp.Ng, dim=-2)
    harmonics = pyro.plate("harmonics", mp.Nhω, dim=-3)
    conditions = pyro.plate("conditions", mp.Nx, dim=-4)
    batches = pyro.plate("batches", mp.Nb, dim=-5)
    with genes:
        loggamma = pyro.sample("logγg", dist.Normal(mp.μγ, mp.σγ))
        logbeta = pyro.sample("logβg", dist.Normal(mp.μβ, mp.σβ))
        if lrmn:
            pyro.sample("rho_real", dist.Normal(mp.rho_mean, mp.rho_std))
        gamma = torch.exp(loggamma)
        nu = pyro.sample("ν", dist.Normal(mp.μνg, mp.σνg).to_event(1))
        if mp.with_delta_nu:
            with batches:
                dnu = pyro.sample("Δν", dist.Normal(torch.tensor(0.0, device=dev), torch.tensor(0.01, device=dev)))
    with cells:
        phixy = pyro.sample("ϕxy", dist.Normal(mp.φxy_prior, torch.tensor(1.0, device=dev)).to_event(1))
    phi = pack_direction(phixy)
    H, Hw = mp.kwargsζ["num_harmonics"], mp.kwargsζω["num_harmonics"]
    zeta = fourier_basis(phi, H, 0).to(dev)
    zeta_d = fourier_basis(phi, H, 1).to(dev)
    with harmonics, conditions:
        nu_omega = pyro.sample("νω", dist.Normal(mp.μνω, mp.σνω))
    zeta_w = fourier_basis(phi, Hw, 0).to(dev).T
    if mp.with_delta_nu:
        ElogS = torch.einsum("...gch,...ch->gc", nu, zeta) + torch.einsum("bxhgc,bxhgc->gc", mp.Db, dnu) + mp.count_factor
    else:
        ElogS = torch.einsum("...gch,...ch->gc", nu, zeta) + mp.count_factor
    omega = torch.einsum("...xhgc,hc...,xhgc->gc", [nu_omega, zeta_w, mp.D])
    ElogU = -logbeta + torch.log(torch.relu(torch.einsum("...gch,...ch->gc", nu, zeta_d) * omega + gamma) + 1e-5) + ElogS
    with genes:
        shape_inv = pyro.sample("shape_inv", dist.Gamma(mp.gamma_alpha, mp.gamma_beta))
    with cells, genes:
        pyro.sample("S", dist.GammaPoisson(1.0 / shape_inv, 1.0 / (shape_inv * torch.exp(ElogS))), obs=mp.S)
        pyro.sample("U", dist.GammaPoisson(1.0 / shape_inv, 1.0 / (shape_inv * torch.exp(ElogU))), obs=mp.U)


def velocity_model_unfused(mp):
    return _velocity_unfused(mp, lrmn=False)


def velocity_model_unfused_lrmn(mp):
    return _velocity_unfused(mp, lrmn=True)
