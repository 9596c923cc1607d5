"""Posterior draws of the latent and deterministic sites, batched over the draws.

The fit drivers finish with ``Predictive(model, guide=guide, num_samples=500, return_sites=[...])`` in bins of 50
(``phase_inference_model.py:216-263``, ``velocity_inference_model.py:198-262``): per draw one guide trace and one model
replay through the effect handlers -- and, in the reference, both GammaPoisson sites over the (Ng, Nc) matrices although
nobody asks for them.  For the package's own model / guide pairs (unconditioned, or conditioned on ν, Δν, shape_inv, ϕxy
like the tutorials' velocity stage) the requested sites are functions of the guide parameters and the standard-normal
draws alone, so a bin is evaluated here in one go: the draws are made per sample in the guide's order with torch's
generator (a seed gives the same values as the sequential ``Predictive``), everything else is a handful of tensor ops over
(num_samples, ...) arrays on the device.  No count byte is read.
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import torch

from .ppl import backend
from .utils import pack_direction, torch_fourier_basis

__all__ = ["batched_posterior"]


def _cond_ids(counts, Nc: int, dev) -> torch.Tensor:
    """Condition id of every cell in the CALLER's cell order (PackedCounts may hold the rows sorted by batch)."""
    if counts is None or counts.cond_id is None:
        return torch.zeros(Nc, dtype=torch.long, device=dev)
    cid = counts.cond_id if counts.perm is None else counts.cond_id[counts.inv_perm]
    return cid.long()


@torch.no_grad()
def batched_posterior(mp, code: int, conditioned: Optional[dict], num_samples: int, return_sites: Iterable[str],
                      counts=None, pad: Optional[Dict[str, int]] = None) -> Dict[str, torch.Tensor]:
    """``{site: (num_samples, 1, ..., 1, *site_shape)}`` on ``mp.device`` for the sites of ``return_sites`` that the model
    defines; ``code`` / ``conditioned`` as returned by ``faststep.model_code``; ``pad[site]`` = the number of singleton dims
    ``Predictive`` puts behind the sample dim (``ppl.infer.predictive_padding``; none when ``pad`` is None)."""
    pyro, _, _, _, _ = backend.get()
    dev = torch.device(mp.device)
    n, rs, cond = int(num_samples), set(return_sites), dict(conditioned or {})
    Nc, Ng = int(mp.Nc), int(mp.Ng)
    K = int(mp.μνg.shape[-1])
    H = (K - 1) // 2
    velocity = code != 0
    par = lambda name: pyro.param(name).detach()
    normal = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev).normal_()
    shard = getattr(mp, "shard", None)

    # ---- the draws, sample by sample, in the order the guide makes them (faststep.FusedStep._draw) ----------------------
    eps = {k: [] for k in ("W", "D", "nu", "loggamma", "logbeta", "nuw", "phixy")}
    if code == 2:
        n_joint, rank = Ng + int(mp.Nx) * int(mp.Nhω), int(mp.rho_rank)
    for _ in range(n):
        if code == 2:
            normal(n_joint, rank)  # the guide re-evaluates cov_factor's random initialiser on every call
            eps["W"].append(normal(rank))
            eps["D"].append(normal(n_joint))
            eps["nu"].append(normal(Ng, 1, K))
            eps["logbeta"].append(normal(Ng, 1))
        elif code == 1:
            eps["loggamma"].append(normal(Ng, 1))
            eps["logbeta"].append(normal(Ng, 1))
            eps["nu"].append(normal(Ng, 1, K))
            eps["nuw"].append(normal(int(mp.Nx), int(mp.Nhω), 1, 1))
        else:
            eps["nu"].append(normal(Ng, 1, K))
        if shard is not None and shard.world > 1:
            full = normal(shard.Nc_global, 2)
            eps["phixy"].append(full[shard.cell_offset: shard.cell_offset + Nc])
        else:
            eps["phixy"].append(normal(Nc, 2))
    E = {k: torch.stack(v) for k, v in eps.items() if v}
    rep = lambda t: t.unsqueeze(0).expand(n, *t.shape)
    given = lambda site: rep(torch.as_tensor(cond[site]).to(dev))

    out: Dict[str, torch.Tensor] = {}
    # ---- gene-level / global sites ------------------------------------------------------------------------------------------
    nu = given("ν") if "ν" in cond else par("ν_locs") + par("ν_scales") * E["nu"]
    out["ν"] = nu
    out["shape_inv"] = given("shape_inv") if "shape_inv" in cond else rep(par("shape_inv_locs"))
    if mp.with_delta_nu:
        out["Δν"] = given("Δν") if "Δν" in cond else rep(par("Δν_locs"))
    if code == 1:
        loggamma = par("logγg_locs") + par("logγg_scales") * E["loggamma"]
        logbeta = par("logβg_locs") + par("logβg_scales") * E["logbeta"]
        nuw = par("νω_locs") + par("νω_scales") * E["nuw"]
    elif code == 2:
        loc, W, D = par("loc"), par("cov_factor"), par("cov_diag")
        joint = loc + E["W"] @ W.T + D.sqrt() * E["D"]                          # (n, n_joint)
        loggamma = joint[:, :Ng].unsqueeze(-1)
        rho_real = rep(par("rho_real_loc").unsqueeze(-1))
        out["rho_real"] = rho_real
        rho = torch.sigmoid(rho_real.squeeze(-1) / mp.rho_scale.to(dev)) * 1.998 - 0.999
        gamma_sd = torch.sqrt((W[:Ng] * W[:Ng]).sum(-1) + D[:Ng])
        s = par("logβg_scales").squeeze()
        cond_mean = par("logβg_locs").squeeze() + rho * s * (joint[:, :Ng] - loc[:Ng]) / gamma_sd
        cond_sd = s * torch.sqrt(1 - rho ** 2)
        logbeta = (cond_mean + cond_sd * E["logbeta"].squeeze(-1)).unsqueeze(-1)
        tail = joint[:, Ng:]
        # (for Nx = 1 the guide hands over a (Kw, 1, 1) value that the harmonics / conditions plates expand to (1, Kw, 1, 1))
        nuw = tail.reshape(n, int(mp.Nx), int(mp.Nhω)).unsqueeze(-1).unsqueeze(-1)
    if velocity:
        out["logγg"], out["logβg"], out["νω"] = loggamma, logbeta, nuw
        out["γg"] = torch.exp(loggamma)
    # ---- per-cell sites -------------------------------------------------------------------------------------------------------
    phixy = given("ϕxy") if "ϕxy" in cond else par("ϕxy_locs") + E["phixy"]
    out["ϕxy"] = phixy
    need_cells = rs & {"ϕ", "ζ", "ζ_dϕ", "ζω", "ω"}
    if need_cells:
        phi = pack_direction(phixy)                                              # (n, Nc)
        out["ϕ"] = phi
        if "ζ" in rs:
            out["ζ"] = torch_fourier_basis(phi, num_harmonics=H, der=0)
        if velocity:
            if "ζ_dϕ" in rs:
                out["ζ_dϕ"] = torch_fourier_basis(phi, num_harmonics=H, der=1)
            if rs & {"ζω", "ω"}:
                Hw = int(mp.kwargsζω["num_harmonics"])
                zw = torch_fourier_basis(phi, num_harmonics=Hw, der=0)           # (n, Nc, Kw)
                out["ζω"] = zw.transpose(1, 2)
                if "ω" in rs:
                    cid = _cond_ids(counts, Nc, dev)
                    nw = nuw.reshape(n, int(mp.Nx), int(mp.Nhω))
                    out["ω"] = (nw[:, cid, :] * zw).sum(-1).unsqueeze(1)
    pad = pad or {}
    return {k: v.reshape((n,) + (1,) * int(pad.get(k, 0)) + tuple(v.shape[1:])) for k, v in out.items() if k in rs}
