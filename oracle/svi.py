"""Oracle: closed-form model log-joint of the phase / velocity models and its gradient w.r.t. the latents
(TEST INFRASTRUCTURE; never imported by the product).

Independent of both the reference's op chain and of ``velocycle_b200.ppl``: priors are written out by hand
(Normal / Gamma log-densities and their derivatives), the likelihood uses ``oracle.likelihood.analytic_gradients``.
Restates ``phase_inference_model.py:361-393`` and ``velocity_inference_model.py:323-386`` (LRMN: ``:404-469``).
Pinned against ``tests/golden/case_*.npz``, which were produced by executing the reference's own model source.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .likelihood import analytic_gradients

LOG_SQRT_2PI = 0.5 * math.log(2.0 * math.pi)


def normal_lp(x, mu, sd):
    """sum log N(x | mu, sd) and d/dx."""
    x, mu, sd = torch.broadcast_tensors(x.double(), torch.as_tensor(mu).double(), torch.as_tensor(sd).double())
    z = (x - mu) / sd
    return (-0.5 * z * z - torch.log(sd) - LOG_SQRT_2PI).sum(), -z / sd


def gamma_lp(x, alpha, beta):
    """sum log Gamma(x | alpha, rate beta) and d/dx."""
    x = x.double()
    a, b = float(alpha), float(beta)
    lp = a * math.log(b) + (a - 1.0) * torch.log(x) - b * x - math.lgamma(a)
    return lp.sum(), (a - 1.0) / x - b


def model_logjoint(kind: str, inp: Dict[str, torch.Tensor], draws: Dict[str, torch.Tensor],
                   with_delta_nu: bool = True) -> Dict[str, torch.Tensor]:
    """kind in {"phase", "velocity", "velocity_lrmn"}.  ``inp`` as stored in the golden files (``in/...``),
    ``draws`` the latent values keyed by site name.  Returns per-site log-probs ``lp/<site>`` and
    ``d/<site>`` = d(log-joint)/d(site value) in the site's own shape."""
    velocity = kind != "phase"
    S = inp["S"].double()
    Nc, Ng = S.shape
    H, Hw, Nb, Nx = int(inp["H"]), int(inp["Hw"]), int(inp["Nb"]), int(inp["Nx"])
    out: Dict[str, torch.Tensor] = {}
    nu = draws["ν"].double().reshape(Ng, -1)
    phixy = draws["ϕxy"].double()
    sinv = draws["shape_inv"].double().reshape(Ng)
    x, y = phixy[:, 0], phixy[:, 1]
    phi = torch.atan2(y, x)

    lp, g = normal_lp(nu, inp["mu_nu"], inp["sd_nu"])
    out["lp/ν"], d_nu = lp, g
    lp, g = normal_lp(phixy, inp["phixy_prior"], 1.0)
    out["lp/ϕxy"], d_phixy = lp, g
    lp, g = gamma_lp(sinv, 1.0, 2.0)
    out["lp/shape_inv"], d_sinv = lp, g
    prob = dict(S=S, phi=phi, cf=inp["cf"].double(), batch_id=inp["batch_id"], nu=nu, shape_inv=sinv)
    dnu = None
    if with_delta_nu:
        dnu = draws["Δν"].double().reshape(Nb, Ng)
        lp, g = normal_lp(dnu, 0.0, 0.01 if velocity else 0.5)  # velocity models hard-code 0.01 (:332, :416)
        out["lp/Δν"], d_dnu = lp, g
        prob["dnu"] = dnu
    if velocity:
        lg = draws["logγg"].double().reshape(Ng)
        lb = draws["logβg"].double().reshape(Ng)
        nw = draws["νω"].double().reshape(Nx, 2 * Hw + 1)
        lp, g = normal_lp(lg, 0.0, 0.5)
        out["lp/logγg"], d_lg = lp, g
        lp, g = normal_lp(lb, 2.0, 3.0)
        out["lp/logβg"], d_lb = lp, g
        lp, g = normal_lp(nw, inp["mu_nw"], inp["sd_nw"])
        out["lp/νω"], d_nw = lp, g
        if kind == "velocity_lrmn":
            rr = draws["rho_real"].double().reshape(Ng)
            lp, g = normal_lp(rr, 4.0, 1.0)
            out["lp/rho_real"], out["d/rho_real"] = lp, g.reshape(draws["rho_real"].shape)
        prob.update(U=inp["U"].double(), cond_id=inp["cond_id"], logbeta=lb, gamma=torch.exp(lg), nu_omega=nw)
    a = analytic_gradients(prob)
    out["lp/S"] = a["lp_S"].sum()
    if velocity:
        out["lp/U"] = a["lp_U"].sum()
    # chain rule through phi = atan2(y, x)
    r2 = x * x + y * y
    dphi = a["d_phi"]
    d_phixy = d_phixy + torch.stack([-y / r2 * dphi, x / r2 * dphi], dim=-1)
    out["d/ν"] = (d_nu + a["d_nu"]).reshape(draws["ν"].shape)
    out["d/ϕxy"] = d_phixy.reshape(draws["ϕxy"].shape)
    out["d/shape_inv"] = (d_sinv + a["d_shape_inv"]).reshape(draws["shape_inv"].shape)
    if with_delta_nu:
        out["d/Δν"] = (d_dnu + a["d_dnu"]).reshape(draws["Δν"].shape)
    if velocity:
        out["d/logγg"] = (d_lg + a["d_gamma"] * torch.exp(lg)).reshape(draws["logγg"].shape)
        out["d/logβg"] = (d_lb + a["d_logbeta"]).reshape(draws["logβg"].shape)
        out["d/νω"] = (d_nw + a["d_nu_omega"]).reshape(draws["νω"].shape)
    out["logjoint"] = sum(v for k, v in out.items() if k.startswith("lp/"))
    return out


def clipped_adam_reference(param, grad, m, v, step, lr0, lrd, betas, eps=1e-8, clip=10.0):
    """One ClippedAdam update in float64 (``pyro/optim/clipped_adam.py``): lr decays before use."""
    g = grad.double().clamp(-clip, clip)
    b1, b2 = betas
    m = b1 * m.double() + (1 - b1) * g
    v = b2 * v.double() + (1 - b2) * g * g
    lr = lr0 * lrd ** step
    step_size = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
    return param.double() - step_size * m / (v.sqrt() + eps), m, v
