"""Posterior summaries the reference's fit drivers attach to ``self.posterior`` after the SVI loop: the expected log counts
``ElogS`` / ``ElogU`` at the fitted parameters and their size-factor-averaged versions ``ElogS2`` / ``ElogU2``
(``phase_inference_model.py:241-256``, ``velocity_inference_model.py:232-260``).  Plain torch on flat tensors, evaluated once
after the fit (not on the per-step path); the drivers skip them above ``max_dense_elements`` because each is a dense (Ng, Nc)
matrix -- 8 GB at 1M x 2k -- which the reference materialises unconditionally.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .utils import torch_fourier_basis

__all__ = ["expected_log_spliced", "expected_log_unspliced", "expected_log_counts_summary"]


def expected_log_spliced(nu: torch.Tensor, phis: torch.Tensor, count_factor: torch.Tensor, dnu: Optional[torch.Tensor] = None,
                         batch_id: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``ElogS[g, c] = nu[g] . zeta(phi_c) + dnu[batch_c, g] + count_factor[c]``; nu (Ng, K), dnu (Nb, Ng)."""
    H = (nu.shape[-1] - 1) // 2
    zeta = torch_fourier_basis(phis.reshape(-1), num_harmonics=H, der=0)          # (Nc, K)
    out = nu @ zeta.T + count_factor.reshape(1, -1)
    if dnu is not None:
        out = out + dnu[batch_id.long()].T
    return out


def expected_log_unspliced(ElogS: torch.Tensor, nu: torch.Tensor, phis: torch.Tensor, nu_omega: torch.Tensor, cond_id: torch.Tensor,
                           gamma: torch.Tensor, logbeta: torch.Tensor) -> torch.Tensor:
    """``ElogU = -logbeta_g + log(relu(nu_g . zeta'(phi_c) omega_c + gamma_g) + 1e-5) + ElogS`` with
    ``omega_c = nu_omega[cond_c] . zeta_omega(phi_c)``; nu_omega (Nx, Kw), gamma / logbeta (Ng,)."""
    H = (nu.shape[-1] - 1) // 2
    Hw = (nu_omega.shape[-1] - 1) // 2
    phis = phis.reshape(-1)
    dzeta = torch_fourier_basis(phis, num_harmonics=H, der=1)                     # (Nc, K)
    zeta_w = torch_fourier_basis(phis, num_harmonics=Hw, der=0)                   # (Nc, Kw)
    omega = (zeta_w * nu_omega[cond_id.long()]).sum(-1)                           # (Nc,)
    a = (nu @ dzeta.T) * omega.reshape(1, -1) + gamma.reshape(-1, 1)
    return -logbeta.reshape(-1, 1) + torch.log(torch.relu(a) + 1e-5) + ElogS


def expected_log_counts_summary(nu, phis, count_factor, dnu=None, batch_id=None, velocity: Optional[dict] = None
                                ) -> Dict[str, torch.Tensor]:
    """The four (two without ``velocity``) matrices the reference stores; ``velocity`` = dict(nu_omega, cond_id, gamma, logbeta).
    The "2" versions replace every cell's size factor by the mean size factor (``metaparams_avg``)."""
    cf = count_factor.reshape(-1)
    out = {"ElogS": expected_log_spliced(nu, phis, cf, dnu, batch_id),
           "ElogS2": expected_log_spliced(nu, phis, torch.full_like(cf, float(cf.mean())), dnu, batch_id)}
    if velocity is not None:
        for tag in ("", "2"):
            out["ElogU" + tag] = expected_log_unspliced(out["ElogS" + tag], nu, phis, velocity["nu_omega"], velocity["cond_id"],
                                                        velocity["gamma"], velocity["logbeta"])
    return out
