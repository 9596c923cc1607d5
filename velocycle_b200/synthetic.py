"""Synthetic spliced/unspliced count matrices with known cell-cycle phases and angular speed.

The generative recipe is the one of the reference's (broken as shipped, sound as a recipe)
``velocycle/utils.py:508-584`` ``simulate_data``: per-gene Fourier coefficients
nu0 ~ N(0.4, 1.2), higher harmonics ~ N(0, 0.2), log gamma ~ N(0, 0.5), log beta ~ N(2, 1),
dispersion shape_inv ~ Gamma(1, 2), phases uniform on the circle, and negative-binomial counts
with means exp(ElogS), exp(ElogU) built exactly like the model
(``velocity_inference_model.py:359-368``).  Everything is generated on ``device`` in cell chunks so
that the 16 GB matrices of the large configurations never exist on the host.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import torch

__all__ = ["SyntheticCycleData", "make_synthetic", "fourier_rows"]


def fourier_rows(phi: torch.Tensor, H: int, der: int = 0) -> torch.Tensor:
    """(Nc, 2H+1) Fourier design rows in the reference's column order [1, sin, cos, sin2, cos2, ...]."""
    n = torch.arange(1, H + 1, device=phi.device, dtype=phi.dtype)
    arg = phi[:, None] * n
    s, c = torch.sin(arg), torch.cos(arg)
    if der == 0:
        first, a, b = torch.ones_like(phi), s, c
    elif der == 1:
        first, a, b = torch.zeros_like(phi), n * c, -n * s
    elif der == 2:
        first, a, b = torch.zeros_like(phi), -n * n * s, -n * n * c
    else:
        raise ValueError(f"Value {der=} is not allowed, use 0, 1 or 2 instead")
    return torch.cat([first[:, None], torch.stack([a, b], dim=-1).reshape(phi.shape[0], 2 * H)], dim=1)


@dataclass
class SyntheticCycleData:
    """Device-resident synthetic dataset (cell-major counts, genes contiguous, pitch ``ld``)."""

    S: torch.Tensor  # (Nc, ld) float32, columns >= Ng are zero padding
    U: torch.Tensor
    Nc: int
    Ng: int
    ld: int
    H: int
    Hw: int
    Nb: int
    Nx: int
    phi: torch.Tensor  # (Nc,) true phases
    cf: torch.Tensor  # (Nc,) log size factors (count_factor)
    batch_id: torch.Tensor  # (Nc,) int32
    cond_id: torch.Tensor  # (Nc,) int32
    nu: torch.Tensor  # (Ng, K)
    dnu: torch.Tensor  # (Nb, Ng)
    shape_inv: torch.Tensor  # (Ng,)
    logbeta: torch.Tensor
    loggamma: torch.Tensor
    nu_omega: torch.Tensor  # (Nx, Kw)
    zero_frac_S: float = field(default=float("nan"))
    zero_frac_U: float = field(default=float("nan"))


def make_synthetic(
    Nc: int,
    Ng: int,
    H: int = 3,
    Hw: int = 1,
    Nb: int = 1,
    Nx: int = 1,
    seed: int = 0,
    device="cpu",
    sorted_batches: bool = True,
    chunk_cells: int = 32768,
    pad_to: int = 4,
    stats: bool = True,
) -> SyntheticCycleData:
    """Draw one dataset.  ``seed`` should be ``base + rank`` under cell sharding."""
    device = torch.device(device)
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    f32 = dict(device=device, dtype=torch.float32)
    K, Kw = 2 * H + 1, 2 * Hw + 1
    ld = ((Ng + pad_to - 1) // pad_to) * pad_to

    def randn(*shape):
        return torch.randn(*shape, generator=g, **f32)

    nu = torch.cat([0.4 + 1.2 * randn(Ng, 1), 0.2 * randn(Ng, K - 1)], dim=1)
    loggamma = 0.5 * randn(Ng)
    logbeta = 2.0 + 1.0 * randn(Ng)
    # Gamma(1, 2) = Exponential(rate 2); keep away from 0 so that 1/shape_inv stays finite
    shape_inv = (-torch.log1p(-torch.rand(Ng, generator=g, **f32)) / 2.0).clamp_min(0.02)
    dnu = 0.1 * randn(Nb, Ng)
    nu_omega = torch.zeros(Nx, Kw, **f32)
    nu_omega[:, 0] = torch.tensor([0.4 - 0.1 * (x % 3) for x in range(Nx)], **f32)
    if Kw > 1:
        nu_omega[:, 1:] = 0.05 * randn(Nx, Kw - 1)

    phi = (torch.rand(Nc, generator=g, **f32) * 2.0 - 1.0) * math.pi
    cf = 0.2 * randn(Nc)
    cf = cf - cf.mean()
    if sorted_batches:  # concatenated samples: contiguous blocks, like anndata.concat
        batch_id = (torch.arange(Nc, device=device) * Nb // max(Nc, 1)).to(torch.int32)
        cond_id = (torch.arange(Nc, device=device) * Nx // max(Nc, 1)).to(torch.int32)
    else:
        batch_id = torch.randint(0, Nb, (Nc,), generator=g, device=device, dtype=torch.int32)
        cond_id = torch.randint(0, Nx, (Nc,), generator=g, device=device, dtype=torch.int32)

    S = torch.zeros(Nc, ld, **f32)
    U = torch.zeros(Nc, ld, **f32)
    r = 1.0 / shape_inv
    gamma = torch.exp(loggamma)
    zS = zU = 0
    for c0 in range(0, Nc, chunk_cells):
        c1 = min(Nc, c0 + chunk_cells)
        ph = phi[c0:c1]
        z0, z1 = fourier_rows(ph, H, 0), fourier_rows(ph, H, 1)
        omega = (fourier_rows(ph, Hw, 0) * nu_omega[cond_id[c0:c1].long()]).sum(-1)
        etaS = z0 @ nu.T + dnu[batch_id[c0:c1].long()] + cf[c0:c1, None]
        etaU = -logbeta + torch.log(torch.relu((z1 @ nu.T) * omega[:, None] + gamma) + 1e-5) + etaS
        for eta, out in ((etaS, S), (etaU, U)):
            mu = torch.exp(eta.clamp_max(8.0))
            lam = torch._standard_gamma(r.expand_as(mu).contiguous(), generator=g) * (mu / r)
            out[c0:c1, :Ng] = torch.poisson(lam, generator=g)
        if stats:
            zS += int((S[c0:c1, :Ng] == 0).sum())
            zU += int((U[c0:c1, :Ng] == 0).sum())
    data = SyntheticCycleData(
        S=S, U=U, Nc=Nc, Ng=Ng, ld=ld, H=H, Hw=Hw, Nb=Nb, Nx=Nx, phi=phi, cf=cf,
        batch_id=batch_id, cond_id=cond_id, nu=nu, dnu=dnu, shape_inv=shape_inv,
        logbeta=logbeta, loggamma=loggamma, nu_omega=nu_omega,
    )
    if stats and Nc * Ng > 0:
        data.zero_frac_S = zS / float(Nc * Ng)
        data.zero_frac_U = zU / float(Nc * Ng)
    return data
