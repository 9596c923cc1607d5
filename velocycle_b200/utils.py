"""Numeric helpers with the reference's names and conventions (``velocycle/utils.py:400-506``).

Column order of the Fourier basis is [1, sin phi, cos phi, sin 2phi, cos 2phi, ...]; ``der=1`` is the phase
derivative [0, cos phi, -sin phi, 2 cos 2phi, -2 sin 2phi, ...].  ``der=2`` is an addition used by tests.
These run in plain torch: they only ever touch per-cell (Nc x K) tensors.  The (Ng,Nc) work is in the kernels.
"""
from __future__ import annotations

import torch

__all__ = ["torch_fourier_basis", "torch_basis", "pack_direction", "unpack_direction", "circular_corrcoef"]


def torch_fourier_basis(ϕ: torch.Tensor, num_harmonics: int, der: int = 0, device=None) -> torch.Tensor:
    H = int(num_harmonics)
    if der not in (0, 1, 2):
        raise ValueError(f"Value {der=} is not allowed, use 0 or 1 instead")
    ϕ = ϕ if device is None else ϕ.to(device)
    lead = torch.ones_like(ϕ) if der == 0 else torch.zeros_like(ϕ)
    if H == 0:
        return lead.unsqueeze(-1)
    n = torch.arange(1, H + 1, device=ϕ.device, dtype=ϕ.dtype)
    arg = ϕ.unsqueeze(-1) * n
    s, c = torch.sin(arg), torch.cos(arg)
    if der == 0:
        pair = (s, c)
    elif der == 1:
        pair = (n * c, -n * s)
    else:
        pair = (-n * n * s, -n * n * c)
    inter = torch.stack(pair, dim=-1).reshape(*ϕ.shape, 2 * H)
    return torch.cat([lead.unsqueeze(-1), inter], dim=-1)


def torch_basis(x: torch.Tensor, der: int = 0, kind: str = "fourier", device=None, **kwargs) -> torch.Tensor:
    if kind != "fourier":
        raise ValueError(f"{kind=} is not a valid entry use `fourier`")
    if "num_harmonics" not in kwargs:
        raise ValueError("num_harmonics needs to be provided if kind=`fourier`")
    return torch_fourier_basis(x, num_harmonics=kwargs["num_harmonics"], der=der, device=device)


def unpack_direction(loc: torch.Tensor, concentration: float = 1.0) -> torch.Tensor:
    return torch.stack([torch.cos(loc), torch.sin(loc)], dim=-1) * concentration


def pack_direction(xy_pair: torch.Tensor) -> torch.Tensor:
    return torch.atan2(xy_pair[..., 1], xy_pair[..., 0])


def circular_corrcoef(x1, x2) -> float:
    """|mean(exp(i (x1 - x2)))|: 1 when two sets of angles agree up to a common rotation (``utils.py:586-611``); the figure
    of merit for inferred phases against ground truth."""
    import numpy as np

    x1, x2 = np.asarray(x1, dtype=np.float64), np.asarray(x2, dtype=np.float64)
    assert len(x1) == len(x2), "Input arrays must have the same length"
    return float(np.abs(np.mean(np.exp(1j * (x1 - x2)))))
