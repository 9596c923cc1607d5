"""Cached scalar device constants.

``torch.tensor(1.0, device="cuda")`` and Python numbers passed to ``torch.distributions`` both copy a scalar from
pageable host memory, which is illegal while a CUDA graph is being captured (svi.GraphedSVI).  ``const`` builds the
scalar once with a fill kernel and reuses it."""
from __future__ import annotations

from typing import Dict, Tuple

import torch

_CACHE: Dict[Tuple[str, float], torch.Tensor] = {}


def const(value: float, device) -> torch.Tensor:
    key = (str(torch.device(device)), float(value))
    t = _CACHE.get(key)
    if t is None:
        t = torch.full((), float(value), dtype=torch.float32, device=device)
        _CACHE[key] = t
    return t
