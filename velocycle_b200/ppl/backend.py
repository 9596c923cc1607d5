"""Which probabilistic-programming runtime the model/guide modules bind to."""
from __future__ import annotations

import importlib


def use_real_pyro():
    """Return the real ``pyro`` module if it is importable (it is not in this image), else None."""
    try:
        return importlib.import_module("pyro")
    except Exception:
        return None


def get():
    """(pyro, dist, poutine, infer, optim): real Pyro when present, otherwise velocycle_b200.ppl."""
    real = use_real_pyro()
    if real is not None:  # pragma: no cover - Pyro is absent from the build image
        import pyro.distributions as dist
        from pyro import infer, optim, poutine

        return real, dist, poutine, infer, optim
    from velocycle_b200 import ppl
    from velocycle_b200.ppl import distributions as dist
    from velocycle_b200.ppl import infer, optim, poutine

    return ppl, dist, poutine, infer, optim
