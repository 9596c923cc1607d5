"""The fused likelihood as seen from a model function: packed-count sidecar lookup and the site distribution.

``pyro.sample("S", FusedCountLikelihood(lp_S, ...), obs=mp.S)`` replaces
``pyro.sample("S", dist.GammaPoisson(1/shape_inv, 1/(shape_inv*exp(ElogS))), obs=mp.S)``
(``phase_inference_model.py:391-393``, ``velocity_inference_model.py:383-386``): the site keeps its name and its
observed value, but its ``log_prob`` is the per-gene sum over cells that the CUDA pass already produced, shape
(Ng,1) inside the ``genes`` plate -- the (Ng,Nc) matrix of log-probs is never materialised (SURVEY 8b option i).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch.distributions import constraints

from . import _lib
from .fused import PackedCounts

__all__ = ["FusedCountLikelihood", "packed_counts_for", "attach_packed_counts", "without_count_sites", "count_sites_enabled"]

_COUNT_SITES = True


class without_count_sites:
    """Context manager: the model functions of this package skip their observed count sites ("S", "U") -- the fused
    likelihood pass over the count matrices -- while it is active.  The fit drivers use it for posterior draws
    (``Predictive`` with ``return_sites`` that name latent / deterministic sites only, ``velocity_inference_model.py:213-224``):
    the reference re-evaluates both GammaPoisson sites over the (Ng, Nc) matrices for each of the 500 draws although nobody
    reads them; here a draw touches no count byte."""

    def __enter__(self):
        global _COUNT_SITES
        self._prev, _COUNT_SITES = _COUNT_SITES, False
        return self

    def __exit__(self, *exc):
        global _COUNT_SITES
        _COUNT_SITES = self._prev
        return False


def count_sites_enabled() -> bool:
    return _COUNT_SITES

_ATTR = "_vcb_packed_counts"  # the sidecar rides on the count TENSOR: it lives and dies with mp.S


def _design_key(mp, need_U: bool) -> Tuple:
    """What the packed counts were derived from besides ``mp.S``: the unspliced matrix and the design tensors (identity and
    in-place version), so that re-preprocessing the same matrix with another design, or freeing and re-allocating a
    dataset at the same address, can never hand back stale counts / batch ids."""
    def ident(t):
        return None if t is None else (id(t), t.data_ptr(), getattr(t, "_version", 0), tuple(t.shape))

    return (ident(getattr(mp, "U", None)) if need_U else None, ident(getattr(mp, "Db", None)),
            ident(getattr(mp, "D", None)) if need_U else None)


def attach_packed_counts(mp, counts: PackedCounts) -> None:
    """Register ready-made packed counts for ``mp`` (what ``preprocess_for_*`` of this package does)."""
    setattr(mp.S, _ATTR, {True: (_design_key(mp, True), counts), False: (_design_key(mp, False), counts)})


def packed_counts_for(mp, need_U: bool) -> PackedCounts:
    """Packed, device-resident counts for a metaparameter tuple; built once per dataset and kept ON ``mp.S`` (an attribute of
    the tensor object: no global table, nothing outlives the data).

    Works for ``mp`` objects made by the reference's own preprocessing: ``mp.S`` / ``mp.U`` logical (Ng,Nc)
    float tensors, one-hot ``mp.Db`` and ``mp.D``."""
    pc = getattr(mp, "packed_counts", None)
    if pc is not None:
        return pc
    cache = getattr(mp.S, _ATTR, None)
    if cache is None:
        cache = {}
        setattr(mp.S, _ATTR, cache)
    key = _design_key(mp, need_U)
    hit = cache.get(need_U)
    if hit is not None and hit[0] == key:
        return hit[1]
    if not need_U:  # counts packed together with U serve the phase model as well (same S, same batch design)
        hit = cache.get(True)
        if hit is not None and hit[0][1] == key[1]:
            return hit[1]
    if not mp.S.is_cuda:
        raise _lib.VcbError(
            "velocycle_b200 models need mp.S / mp.U on a CUDA device: there is no CPU path "
            "(pass device=torch.device('cuda') to preprocess_for_*_estimation)"
        )
    pc = PackedCounts.from_model_tensors(
        mp.S, mp.U if need_U else None, getattr(mp, "Db", None), getattr(mp, "D", None) if need_U else None
    )
    cache[need_U] = (key, pc)
    return pc


class FusedCountLikelihood(torch.distributions.Distribution):
    """Site distribution whose ``log_prob`` is the precomputed per-gene sum over cells of NB log-pmfs."""

    arg_constraints: dict = {}
    support = constraints.nonnegative_integer
    has_rsample = False

    def __init__(self, log_prob_per_gene: torch.Tensor, name: str = "S"):
        self._lp = log_prob_per_gene.reshape(-1, 1)
        self._name = name
        super().__init__(batch_shape=self._lp.shape, event_shape=torch.Size(), validate_args=False)

    def expand(self, batch_shape, _instance=None):
        if tuple(batch_shape) != tuple(self.batch_shape):
            raise ValueError(f"FusedCountLikelihood cannot be expanded to {tuple(batch_shape)}")
        return self

    def log_prob(self, value):
        return self._lp

    def sample(self, sample_shape=torch.Size()):
        raise NotImplementedError(
            f"site '{self._name}' is an observed-count likelihood; sampling counts from the fused site is not "
            "supported (condition it on data as the fit drivers do)"
        )

    def __call__(self, *a, **k):
        return self.sample(*a, **k)
