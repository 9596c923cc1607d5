"""Mean-field guide of the phase (manifold-learning) model.

Same parameter names, shapes, constraints and -- crucially -- the same order of RNG draws as
``velocycle/phase_inference_guide.py:10-56``: nu ~ Normal(nu_locs, nu_scales) first, then the Delta sites
(shape_inv, Delta-nu: no RNG), then phixy ~ Normal(phixy_locs, 1).  The guide is host-side Python on purpose:
it only touches O(Ng K + Nc) numbers.
"""
from __future__ import annotations

import torch

from ._const import const
from .ppl import backend

__all__ = ["phase_latent_variable_guide"]


def _phixy_guide_dist(dist, mp, locs, scale):
    """Normal(phixy_locs, 1) over the cells of this rank.  Under cell sharding the noise is the rank's slice of
    the global draw, so that all ranks consume the RNG stream like the single-GPU run (sharding.ShardedNormal)."""
    shard = getattr(mp, "shard", None)
    if shard is not None and shard.world > 1:
        from .sharding import ShardedNormal

        return ShardedNormal(locs, scale, shard, event_dims=1)
    return dist.Normal(locs, scale).to_event(1)


def phase_latent_variable_guide(mp):
    pyro, dist, _, _, _ = backend.get()
    dev = mp.device
    positive = dist.constraints.positive
    cells = pyro.plate("cells", mp.Nc, dim=-1, device=dev)
    genes = pyro.plate("genes", mp.Ng, dim=-2, device=dev)
    batches = pyro.plate("batches", mp.Nb, dim=-3, device=dev)

    nu_locs = pyro.param("ν_locs", mp.μνg.detach().clone().to(dev))
    nu_scales = pyro.param("ν_scales", mp.σνg.detach().clone().to(dev), constraint=positive)
    if mp.with_delta_nu:
        dnu_locs = pyro.param("Δν_locs", torch.ones((mp.Nb, mp.Ng, 1), device=dev) * mp.μΔν)
    phixy_locs = pyro.param("ϕxy_locs", mp.φxy_prior.detach().clone().to(dev))
    nb = mp.noisemodel == "NegativeBinomial"
    if nb:
        shape_inv_locs = pyro.param(
            "shape_inv_locs", torch.ones((mp.Ng, 1), device=dev) * mp.gamma_alpha / mp.gamma_beta, constraint=positive
        )

    with genes:
        pyro.sample("ν", dist.Normal(nu_locs, nu_scales).to_event(1))
        if nb:
            pyro.sample("shape_inv", dist.Delta(shape_inv_locs))
        if mp.with_delta_nu:
            with batches:
                pyro.sample("Δν", dist.Delta(dnu_locs))
    with cells:
        pyro.sample("ϕxy", _phixy_guide_dist(dist, mp, phixy_locs, const(1.0, dev)))
