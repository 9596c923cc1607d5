"""Distributions with Pyro's surface (``to_event``, ``mask``, ``__call__`` = (r)sample, ``expand``).

Normal / Gamma / Poisson / Uniform / Bernoulli / LowRankMultivariateNormal wrap ``torch.distributions`` exactly
as ``pyro.distributions`` does.  ``Delta`` and ``GammaPoisson`` are Pyro-only; they are restated from Pyro 1.8.6
(``pyro/distributions/delta.py``, ``pyro/distributions/conjugate.py``, ``pyro/ops/special.py:log_beta``).
"""
from __future__ import annotations

import numbers

import torch
import torch.distributions as td
from torch.distributions import constraints  # noqa: F401  (re-exported like pyro.distributions.constraints)
from torch.distributions.utils import broadcast_all

__all__ = [
    "constraints", "Normal", "Gamma", "Poisson", "Uniform", "Bernoulli", "LowRankMultivariateNormal",
    "MultivariateNormal", "Delta", "GammaPoisson", "Independent", "MaskedDistribution", "TorchDistributionMixin",
]


class TorchDistributionMixin:
    """The part of ``pyro.distributions.torch_distribution.TorchDistributionMixin`` the models use."""

    def __call__(self, sample_shape=torch.Size()):
        return self.rsample(sample_shape) if self.has_rsample else self.sample(sample_shape)

    @property
    def event_dim(self) -> int:
        return len(self.event_shape)

    def shape(self, sample_shape=torch.Size()):
        return torch.Size(sample_shape) + self.batch_shape + self.event_shape

    def to_event(self, reinterpreted_batch_ndims=None):
        if reinterpreted_batch_ndims is None:
            reinterpreted_batch_ndims = len(self.batch_shape)
        if reinterpreted_batch_ndims == 0:
            return self
        return Independent(self, reinterpreted_batch_ndims)

    def mask(self, mask):
        return MaskedDistribution(self, mask)

    def expand_by(self, sample_shape):
        return self.expand(torch.Size(sample_shape) + self.batch_shape)


def _wrap(name, base):
    cls = type(name, (base, TorchDistributionMixin), {"__doc__": f"``pyro.distributions.{name}`` (wraps torch)."})
    return cls


Normal = _wrap("Normal", td.Normal)
Gamma = _wrap("Gamma", td.Gamma)
Poisson = _wrap("Poisson", td.Poisson)
Uniform = _wrap("Uniform", td.Uniform)
Bernoulli = _wrap("Bernoulli", td.Bernoulli)
LowRankMultivariateNormal = _wrap("LowRankMultivariateNormal", td.LowRankMultivariateNormal)
MultivariateNormal = _wrap("MultivariateNormal", td.MultivariateNormal)


class Independent(td.Independent, TorchDistributionMixin):
    def expand(self, batch_shape, _instance=None):
        batch_shape = torch.Size(batch_shape)
        base = self.base_dist.expand(batch_shape + self.event_shape[: self.reinterpreted_batch_ndims])
        return Independent(base, self.reinterpreted_batch_ndims)


class MaskedDistribution(td.Distribution, TorchDistributionMixin):
    """``dist.mask(m)``: log_prob multiplied by a boolean mask (False -> contributes exactly zero)."""

    arg_constraints = {}

    def __init__(self, base_dist, mask):
        self.base_dist = base_dist
        self._mask = mask
        super().__init__(base_dist.batch_shape, base_dist.event_shape, validate_args=False)

    @property
    def has_rsample(self):
        return self.base_dist.has_rsample

    @property
    def support(self):
        return self.base_dist.support

    def expand(self, batch_shape, _instance=None):
        return MaskedDistribution(self.base_dist.expand(batch_shape), self._mask)

    def sample(self, sample_shape=torch.Size()):
        return self.base_dist.sample(sample_shape)

    def rsample(self, sample_shape=torch.Size()):
        return self.base_dist.rsample(sample_shape)

    def log_prob(self, value):
        if self._mask is False:
            shape = torch.broadcast_shapes(self.base_dist.batch_shape, value.shape[: value.dim() - len(self.event_shape)])
            return torch.zeros(shape, dtype=value.dtype if value.is_floating_point() else torch.float32,
                               device=value.device)
        lp = self.base_dist.log_prob(value)
        if self._mask is True:
            return lp
        return torch.where(self._mask, lp, torch.zeros_like(lp))


class Delta(td.Distribution, TorchDistributionMixin):
    """Point mass at ``v`` with optional log-density; ``event_dim`` rightmost dims of ``v`` are event dims."""

    has_rsample = True
    arg_constraints = {"v": constraints.dependent, "log_density": constraints.real}
    support = constraints.real

    def __init__(self, v, log_density=0.0, event_dim=0, validate_args=None):
        if event_dim > v.dim():
            raise ValueError(f"Expected event_dim <= v.dim(), actual {event_dim} vs {v.dim()}")
        batch_dim = v.dim() - event_dim
        batch_shape = v.shape[:batch_dim]
        event_shape = v.shape[batch_dim:]
        if isinstance(log_density, numbers.Number):
            log_density = torch.full(batch_shape, float(log_density), dtype=v.dtype, device=v.device)
        elif validate_args and log_density.shape != batch_shape:
            raise ValueError(f"Expected log_density.shape = {batch_shape}, actual {log_density.shape}")
        self.v = v
        self.log_density = log_density
        super().__init__(batch_shape, event_shape, validate_args=False)

    def expand(self, batch_shape, _instance=None):
        batch_shape = torch.Size(batch_shape)
        v = self.v.expand(batch_shape + self.event_shape)
        log_density = self.log_density.expand(batch_shape)
        return Delta(v, log_density, event_dim=len(self.event_shape))

    def rsample(self, sample_shape=torch.Size()):
        shape = torch.Size(sample_shape) + self.v.shape
        return self.v.expand(shape)

    sample = rsample

    def log_prob(self, x):
        v = self.v.expand(self.batch_shape + self.event_shape)
        log_prob = (x == v).type(x.dtype).log()
        if len(self.event_shape):
            log_prob = log_prob.reshape(log_prob.shape[: log_prob.dim() - len(self.event_shape)] + (-1,)).sum(-1)
        return log_prob + self.log_density

    @property
    def mean(self):
        return self.v

    @property
    def variance(self):
        return torch.zeros_like(self.v)


def log_beta(x, y):
    """Pyro's ``log_beta`` with the default ``tol=0`` (exact branch)."""
    return x.lgamma() + y.lgamma() - (x + y).lgamma()


class GammaPoisson(td.Distribution, TorchDistributionMixin):
    """Compound Gamma(concentration, rate)-Poisson = negative binomial, as ``pyro.distributions.GammaPoisson``.

    ``log_prob`` is the op chain the reference evaluates over the full (Ng,Nc) matrix
    (``phase_inference_model.py:393``, ``velocity_inference_model.py:385-386``)."""

    arg_constraints = {"concentration": constraints.positive, "rate": constraints.positive}
    support = constraints.nonnegative_integer

    def __init__(self, concentration, rate, validate_args=None):
        concentration, rate = broadcast_all(concentration, rate)
        self._gamma = td.Gamma(concentration, rate, validate_args=validate_args)
        super().__init__(self._gamma.batch_shape, validate_args=validate_args)

    @property
    def concentration(self):
        return self._gamma.concentration

    @property
    def rate(self):
        return self._gamma.rate

    def expand(self, batch_shape, _instance=None):
        batch_shape = torch.Size(batch_shape)
        return GammaPoisson(self.concentration.expand(batch_shape), self.rate.expand(batch_shape),
                            validate_args=False)

    def sample(self, sample_shape=torch.Size()):
        rate = self._gamma.sample(sample_shape)
        return torch.poisson(rate)

    def log_prob(self, value):
        if self._validate_args:
            self._validate_sample(value)
        post_value = self.concentration + value
        return (
            -log_beta(self.concentration, value + 1)
            - post_value.log()
            + self.concentration * self.rate.log()
            - post_value * (1 + self.rate).log()
        )

    @property
    def mean(self):
        return self.concentration / self.rate

    @property
    def variance(self):
        return self.concentration / self.rate.pow(2) * (1 + self.rate)
