"""``pyro.infer`` subset: Trace_ELBO (1..n particles, reparameterised sites), SVI, Predictive.

Restated from Pyro 1.8.6 (``pyro/infer/trace_elbo.py``, ``pyro/infer/svi.py``, ``pyro/infer/predictive.py``,
``pyro/infer/elbo.py``): SVI.step = trace the params, ``loss_and_grads``, one optimizer per parameter tensor,
zero the grads, return the loss as a Python float.
"""
from __future__ import annotations

import warnings
from typing import Callable, Dict, Iterable, Optional

import torch

from . import poutine
from .primitives import get_param_store, validation_enabled

__all__ = ["Trace_ELBO", "TraceEnum_ELBO", "SVI", "Predictive", "config_enumerate", "autoguide"]


def torch_item(x):
    return x if isinstance(x, (int, float)) else x.item()


def _check_model_guide_match(model_trace, guide_trace) -> None:
    guide_vars = set(guide_trace.stochastic_nodes())
    model_vars = set(n for n in model_trace.stochastic_nodes()) | set(
        n for n, s in model_trace.nodes.items() if s["type"] == "sample" and s["is_observed"]
    )
    missing = guide_vars - set(model_trace.nodes)
    if missing:
        warnings.warn(f"Found vars in guide but not model: {missing}")
    unsampled = set(model_trace.stochastic_nodes()) - guide_vars
    if unsampled:
        warnings.warn(f"Found non-auxiliary vars in model but not guide: {unsampled}")


class Trace_ELBO:
    def __init__(self, num_particles: int = 1, max_plate_nesting=float("inf"), vectorize_particles: bool = False,
                 retain_graph: Optional[bool] = None, **_):
        self.num_particles = num_particles
        self.retain_graph = retain_graph
        if vectorize_particles:
            raise NotImplementedError("vectorize_particles is not supported")

    def _get_trace(self, model, guide, args, kwargs):
        guide_trace = poutine.trace(guide).get_trace(*args, **kwargs)
        model_trace = poutine.trace(poutine.replay(model, trace=guide_trace)).get_trace(*args, **kwargs)
        if validation_enabled():
            _check_model_guide_match(model_trace, guide_trace)
        model_trace.compute_log_prob()
        guide_trace.compute_score_parts()
        return model_trace, guide_trace

    def _particle(self, model, guide, args, kwargs):
        model_trace, guide_trace = self._get_trace(model, guide, args, kwargs)
        elbo_particle = 0.0
        surrogate = 0.0
        for name, site in model_trace.nodes.items():
            if site["type"] == "sample":
                elbo_particle = elbo_particle + torch_item(site["log_prob_sum"])
                surrogate = surrogate + site["log_prob_sum"]
        for name, site in guide_trace.nodes.items():
            if site["type"] == "sample":
                _, _, entropy_term = site["score_parts"]
                elbo_particle = elbo_particle - torch_item(site["log_prob_sum"])
                if not isinstance(entropy_term, (int, float)):
                    surrogate = surrogate - entropy_term.sum()
        return -elbo_particle, -surrogate, model_trace, guide_trace

    def loss(self, model, guide, *args, **kwargs) -> float:
        elbo = 0.0
        with torch.no_grad():
            for _ in range(self.num_particles):
                loss_particle, _, _, _ = self._particle(model, guide, args, kwargs)
                elbo += loss_particle / self.num_particles
        if elbo != elbo:
            warnings.warn("Encountered NaN: loss")
        return elbo

    def differentiable_loss(self, model, guide, *args, **kwargs):
        total = 0.0
        for _ in range(self.num_particles):
            _, surrogate, _, _ = self._particle(model, guide, args, kwargs)
            total = total + surrogate / self.num_particles
        return total

    def loss_and_grads(self, model, guide, *args, **kwargs) -> float:
        loss = 0.0
        for _ in range(self.num_particles):
            loss_particle, surrogate, model_trace, guide_trace = self._particle(model, guide, args, kwargs)
            loss += loss_particle / self.num_particles
            trainable = any(s["type"] == "param" for tr in (model_trace, guide_trace) for s in tr.nodes.values())
            if trainable and getattr(surrogate, "requires_grad", False):
                (surrogate / self.num_particles).backward(retain_graph=self.retain_graph)
        if loss != loss:
            warnings.warn("Encountered NaN: loss")
        return loss


class TraceEnum_ELBO(Trace_ELBO):
    """Placeholder: parallel enumeration is only needed by the reference's unreachable LBA model."""

    def _get_trace(self, *a, **k):  # pragma: no cover
        raise NotImplementedError("TraceEnum_ELBO (discrete enumeration) is out of scope: SURVEY.md section 2")


def config_enumerate(fn=None, default="parallel", **_):
    if fn is None:
        return lambda f: f
    return fn


class SVI:
    def __init__(self, model, guide, optim, loss, loss_and_grads=None, **_):
        self.model, self.guide, self.optim = model, guide, optim
        if isinstance(loss, Trace_ELBO):
            self.loss, self.loss_and_grads = loss.loss, loss.loss_and_grads
        else:
            self.loss = loss
            self.loss_and_grads = loss_and_grads
            if loss_and_grads is None:
                def _loss_and_grads(model, guide, *args, **kwargs):
                    val = loss(model, guide, *args, **kwargs)
                    if getattr(val, "requires_grad", False):
                        val.backward(retain_graph=True)
                    return val
                self.loss_and_grads = _loss_and_grads

    def evaluate_loss(self, *args, **kwargs) -> float:
        with torch.no_grad():
            return torch_item(self.loss(self.model, self.guide, *args, **kwargs))

    def step(self, *args, **kwargs) -> float:
        with poutine.trace(param_only=True) as param_capture:
            loss = self.loss_and_grads(self.model, self.guide, *args, **kwargs)
        store = get_param_store()
        params = []
        seen = set()
        for site in param_capture.trace.nodes.values():
            p = store.get_unconstrained(site["name"])
            if id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        self.optim(params)
        for p in params:  # pyro.infer.util.zero_grads
            p.grad = None
        return torch_item(loss)


def _batch_ndim(site: dict) -> int:
    """Length of the site distribution's batch shape as Pyro's BroadcastMessenger leaves it: plates expand every sample
    site -- ``pyro.deterministic`` sites included, whose Delta starts with an empty batch shape -- up to their dim."""
    n = len(getattr(site["fn"], "batch_shape", ()))
    for frame in site["cond_indep_stack"]:
        n = max(n, -frame.dim)
    return n


def predictive_site_shapes(model_trace, num_samples: int, return_sites, posterior_names=()) -> Dict[str, tuple]:
    """Which sites ``Predictive`` returns and the shape of each: Pyro 1.8.6 ``pyro/infer/predictive.py`` (``_predictive``):
    ``(num_samples,) + (1,) * (max_plate_nesting - len(fn.batch_shape)) + value.shape`` for the sites in ``return_sites``;
    every sample site (latent, observed, deterministic) when ``return_sites`` is None; the sites that are not among the
    posterior samples when it is empty."""
    sites = [(k, v) for k, v in model_trace.nodes.items() if v["type"] == "sample"]
    dims = [frame.dim for _, v in sites for frame in v["cond_indep_stack"]]
    mpn = -min(dims) if dims else 0
    shapes: Dict[str, tuple] = {}
    for name, site in sites:
        shape = (int(num_samples),) + (1,) * (mpn - _batch_ndim(site)) + tuple(site["value"].shape)
        if return_sites:
            if name in return_sites:
                shapes[name] = shape
        elif return_sites is None or name not in posterior_names:
            shapes[name] = shape
    return shapes


def predictive_padding(model, guide, *args, device=None, **kwargs) -> Dict[str, int]:
    """site -> number of singleton dims ``Predictive`` inserts behind the sample dim (``predictive_site_shapes``), from
    one guide trace and model replay that leave the RNG state untouched."""
    dev = None if device is None else torch.device(device)
    devices = [dev.index if dev.index is not None else torch.cuda.current_device()] if dev is not None and dev.type == "cuda" else []
    with torch.random.fork_rng(devices=devices), torch.no_grad():
        guide_trace = poutine.trace(guide).get_trace(*args, **kwargs)
        model_trace = poutine.trace(poutine.replay(model, trace=guide_trace)).get_trace(*args, **kwargs)
    shapes = predictive_site_shapes(model_trace, 1, None)
    return {k: len(shape) - 1 - model_trace.nodes[k]["value"].dim() for k, shape in shapes.items()}


class Predictive:
    """Sequential ``Predictive(model, guide=guide, num_samples=N, return_sites=...)`` as used by the fit drivers
    (``velocity_inference_model.py:279-291``): draw the guide, replay the model, collect the requested sites.  Site selection
    and output shapes follow Pyro 1.8.6 (``predictive_site_shapes``): with a guide and no ``return_sites`` every model site
    comes back, and every value is left-padded with singleton dims up to the model's plate nesting.  (Pyro spends two guide
    and two model executions on shape discovery before the draws; that RNG consumption is not reproduced.)"""

    def __init__(self, model, posterior_samples=None, guide=None, num_samples=None, return_sites=(), parallel=False):
        if parallel:
            raise NotImplementedError("parallel Predictive is not supported (the reference uses the sequential mode)")
        if num_samples is None and posterior_samples is None:
            raise ValueError("num_samples or posterior_samples is required")
        self.model, self.guide = model, guide
        self.posterior_samples = posterior_samples
        self.num_samples = num_samples
        self.return_sites = None if return_sites is None else tuple(return_sites)

    @torch.no_grad()
    def __call__(self, *args, **kwargs) -> Dict[str, torch.Tensor]:
        n = self.num_samples
        if n is None:
            n = next(iter(self.posterior_samples.values())).shape[0]
        return_sites = self.return_sites
        if self.guide is not None and not return_sites:
            return_sites = None  # "return all sites by default if a guide is provided"
        shapes = None
        collected: Dict[str, list] = {}
        for i in range(n):
            if self.guide is not None:
                guide_trace = poutine.trace(self.guide).get_trace(*args, **kwargs)
                model_trace = poutine.trace(poutine.replay(self.model, trace=guide_trace)).get_trace(*args, **kwargs)
                names = [k for k, v in guide_trace.nodes.items() if v["type"] == "sample"]
            else:
                data = {k: v[i] for k, v in (self.posterior_samples or {}).items()}
                model_trace = poutine.trace(poutine.condition(self.model, data=data)).get_trace(*args, **kwargs)
                names = list(data)
            if shapes is None:
                shapes = predictive_site_shapes(model_trace, n, return_sites, names)
            for name in shapes:
                collected.setdefault(name, []).append(model_trace.nodes[name]["value"].detach())
        return {k: torch.stack(v).reshape(shapes[k]) for k, v in collected.items()}

    forward = __call__


class _AutoGuideStub:
    def __init__(self, *a, **k):  # pragma: no cover
        raise NotImplementedError("autoguides are not part of the VeloCycle hot path (guides are hand-written)")


class _AutoGuideNamespace:
    AutoNormal = AutoDiagonalNormal = AutoDelta = AutoGuideList = _AutoGuideStub

    @staticmethod
    def init_to_mean(site=None, *, fallback=None):
        return lambda site: None

    @staticmethod
    def init_to_median(site=None, *, num_samples=15, fallback=None):
        return lambda site: None


autoguide = _AutoGuideNamespace()
