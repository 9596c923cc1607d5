"""``pyro.poutine`` subset: Trace, trace, replay, condition, block (Pyro 1.8.6 semantics restated).

Conditioning semantics the fit drivers rely on (``velocity_inference_model.py:61-66``):
``condition(model, data)`` marks the named sites observed at the given values (their log-prob stays in the
ELBO); ``block(guide, hide=[names])`` hides those *sample sites* from outer handlers -- the guide body still
runs, so the RNG stream is consumed exactly as without conditioning.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, Dict, Iterable, Optional

import torch

from .primitives import Messenger, validation_enabled

__all__ = ["Trace", "trace", "replay", "condition", "block", "TraceMessenger"]


def _is_identically_zero(x) -> bool:
    return isinstance(x, (int, float)) and x == 0


class Trace:
    """Ordered record of the sites a program visited."""

    def __init__(self):
        self.nodes: "OrderedDict[str, dict]" = OrderedDict()

    def add_node(self, site_name: str, **site) -> None:
        if site_name in self.nodes:
            raise RuntimeError(f"Multiple sites named '{site_name}'")
        self.nodes[site_name] = site

    def __contains__(self, name):
        return name in self.nodes

    def __iter__(self):
        return iter(self.nodes)

    def stochastic_nodes(self):
        return [n for n, s in self.nodes.items() if s["type"] == "sample" and not s["is_observed"]]

    def observation_nodes(self):
        return [n for n, s in self.nodes.items() if s["type"] == "sample" and s["is_observed"]]

    def param_nodes(self):
        return [n for n, s in self.nodes.items() if s["type"] == "param"]

    def compute_log_prob(self, site_filter=lambda name, site: True) -> None:
        for name, site in self.nodes.items():
            if site["type"] != "sample" or not site_filter(name, site) or "log_prob" in site:
                continue
            if site["infer"].get("_deterministic"):
                zero = torch.zeros((), device=site["value"].device)
                site["log_prob"], site["log_prob_sum"] = zero, zero
                continue
            log_p = site["fn"].log_prob(site["value"], *site["args"], **site["kwargs"])
            if site["mask"] is False:
                log_p = torch.zeros_like(log_p)
            elif site["mask"] is not None and site["mask"] is not True:
                log_p = torch.where(site["mask"], log_p, torch.zeros_like(log_p))
            if site["scale"] != 1.0:
                log_p = log_p * site["scale"]
            site["log_prob"] = log_p
            site["log_prob_sum"] = log_p.sum()
            if validation_enabled():
                _check_site_shape(name, site)

    def compute_score_parts(self) -> None:
        """All guide sites of this package are reparameterised (Normal, Delta, Gamma): the score-function term
        is zero and the entropy term is the log-prob (``Trace_ELBO`` then needs nothing else)."""
        self.compute_log_prob()
        for name, site in self.nodes.items():
            if site["type"] == "sample" and "score_parts" not in site:
                if not getattr(site["fn"], "has_rsample", False) and not site["is_observed"]:
                    raise NotImplementedError(f"non-reparameterised guide site '{name}' is not supported")
                site["score_parts"] = (site["log_prob"], 0, site["log_prob"])

    def log_prob_sum(self):
        self.compute_log_prob()
        return sum(s["log_prob_sum"] for s in self.nodes.values() if s["type"] == "sample")

    def format_shapes(self, title="Trace Shapes:") -> str:
        """Same information as Pyro's ``Trace.format_shapes``: per site the dist batch | event shape and the
        value shape (the only 'golden output' the reference records: SURVEY Appendix B)."""
        rows = [[title]]
        for name, site in self.nodes.items():
            if site["type"] == "param":
                rows.append(["Param Sites:"]) if ["Param Sites:"] not in rows else None
                rows.append([name, *[str(s) for s in site["value"].shape]])
        rows.append(["Sample Sites:"])
        for name, site in self.nodes.items():
            if site["type"] != "sample":
                continue
            fn = site["fn"]
            b, e = tuple(getattr(fn, "batch_shape", ())), tuple(getattr(fn, "event_shape", ()))
            rows.append([name + " dist", *map(str, b), "|", *map(str, e)])
            v = site["value"]
            vs = tuple(v.shape) if hasattr(v, "shape") else ()
            ed = len(e)
            rows.append(["value", *map(str, vs[: len(vs) - ed]), "|", *map(str, vs[len(vs) - ed:])])
        return "\n".join(" ".join(r) for r in rows)

    def site_shapes(self) -> Dict[str, tuple]:
        out = {}
        for name, site in self.nodes.items():
            if site["type"] == "sample":
                fn = site["fn"]
                out[name] = (tuple(getattr(fn, "batch_shape", ())), tuple(getattr(fn, "event_shape", ())),
                             tuple(site["value"].shape))
        return out


def _check_site_shape(name, site) -> None:
    """log_prob must broadcast against the enclosing plates (Pyro's check_site_shape)."""
    shape = list(site["log_prob"].shape)
    for f in site["cond_indep_stack"]:
        if f.dim is None:
            continue
        k = -f.dim
        if len(shape) < k:
            continue  # broadcastable
        if shape[f.dim] not in (1, f.size):
            raise ValueError(
                f"at site '{name}', log_prob shape {tuple(shape)} does not match plate '{f.name}' "
                f"(dim {f.dim}, size {f.size})"
            )


class TraceMessenger(Messenger):
    def __init__(self, fn=None, graph_type="flat", param_only=False):
        super().__init__(fn)
        self.param_only = param_only
        self.trace = Trace()

    def __enter__(self):
        self.trace = Trace()
        return super().__enter__()

    def get_trace(self, *args, **kwargs) -> Trace:
        self(*args, **kwargs)
        return self.trace

    def _pyro_post_sample(self, msg):
        if self.param_only:
            return
        self.trace.add_node(msg["name"], **{k: v for k, v in msg.items()})

    def _pyro_post_param(self, msg):
        if msg["name"] not in self.trace.nodes:
            self.trace.add_node(msg["name"], **{k: v for k, v in msg.items()})


def trace(fn: Optional[Callable] = None, graph_type="flat", param_only=False) -> TraceMessenger:
    return TraceMessenger(fn, graph_type=graph_type, param_only=param_only)


class ReplayMessenger(Messenger):
    def __init__(self, fn=None, trace: Optional[Trace] = None, params=None):
        super().__init__(fn)
        self.guide_trace = trace

    def _pyro_sample(self, msg):
        name = msg["name"]
        if self.guide_trace is not None and name in self.guide_trace.nodes:
            guide_msg = self.guide_trace.nodes[name]
            if msg["is_observed"]:
                return
            if guide_msg["type"] != "sample" or guide_msg["is_observed"]:
                raise RuntimeError(f"site {name} must be a latent sample site in the replayed trace")
            msg["done"] = True
            msg["value"] = guide_msg["value"]
            msg["infer"] = guide_msg["infer"]


def replay(fn=None, trace=None, params=None) -> ReplayMessenger:
    return ReplayMessenger(fn, trace=trace, params=params)


class ConditionMessenger(Messenger):
    def __init__(self, fn=None, data: Optional[dict] = None):
        super().__init__(fn)
        self.data = data or {}

    def _pyro_sample(self, msg):
        name = msg["name"]
        if name in self.data:
            if msg["is_observed"] and not msg["infer"].get("_deterministic"):
                raise RuntimeError(f"cannot condition on already observed site '{name}'")
            msg["value"] = self.data[name]
            msg["is_observed"] = True


def condition(fn=None, data=None) -> ConditionMessenger:
    return ConditionMessenger(fn, data=data)


class BlockMessenger(Messenger):
    def __init__(self, fn=None, hide_fn=None, expose_fn=None, hide_all=True, hide: Optional[Iterable[str]] = None,
                 expose: Optional[Iterable[str]] = None, hide_types=None, expose_types=None):
        super().__init__(fn)
        hide = None if hide is None else list(hide)
        self.hide = None if hide is None else frozenset(hide)  # (read by the fused step to recognise block(guide, hide=sites))
        self.plain_hide = hide is not None and hide_fn is None and expose_fn is None and expose is None and not hide_types \
            and not expose_types
        if hide_fn is not None:
            self.hide_fn = hide_fn
        elif expose_fn is not None:
            self.hide_fn = lambda msg: not expose_fn(msg)
        elif hide is not None or hide_types is not None:
            hide = set(hide or ())
            hide_types = set(hide_types or ())
            self.hide_fn = lambda msg: msg["name"] in hide or msg["type"] in hide_types
        elif expose is not None or expose_types is not None:
            expose = set(expose or ())
            expose_types = set(expose_types or ())
            self.hide_fn = lambda msg: not (msg["name"] in expose or msg["type"] in expose_types)
        else:
            self.hide_fn = lambda msg: True

    def _process_message(self, msg):
        msg["stop"] = bool(self.hide_fn(msg))


def block(fn=None, **kwargs) -> BlockMessenger:
    return BlockMessenger(fn, **kwargs)
