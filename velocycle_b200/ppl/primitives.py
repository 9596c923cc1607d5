"""Effect-handler core: message stack, ``sample`` / ``param`` / ``deterministic`` / ``plate`` and the
parameter store.  Semantics follow Pyro 1.8.6 (``pyro/primitives.py``, ``pyro/poutine/runtime.py``,
``pyro/poutine/plate_messenger.py``, ``pyro/poutine/broadcast_messenger.py``, ``pyro/params/param_store.py``),
restated -- Pyro itself is not installable in this image.
"""
from __future__ import annotations

import weakref
from collections import OrderedDict, namedtuple
from typing import Any, Callable, Dict, Optional

import torch
from torch.distributions import constraints, transform_to

_PYRO_STACK: list = []
_VALIDATION = True


def enable_validation(is_validate: bool = True) -> None:
    global _VALIDATION
    _VALIDATION = bool(is_validate)


def validation_enabled() -> bool:
    return _VALIDATION


def set_rng_seed(seed: int) -> None:
    torch.manual_seed(seed)
    try:
        import numpy as np
        import random

        random.seed(seed)
        np.random.seed(seed % (2**32))
    except Exception:  # pragma: no cover
        pass


# ------------------------------------------------------------------------------------------------------
# Parameter store
# ------------------------------------------------------------------------------------------------------
class ParamStoreDict:
    """name -> unconstrained leaf tensor (+ constraint); ``pyro.param`` returns the constrained view."""

    def __init__(self):
        self._params: Dict[str, torch.Tensor] = OrderedDict()
        self._constraints: Dict[str, Any] = {}

    def clear(self) -> None:
        self._params.clear()
        self._constraints.clear()

    def __contains__(self, name: str) -> bool:
        return name in self._params

    def __len__(self) -> int:
        return len(self._params)

    def keys(self):
        return self._params.keys()

    def items(self):
        for name in self._params:
            yield name, self[name]

    def named_parameters(self):
        return self._params.items()

    def setdefault(self, name: str, init, constraint=constraints.real) -> torch.Tensor:
        if name not in self._params:
            if callable(init) and not isinstance(init, torch.Tensor):
                init = init()
            self.__setitem__(name, init, constraint)
        return self[name]

    def __setitem__(self, name: str, value: torch.Tensor, constraint=constraints.real) -> None:
        with torch.no_grad():
            unconstrained = transform_to(constraint).inv(value.detach()).clone().contiguous()
        unconstrained.requires_grad_(True)
        self._params[name] = unconstrained
        self._constraints[name] = constraint

    def set_unconstrained(self, name: str, unconstrained: torch.Tensor, constraint=constraints.real) -> None:
        """Adopt an existing leaf tensor (used by the fused step to keep every parameter in one flat buffer)."""
        self._params[name] = unconstrained
        self._constraints[name] = constraint

    def __getitem__(self, name: str) -> torch.Tensor:
        unconstrained = self._params[name]
        constraint = self._constraints[name]
        if constraint is constraints.real or constraint == constraints.real:
            constrained = unconstrained
        else:
            constrained = transform_to(constraint)(unconstrained)
        try:
            constrained.unconstrained = weakref.ref(unconstrained)
        except Exception:  # pragma: no cover
            pass
        return constrained

    def get_unconstrained(self, name: str) -> torch.Tensor:
        return self._params[name]

    def get_constraint(self, name: str):
        return self._constraints[name]

    def get_state(self) -> dict:
        return {"params": {k: v.detach().clone() for k, v in self._params.items()},
                "constraints": dict(self._constraints)}

    def set_state(self, state: dict) -> None:
        self.clear()
        for k, v in state["params"].items():
            t = v.detach().clone().requires_grad_(True)
            self._params[k] = t
            self._constraints[k] = state["constraints"][k]


_PARAM_STORE = ParamStoreDict()


def get_param_store() -> ParamStoreDict:
    return _PARAM_STORE


def clear_param_store() -> None:
    _PARAM_STORE.clear()


# ------------------------------------------------------------------------------------------------------
# Messengers
# ------------------------------------------------------------------------------------------------------
class Messenger:
    """Context manager that sits on the handler stack; can also wrap a callable."""

    def __init__(self, fn: Optional[Callable] = None):
        self.fn = fn

    def __enter__(self):
        _PYRO_STACK.append(self)
        return self

    def __exit__(self, exc_type, exc, tb):
        if _PYRO_STACK and _PYRO_STACK[-1] is self:
            _PYRO_STACK.pop()
        elif self in _PYRO_STACK:  # exception unwinding: drop this frame and everything above it
            loc = _PYRO_STACK.index(self)
            del _PYRO_STACK[loc:]
        return False

    def __call__(self, *args, **kwargs):
        if self.fn is None:
            raise TypeError("Messenger used as a function without a wrapped callable")
        with self:
            return self.fn(*args, **kwargs)

    def _process_message(self, msg: dict) -> None:
        method = getattr(self, "_pyro_" + msg["type"], None)
        if method is not None:
            method(msg)

    def _postprocess_message(self, msg: dict) -> None:
        method = getattr(self, "_pyro_post_" + msg["type"], None)
        if method is not None:
            method(msg)


def apply_stack(msg: dict) -> dict:
    """Top of the stack (innermost handler) first; a handler may set ``stop`` to hide the site from the rest."""
    pointer = 0
    for pointer, frame in enumerate(reversed(_PYRO_STACK)):
        frame._process_message(msg)
        if msg["stop"]:
            break
    if msg["value"] is None and not msg["done"]:
        msg["value"] = msg["fn"](*msg["args"], **msg["kwargs"])
    msg["done"] = True
    for frame in _PYRO_STACK[len(_PYRO_STACK) - pointer - 1:] if _PYRO_STACK else []:
        frame._postprocess_message(msg)
    return msg


def _new_msg(type_: str, name: str, fn, args=(), kwargs=None, value=None, is_observed=False, infer=None) -> dict:
    return {
        "type": type_, "name": name, "fn": fn, "is_observed": is_observed, "args": args,
        "kwargs": kwargs or {}, "value": value, "scale": 1.0, "mask": None, "cond_indep_stack": (),
        "done": False, "stop": False, "continuation": None, "infer": {} if infer is None else dict(infer),
    }


def sample(name: str, fn, *args, obs=None, infer=None, **kwargs):
    """``pyro.sample``: draw from ``fn`` (rsample when reparameterised) unless observed / replayed."""
    if not _PYRO_STACK:
        if obs is not None:
            return obs
        return fn(*args, **kwargs)
    msg = _new_msg("sample", name, fn, args, kwargs, value=obs, is_observed=obs is not None, infer=infer)
    apply_stack(msg)
    return msg["value"]


def deterministic(name: str, value: torch.Tensor, event_dim: Optional[int] = None):
    """``pyro.deterministic``: records ``value`` as a Delta-like site (no log-prob contribution)."""
    if not _PYRO_STACK:
        return value
    from .distributions import Delta

    ed = value.dim() if event_dim is None else event_dim
    msg = _new_msg("sample", name, Delta(value, event_dim=ed).mask(False), value=value, is_observed=True,
                   infer={"_deterministic": True})
    msg["mask"] = False
    apply_stack(msg)
    return msg["value"]


def param(name: str, init_tensor=None, constraint=constraints.real, event_dim=None):
    """``pyro.param``: fetch or create a learnable parameter; returns the constrained value."""
    def fn(*a, **k):
        if init_tensor is None:
            return _PARAM_STORE[name]
        return _PARAM_STORE.setdefault(name, init_tensor, constraint)

    if not _PYRO_STACK:
        return fn()
    msg = _new_msg("param", name, fn, (), {})
    apply_stack(msg)
    return msg["value"]


# ------------------------------------------------------------------------------------------------------
# plate
# ------------------------------------------------------------------------------------------------------
CondIndepStackFrame = namedtuple("CondIndepStackFrame", ["name", "dim", "size", "counter"])


class plate(Messenger):
    """``pyro.plate(name, size, dim=...)`` without subsampling (the reference never subsamples:
    ``phase_inference_model.py:356-358``, ``velocity_inference_model.py:315-319``).  Entering the plate adds a
    conditional-independence frame to every sample site and expands the site's distribution so that its batch
    shape has ``size`` at ``dim`` (Pyro's BroadcastMessenger)."""

    def __init__(self, name: str, size: int, subsample_size=None, dim: Optional[int] = None, device=None, **_):
        super().__init__()
        if subsample_size is not None and subsample_size != size:
            raise NotImplementedError("velocycle_b200.ppl.plate does not subsample (neither does the reference)")
        if dim is None:
            raise NotImplementedError("plate needs an explicit negative dim (the reference always passes one)")
        if dim >= 0:
            raise ValueError("plate dim must be negative")
        self.name, self.size, self.dim, self.device = name, int(size), dim, device
        self.counter = 0

    def __enter__(self):
        self.counter += 1
        return super().__enter__()

    def _pyro_sample(self, msg):
        frame = CondIndepStackFrame(self.name, self.dim, self.size, self.counter)
        msg["cond_indep_stack"] = (frame,) + tuple(msg["cond_indep_stack"])
        if msg["infer"].get("_deterministic"):
            return
        dist = msg["fn"]
        orig = list(getattr(dist, "batch_shape", ()))
        batch = list(orig)
        k = -self.dim
        if len(batch) < k:  # Pyro's BroadcastMessenger left-pads the batch shape up to the plate's dim
            batch = [1] * (k - len(batch)) + batch
        if batch[self.dim] == 1:
            batch[self.dim] = self.size
        if batch[self.dim] == self.size:
            if batch != orig:
                msg["fn"] = dist.expand(torch.Size(batch))
        else:
            raise ValueError(
                f"Shape mismatch inside plate('{self.name}') at site {msg['name']} dim {self.dim}: "
                f"{batch[self.dim]} vs {self.size}"
            )

    def _pyro_param(self, msg):
        pass
