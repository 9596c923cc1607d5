"""velocycle_b200.ppl -- the probabilistic-programming runtime the host side runs on.

The reference drives its model/guide functions with Pyro (``pyro-ppl==1.8.6``, ``requirements.txt:105``).
Pyro is a third-party dependency that is absent from this image and cannot be installed offline, so this
package provides the subset of Pyro's public API that VeloCycle's hot path touches, with Pyro's semantics
restated from its published behaviour (effect-handler stack, plates with broadcasting, Trace_ELBO with one
particle, SVI, ClippedAdam, Predictive, condition/block/replay/trace).  The names mirror Pyro's so that the
model and guide functions read exactly like the reference's:

    from velocycle_b200 import ppl as pyro
    import velocycle_b200.ppl.distributions as dist

``use_real_pyro()`` returns the real library when it is importable; the drop-in model/guide modules bind to
whichever is active through ``velocycle_b200.ppl.backend``.
"""
from . import distributions, infer, optim, poutine  # noqa: F401
from .primitives import (  # noqa: F401
    clear_param_store,
    deterministic,
    enable_validation,
    get_param_store,
    param,
    plate,
    sample,
    set_rng_seed,
)
from .backend import use_real_pyro  # noqa: F401
