"""velocycle_b200: B200-native fused ELBO+gradient hot path for VeloCycle's phase / velocity SVI.

See DESIGN.md.  The CUDA library (``libvcb.so``, C-ABI in ``include/vcb.h``) is loaded lazily by
``velocycle_b200._lib``; there is no CPU fallback.

The package carries the reference's module names (``velocycle/__init__.py``: cycle, phases, angularspeed, utils, preprocessing,
phase_inference_model, phase_inference_guide, velocity_inference_model, velocity_inference_guide; ``plots`` is out of scope);
they are imported on first attribute access, so ``import velocycle_b200 as velocycle; velocycle.preprocessing...`` works
without paying for torch at package import.
"""
import importlib

__version__ = "0.1.0"

_SUBMODULES = (
    "cycle", "phases", "angularspeed", "utils", "preprocessing", "phase_inference_model", "phase_inference_guide",
    "velocity_inference_model", "velocity_inference_guide", "fused", "likelihood", "posterior", "sharding", "svi", "synthetic", "ppl",
)


def __getattr__(name):
    if name in _SUBMODULES:
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def __dir__():
    return sorted(list(globals()) + list(_SUBMODULES))
