"""``Predictive`` of the Pyro subset (velocycle_b200/ppl/infer.py): which sites come back and in which shape.

Pyro 1.8.6 (``pyro/infer/predictive.py``, ``_predictive``) returns every requested site as
``(num_samples,) + (1,) * (max_plate_nesting - len(fn.batch_shape)) + value.shape`` -- the plates (BroadcastMessenger) have
expanded every site's batch shape up to their dim, ``pyro.deterministic`` sites included --; with a guide and no
``return_sites`` it returns ALL model sites (latent, observed, deterministic), without a guide and ``return_sites=()`` the
sites that are not among the posterior samples.  The reference hands this dictionary to the user unchanged
(``velocity_inference_model.py:279-291``), so the shapes are part of the drop-in surface.  The toy model has the reference's
plate structure (genes -2, cells -1, batches -3; a deterministic site inside the gene plate like ``γg``,
``velocity_inference_model.py:322-327``, and one outside every plate like ``ζ``).  When real Pyro is importable the same
model is run through it and the shapes are compared (``pyro_available`` is reported by tests/test_pyro_probe_cpu.py)."""
import pytest
import torch

from velocycle_b200 import ppl as shim
from velocycle_b200.ppl import distributions as sdist
from velocycle_b200.ppl.infer import Predictive


def make(P, D):
    def model():
        gp, cp, bp = P.plate("genes", 3, dim=-2), P.plate("cells", 4, dim=-1), P.plate("batches", 2, dim=-3)
        with gp:
            nu = P.sample("nu", D.Normal(torch.zeros(3, 1, 5), 1.0).to_event(1))
            P.deterministic("g", torch.exp(nu[..., 0]))
            with bp:
                P.sample("d", D.Normal(torch.zeros(2, 3, 1), 1.0))
        with cp:
            xy = P.sample("xy", D.Normal(torch.zeros(4, 2), 1.0).to_event(1))
        P.deterministic("phi", torch.atan2(xy[:, 1], xy[:, 0]))
        with gp, cp:
            P.sample("S", D.Poisson(torch.ones(3, 4)), obs=torch.ones(3, 4))

    def guide():
        gp, cp, bp = P.plate("genes", 3, dim=-2), P.plate("cells", 4, dim=-1), P.plate("batches", 2, dim=-3)
        with gp:
            P.sample("nu", D.Normal(torch.zeros(3, 1, 5), 0.1).to_event(1))
            with bp:
                P.sample("d", D.Normal(torch.zeros(2, 3, 1), 0.1))
        with cp:
            P.sample("xy", D.Normal(torch.ones(4, 2), 0.1).to_event(1))

    return model, guide


ALL = {"nu": (5, 1, 3, 1, 5), "g": (5, 1, 3, 1), "d": (5, 2, 3, 1), "xy": (5, 1, 1, 4, 2), "phi": (5, 1, 1, 1, 4),
       "S": (5, 1, 3, 4)}


def shapes(d):
    return {k: tuple(v.shape) for k, v in d.items()}


def test_with_a_guide_every_model_site_comes_back_left_padded():
    model, guide = make(shim, sdist)
    assert shapes(Predictive(model, guide=guide, num_samples=5)()) == ALL
    assert shapes(Predictive(model, guide=guide, num_samples=5, return_sites=())()) == ALL


def test_return_sites_select():
    model, guide = make(shim, sdist)
    got = Predictive(model, guide=guide, num_samples=5, return_sites=("xy", "phi"))()
    assert shapes(got) == {k: ALL[k] for k in ("xy", "phi")}
    xy = got["xy"].reshape(5, 4, 2)  # the deterministic site is a function of the replayed draw
    assert torch.allclose(got["phi"].reshape(5, 4), torch.atan2(xy[..., 1], xy[..., 0]))


def test_without_a_guide_the_default_is_the_sites_outside_the_posterior_samples():
    model, _ = make(shim, sdist)
    post = {"nu": torch.randn(5, 3, 1, 5), "d": torch.randn(5, 2, 3, 1), "xy": torch.randn(5, 4, 2)}
    got = Predictive(model, posterior_samples=post)()
    assert shapes(got) == {k: ALL[k] for k in ("g", "phi", "S")}
    assert torch.allclose(got["g"].reshape(5, 3, 1), torch.exp(post["nu"][..., 0]))
    assert torch.allclose(got["phi"].reshape(5, 4), torch.atan2(post["xy"][..., 1], post["xy"][..., 0]))
    assert shapes(Predictive(model, posterior_samples=post, return_sites=None)()) == ALL


def test_padding_helper_matches_and_leaves_the_rng_alone():
    """``predictive_padding`` (what the batched posterior draws use to pad their outputs) = the singleton dims of the shapes
    above, computed without consuming random numbers."""
    from velocycle_b200.ppl.infer import predictive_padding

    model, guide = make(shim, sdist)
    torch.manual_seed(5)
    before = torch.random.get_rng_state()
    pad = predictive_padding(model, guide)
    assert torch.equal(torch.random.get_rng_state(), before)
    assert pad == {"nu": 1, "g": 1, "d": 0, "xy": 2, "phi": 3, "S": 1}


def _pyro():
    try:
        import pyro  # noqa: F401

        return pyro
    except Exception:
        return None


@pytest.mark.skipif(_pyro() is None, reason="pyro_available: false")
def test_shapes_match_real_pyro():
    import pyro
    import pyro.distributions as pdist
    from pyro.infer import Predictive as RealPredictive

    model, guide = make(pyro, pdist)
    assert shapes(RealPredictive(model, guide=guide, num_samples=5)()) == ALL
    assert shapes(RealPredictive(model, guide=guide, num_samples=5, return_sites=("nu", "phi"))()) == {
        k: ALL[k] for k in ("nu", "phi")}
    post = {"nu": torch.randn(5, 3, 1, 5), "d": torch.randn(5, 2, 3, 1), "xy": torch.randn(5, 4, 2)}
    assert shapes(RealPredictive(model, posterior_samples=post)()) == {k: ALL[k] for k in ("g", "phi", "S")}
