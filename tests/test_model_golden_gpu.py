"""GPU: the drop-in model functions (fused CUDA likelihood) against (a) the golden vectors produced by the
reference's own model source and (b) whole SVI trajectories of the unfused restatement under the same RNG."""
import os

import numpy as np
import pytest
import torch

from test_golden_cpu import CASES, KINDS, load, section

pytestmark = pytest.mark.gpu


def _mp(inp, kind, device="cuda"):
    from velocycle_b200.preprocessing import make_phase_metaparams, make_velocity_metaparams

    common = dict(batch_id=inp["batch_id"], Nb=int(inp["Nb"]), count_factor=inp["cf"], device=device)
    if kind.startswith("phase"):
        return make_phase_metaparams(inp["S"], inp["U"], inp["mu_nu"], inp["sd_nu"], inp["phixy_prior"],
                                     with_delta_nu=(kind == "phase"), **common)
    return make_velocity_metaparams(inp["S"], inp["U"], inp["mu_nu"], inp["sd_nu"], inp["phixy_prior"],
                                    inp["mu_nw"], inp["sd_nw"], cond_id=inp["cond_id"], Nx=int(inp["Nx"]),
                                    model_type="lrmn" if kind.endswith("lrmn") else "normal", **common)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("kind", KINDS)
def test_fused_model_matches_reference_golden(case, kind):
    from velocycle_b200 import ppl as pyro
    from velocycle_b200.ppl import poutine

    z, inp = load(case)
    mp = _mp(inp, kind)
    draws = {k: v.cuda().requires_grad_(True) for k, v in section(z, kind, "draw").items()}
    pyro.clear_param_store()
    tr = poutine.trace(poutine.condition(mp.model_fn, data=draws)).get_trace(mp)
    tr.compute_log_prob()
    for site, ref in section(z, kind, "model_lp").items():
        got = float(tr.nodes[site]["log_prob_sum"])
        assert abs(got - float(ref)) <= 1e-4 * max(1.0, abs(float(ref))), (site, got, float(ref))
    total = sum(s["log_prob_sum"] for s in tr.nodes.values() if s["type"] == "sample")
    assert abs(float(total) - float(z[f"{kind}/model_logjoint"])) <= 1e-4 * abs(float(total))
    total.backward()
    for site, ref in section(z, kind, "dlogjoint").items():
        got = draws[site].grad.cpu().double()
        err = float((got - ref.double()).abs().max() / (ref.double().abs().max() + 1e-30))
        # 1e-4, except the relu-kink sites of the wide-prior case (see tests/test_golden_cpu.py): there the fp32
        # reference itself is 5e-3 away from the fp64 truth
        kink = kind.startswith("velocity") and site in ("logγg", "ν", "ϕxy", "νω") and case == "case_multi"
        tol = 2e-2 if kink else (1e-3 if site == "shape_inv" else 1e-4)
        assert err <= tol, (site, err)


def _unfused(kind):
    from oracle import models

    return {"phase": models.phase_model_unfused, "phase_nodnu": models.phase_model_unfused,
            "velocity": models.velocity_model_unfused, "velocity_lrmn": models.velocity_model_unfused_lrmn}[kind]


@pytest.mark.parametrize("kind", KINDS)
def test_svi_trajectory_matches_unfused_reference_chain(kind):
    """Same seed, same device, 25 SVI steps: fused drop-in vs the unfused (Ng,Nc) op chain of the reference.
    Tolerances (stated): per-step ELBO loss 1e-4 relative; after 25 steps parameters within 2e-3 absolute
    (phases phixy_locs, speeds nu_omega_locs, gene harmonics nu_locs) -- ClippedAdam's sign-like first steps
    amplify 1e-6 gradient differences, hence absolute."""
    from velocycle_b200 import ppl as pyro
    from velocycle_b200.ppl.infer import SVI, Trace_ELBO
    from velocycle_b200.ppl.optim import ClippedAdam

    z, inp = load("case_stereo")
    mp = _mp(inp, kind)
    runs = {}
    for name, model in (("fused", mp.model_fn), ("unfused", _unfused(kind))):
        pyro.clear_param_store()
        pyro.set_rng_seed(123)
        svi = SVI(model, mp.guide_fn, ClippedAdam({"lr": 0.03, "lrd": 0.999, "betas": (0.8, 0.99)}), Trace_ELBO())
        losses = [svi.step(mp) for _ in range(25)]
        runs[name] = (losses, {k: v.detach().clone() for k, v in pyro.get_param_store().named_parameters()})
    lf, lu = np.array(runs["fused"][0]), np.array(runs["unfused"][0])
    assert np.all(np.abs(lf - lu) <= 1e-4 * np.abs(lu)), np.max(np.abs(lf - lu) / np.abs(lu))
    for name, pu in runs["unfused"][1].items():
        pf = runs["fused"][1][name]
        finite = torch.isfinite(pu)  # cov_factor is initialised with exact zeros under a positive constraint (-inf)
        assert torch.equal(finite, torch.isfinite(pf)), name
        assert float((pf[finite] - pu[finite]).abs().max()) <= 2e-3, (name, float((pf[finite] - pu[finite]).abs().max()))


def test_conditioned_velocity_fit_driver_runs_and_matches():
    """The tutorial pattern: condition the velocity model on the phase-stage estimates (phixy, nu, shape_inv, Δν)."""
    from velocycle_b200 import ppl as pyro
    from velocycle_b200.ppl.optim import ClippedAdam
    from velocycle_b200.velocity_inference_model import VelocityFitModel

    z, inp = load("case_stereo")
    mp = _mp(inp, "velocity_lrmn")
    draws = section(z, "velocity_lrmn", "draw")
    cond = {k: draws[k].cuda() for k in ("ϕxy", "ν", "shape_inv", "Δν")}
    pyro.clear_param_store()
    pyro.set_rng_seed(5)
    fit = VelocityFitModel(mp, condition_on=cond, num_samples=4, n_per_bin=2)
    fit.fit(ClippedAdam({"lr": 0.03, "betas": (0.8, 0.99)}), num_steps=15, verbose=False)
    assert len(fit.losses) == 15 and np.isfinite(fit.losses).all()
    assert fit.losses[-1] < fit.losses[0]
    # the driver ran the graphed, fused step with the four sites held at the given values
    g = next(iter(fit._steppers.values()))[0]
    assert g._graph is not None and g.steps_done == 15
    assert g._fast is not None and g._fast.conditioned == sorted(cond)
    assert fit.posterior["νω"].shape[0] == 4 and fit.posterior["ω"].shape[-1] == mp.Nc
    # conditioned sites receive no updates
    assert torch.equal(pyro.param("ϕxy_locs").detach().cpu(), inp["phixy_prior"].float())
    # the estimates come back as container objects, like the reference's fit() (velocity_inference_model.py:160-277)
    K, Kw = mp.μνg.shape[-1], mp.Nhω
    assert fit.cycle_pyro.means.shape == (K, mp.Ng) and fit.cycle_pyro.stds.shape == (K, mp.Ng)
    assert np.array_equal(fit.cycle_pyro.means.values, fit.fourier_coef)
    assert fit.cycle_pyro.log_betas.shape == (mp.Ng,) and fit.cycle_pyro.log_gammas.shape == (mp.Ng,)
    assert fit.cycle_pyro.disp_pyro.shape == (mp.Ng,)
    assert tuple(fit.phase_pyro.phis.shape) == (mp.Nc,) and float(fit.phase_pyro.phis.min()) >= 0.0
    assert fit.speed_pyro.means.shape == (Kw, mp.Nx) and list(fit.speed_pyro.means.index)[0] == "nu0"
    # ... and the expected log counts at the fitted parameters (velocity_inference_model.py:232-260)
    for k in ("ElogS", "ElogU", "ElogS2", "ElogU2"):
        assert tuple(fit.posterior[k].shape) == (mp.Ng, mp.Nc) and bool(torch.isfinite(fit.posterior[k]).all()), k
    d = fit.posterior["ElogS"] - fit.posterior["ElogS2"]                     # differs by cf_c - mean(cf) only
    cf = mp.count_factor.reshape(-1).cpu()
    assert torch.allclose(d, (cf - cf.mean()).expand_as(d), atol=1e-5)


def test_phase_fit_driver_returns_containers():
    from velocycle_b200 import ppl as pyro
    from velocycle_b200.phase_inference_model import PhaseFitModel
    from velocycle_b200.ppl.optim import ClippedAdam

    z, inp = load("case_stereo")
    mp = _mp(inp, "phase")
    pyro.clear_param_store()
    pyro.set_rng_seed(5)
    fit = PhaseFitModel(mp, num_samples=2, n_per_bin=2)
    fit.fit(ClippedAdam({"lr": 0.03, "betas": (0.8, 0.99)}), num_steps=8, verbose=False)
    assert len(fit.losses) == 8 and np.isfinite(fit.losses).all()
    g = next(iter(fit._steppers.values()))[0]
    assert g._graph is not None and g._fast is not None  # the reference's public entry point runs the fused, graphed step
    K = mp.μνg.shape[-1]
    assert fit.cycle_pyro.means.shape == (K, mp.Ng) and np.array_equal(fit.cycle_pyro.stds.values, fit.fourier_coef_sd)
    assert fit.cycle_pyro.disp_pyro.shape == (mp.Ng,)
    assert fit.phase_pyro.phi_xy.shape == (2, mp.Nc)
    assert np.array_equal(fit.phase_pyro.phi_xy_tensor.numpy(), fit.phis_pyro.astype(np.float32))
    assert tuple(fit.posterior["ElogS"].shape) == (mp.Ng, mp.Nc) and "ElogS2" in fit.posterior and "ElogU" not in fit.posterior
    fit.max_dense_elements = 0                                              # the dense summaries are optional
    fit.fit(ClippedAdam({"lr": 0.03, "betas": (0.8, 0.99)}), num_steps=2, verbose=False)
    assert "ElogS" not in fit.posterior and "ν" in fit.posterior


@pytest.mark.parametrize("kind", ["phase", "velocity", "velocity_lrmn"])
@pytest.mark.parametrize("use_graph", [False, True])
def test_graphed_svi_matches_eager_svi(kind, use_graph):
    """GraphedSVI (flat parameters, fused ClippedAdam, optional CUDA-graph replay) against the eager
    ``ppl.infer.SVI`` + per-parameter ClippedAdam on the same seed: losses 1e-5 relative; parameters 5e-3 absolute
    (Adam's normalised update turns a last-bit gradient difference on a near-zero gradient into O(lr) steps)."""
    from velocycle_b200 import ppl as pyro
    from velocycle_b200.ppl.infer import SVI, Trace_ELBO
    from velocycle_b200.ppl.optim import ClippedAdam
    from velocycle_b200.svi import GraphedSVI

    z, inp = load("case_stereo")
    mp = _mp(inp, kind)
    args = {"lr": 0.03, "lrd": 0.999, "betas": (0.8, 0.99)}
    pyro.clear_param_store()
    pyro.set_rng_seed(321)
    svi = SVI(mp.model_fn, mp.guide_fn, ClippedAdam(dict(args)), Trace_ELBO())
    eager_losses = [svi.step(mp) for _ in range(12)]
    eager = {k: v.detach().clone() for k, v in pyro.get_param_store().named_parameters()}
    pyro.clear_param_store()
    pyro.set_rng_seed(321)
    gsvi = GraphedSVI(mp.model_fn, mp.guide_fn, dict(args), mp, use_graph=use_graph)
    losses = [gsvi.step() for _ in range(12)]
    le, lg = np.array(eager_losses), np.array(losses)
    assert np.all(np.abs(le - lg) <= 1e-5 * np.abs(le)), (le, lg)
    for name, ref in eager.items():
        got = pyro.get_param_store().get_unconstrained(name).detach()
        finite = torch.isfinite(ref)
        assert float((got[finite] - ref[finite]).abs().max()) <= 5e-3, name


@pytest.mark.parametrize("case", ["case_stereo", "case_multi"])
@pytest.mark.parametrize("kind", ["phase", "phase_nodnu", "velocity", "velocity_lrmn"])
@pytest.mark.parametrize("conditioned", [False, True])
def test_batched_posterior_equals_sequential_predictive(case, kind, conditioned):
    """``sample_posterior`` of the fit drivers (all draws of a bin at once, fastposterior.py) against the sequential
    ``Predictive`` over the traced guide and model on the same seed: same keys, same shapes, same values (1e-6; the draws
    themselves are bit-identical).  Also with the tutorial's conditioning."""
    from velocycle_b200 import phase_inference_model as pm, ppl as pyro, velocity_inference_model as vm
    from velocycle_b200.likelihood import without_count_sites
    from velocycle_b200.ppl.infer import Predictive

    z, inp = load(case)
    mp = _mp(inp, kind)
    cond = {}
    if conditioned:
        draws = section(z, kind, "draw")
        cond = {k: draws[k].cuda() for k in (("ϕxy", "ν", "shape_inv", "Δν") if kind.startswith("velocity") else ("shape_inv",))
                if k in draws}
    Driver = pm.PhaseFitModel if kind.startswith("phase") else vm.VelocityFitModel
    driver = Driver(mp, condition_on=cond, get_posterior=False)
    rs = (["ν", "ϕxy", "ϕ", "ζ", "shape_inv"] + (["Δν"] if mp.with_delta_nu else [])) if kind.startswith("phase") \
        else driver._return_sites()
    pyro.clear_param_store()
    pyro.set_rng_seed(9)
    with without_count_sites():
        ref = {k: v.cpu() for k, v in Predictive(driver.model, guide=driver.guide, num_samples=3, return_sites=rs)(mp).items()}
    pyro.set_rng_seed(9)
    got = driver._batched_posterior(mp, 3, rs)
    assert got is not None and set(got) == set(ref), (set(got or ()), set(ref))
    for k in ref:
        assert tuple(got[k].shape) == tuple(ref[k].shape), (k, tuple(got[k].shape), tuple(ref[k].shape))
        assert torch.allclose(got[k], ref[k], rtol=1e-6, atol=1e-6), k
    assert torch.equal(got["ϕxy"], ref["ϕxy"])
    # the layout Pyro's Predictive produces (tests/test_predictive_shapes_cpu.py): singleton dims up to the model's plate
    # nesting minus the site's own batch dims.  Nesting: cells -1, genes -2, [harmonics -3, conditions -4 in the velocity
    # model,] batches -3 / -5 -- entered only with Δν
    Ng, Nc, K = int(mp.Ng), int(mp.Nc), int(mp.μνg.shape[-1])
    if kind.startswith("phase"):
        mpn = 3 if mp.with_delta_nu else 2
        sites = {"ν": (2, (Ng, 1, K)), "ϕ": (0, (Nc,)), "ϕxy": (1, (Nc, 2))}
    else:
        mpn = 5 if mp.with_delta_nu else 4
        sites = {"ν": (2, (Ng, 1, K)), "ϕ": (0, (Nc,)), "γg": (2, (Ng, 1)), "νω": (4, (int(mp.Nx), int(mp.Nhω), 1, 1)),
                 "ϕxy": (1, (Nc, 2))}
    for k, (bdim, vshape) in sites.items():
        want = (3,) + (1,) * (mpn - bdim) + vshape
        assert tuple(got[k].shape) == want, (k, tuple(got[k].shape), want)


@pytest.mark.parametrize("kind", ["phase", "velocity_lrmn"])
def test_posterior_draws_touch_no_counts_and_equal_predictive(kind, monkeypatch):
    """``sample_posterior`` with latent / deterministic ``return_sites`` (what ``fit`` asks for, velocity_inference_model.py:
    213-224): same seed => the same tensors as a plain ``Predictive`` over the full model, but the fused likelihood is never
    called (the reference re-evaluates the S / U sites over the (Ng, Nc) matrices for each of the 500 draws)."""
    from velocycle_b200 import phase_inference_model as pm, ppl as pyro, velocity_inference_model as vm
    from velocycle_b200.ppl.infer import Predictive

    z, inp = load("case_stereo")
    mp = _mp(inp, kind)
    mod = pm if kind == "phase" else vm
    driver = (pm.PhaseFitModel if kind == "phase" else vm.VelocityFitModel)(mp, get_posterior=False)
    rs = ["ν", "ϕxy", "ϕ", "ζ", "shape_inv", "Δν"] if kind == "phase" else driver._return_sites()
    pyro.clear_param_store()
    pyro.set_rng_seed(9)
    ref = {k: v.cpu() for k, v in Predictive(driver.model, guide=driver.guide, num_samples=3, return_sites=rs)(mp).items()}
    calls = []
    real = mod.fused_cycle_nb
    monkeypatch.setattr(mod, "fused_cycle_nb", lambda *a, **k: (calls.append(1), real(*a, **k))[1])
    pyro.set_rng_seed(9)
    got = driver.sample_posterior(num_samples=3, rs=rs)
    assert not calls
    assert set(got) == set(ref)
    for k in ref:
        assert torch.allclose(got[k], ref[k], rtol=1e-6, atol=1e-6), k
    driver.sample_posterior(num_samples=1, rs=None)  # the default (every latent site) keeps the full model
    assert calls
