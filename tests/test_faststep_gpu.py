"""GPU: the fused SVI step (csrc/vcb_svi.cu through faststep.FusedStep) against the step traced through the effect
handlers -- the path that is itself pinned to the goldens generated from the reference's model / guide source
(tests/test_model_golden_gpu.py).  Same seed => same torch draws in the same order, so losses, gradients and parameter
trajectories must agree to fp32 rounding."""
import numpy as np
import pytest
import torch

from test_golden_cpu import load
from test_model_golden_gpu import _mp

pytestmark = pytest.mark.gpu

KINDS = ["phase", "phase_nodnu", "velocity", "velocity_lrmn"]
ARGS = {"lr": 0.03, "lrd": 0.999, "betas": (0.8, 0.99)}


def _run(case, kind, fast, use_graph, steps, seed=321):
    from velocycle_b200 import ppl as pyro
    from velocycle_b200.svi import GraphedSVI

    z, inp = load(case)
    mp = _mp(inp, kind)
    pyro.clear_param_store()
    pyro.set_rng_seed(seed)
    g = GraphedSVI(mp.model_fn, mp.guide_fn, dict(ARGS), mp, use_graph=use_graph, fast=fast)
    losses = [g.step() for _ in range(steps)]
    assert (g._fast is not None) == fast
    return g, np.array(losses)


@pytest.mark.parametrize("case", ["case_stereo", "case_multi"])
@pytest.mark.parametrize("kind", KINDS)
def test_first_step_loss_and_every_gradient(case, kind):
    """One step from the initial parameters: ELBO loss 1e-5 relative, every parameter gradient 1e-4 normwise.  The velocity
    models get the relu-kink allowance of the golden tests (tests/test_golden_cpu.py: a last-bit difference in phi moves
    dL/da = kU / m by per cent where a = 0+): 1e-3 here, 2e-2 on the wide-prior case.  cov_factor rows that start at
    log 0 = -inf carry a zero gradient in both."""
    gt, lt = _run(case, kind, fast=False, use_graph=False, steps=1)
    grads_t = {n: gt.flat_grad[o: o + sz].clone() for n, (o, sz) in gt.param_slices.items()}
    gf, lf = _run(case, kind, fast=True, use_graph=False, steps=1)
    assert abs(lf[0] - lt[0]) <= 1e-5 * abs(lt[0]), (lf, lt)
    assert set(gf.param_slices) == set(grads_t)
    for n, (o, sz) in gf.param_slices.items():
        got, ref = gf.flat_grad[o: o + sz].double(), grads_t[n].double()
        assert torch.isfinite(got).all(), n
        err = float((got - ref).abs().max() / (ref.abs().max() + 1e-30))
        tol = 1e-4 if kind.startswith("phase") else (2e-2 if case == "case_multi" else 1e-3)
        assert err <= tol, (n, err)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("use_graph", [False, True])
def test_trajectory_matches_traced_step(kind, use_graph):
    """12 steps, same seed: losses 1e-5 relative, parameters 5e-3 absolute (the tolerances of
    test_graphed_svi_matches_eager_svi: Adam's normalised update amplifies last-bit gradient differences)."""
    gt, lt = _run("case_stereo", kind, fast=False, use_graph=False, steps=12)
    pt = gt.flat_param.clone()
    gf, lf = _run("case_stereo", kind, fast=True, use_graph=use_graph, steps=12)
    assert np.all(np.abs(lf - lt) <= 1e-5 * np.abs(lt)), (lf, lt)
    pf = gf.flat_param
    finite = torch.isfinite(pt)
    assert torch.equal(finite, torch.isfinite(pf))
    assert float((pf[finite] - pt[finite]).abs().max()) <= 5e-3


def _conditioned(kind, sites, fast, use_graph, steps, seed=77):
    from test_golden_cpu import section
    from velocycle_b200 import ppl as pyro
    from velocycle_b200.ppl import poutine
    from velocycle_b200.svi import GraphedSVI

    z, inp = load("case_stereo")
    mp = _mp(inp, kind)
    draws = section(z, kind, "draw")
    cond = {k: draws[k].cuda() for k in sites}
    pyro.clear_param_store()
    pyro.set_rng_seed(seed)
    g = GraphedSVI(poutine.condition(mp.model_fn, data=cond), poutine.block(mp.guide_fn, hide=list(cond)), dict(ARGS), mp,
                   use_graph=use_graph, fast=fast)
    losses = np.array([g.step() for _ in range(steps)])
    return g, losses


@pytest.mark.parametrize("kind,sites", [("velocity_lrmn", ("ϕxy", "ν", "shape_inv", "Δν")), ("velocity", ("ϕxy", "ν")),
                                        ("phase", ("shape_inv",)), ("velocity_lrmn", ("Δν",))])
def test_conditioned_fit_uses_the_fused_step_and_matches_the_traced_one(kind, sites):
    """The tutorial pattern (condition the velocity model on the phase stage's ϕxy, ν, shape_inv, Δν; block them in the
    guide) and subsets of it: conditioned sites take the given values, keep their prior term, lose the guide term, their
    parameters do not move; the draws are consumed all the same.  Loss 1e-5, gradients 1e-3 (velocity: the relu kink),
    8-step trajectory 5e-3."""
    gt, lt = _conditioned(kind, sites, fast=False, use_graph=False, steps=1)
    grads_t = {n: gt.flat_grad[o: o + sz].clone() for n, (o, sz) in gt.param_slices.items()}
    gf, lf = _conditioned(kind, sites, fast=True, use_graph=False, steps=1)
    assert gf._fast is not None and gf._fast.conditioned == sorted(sites) and gt._fast is None
    assert abs(lf[0] - lt[0]) <= 1e-5 * abs(lt[0]), (lf, lt)
    for n, (o, sz) in gf.param_slices.items():
        got, ref = gf.flat_grad[o: o + sz].double(), grads_t[n].double()
        err = float((got - ref).abs().max() / (ref.abs().max() + 1e-30))
        assert err <= (1e-4 if kind == "phase" else 1e-3), (n, err)
    gt, lt = _conditioned(kind, sites, fast=False, use_graph=False, steps=8)
    pt = gt.flat_param.clone()
    gf, lf = _conditioned(kind, sites, fast=True, use_graph=True, steps=8)
    assert np.all(np.abs(lf - lt) <= 1e-5 * np.abs(lt)), (lf, lt)
    finite = torch.isfinite(pt)
    assert float((gf.flat_param[finite] - pt[finite]).abs().max()) <= 5e-3
    if "ϕxy" in sites:  # conditioned sites receive no updates
        o, sz = gf.param_slices["ϕxy_locs"]
        z, inp = load("case_stereo")
        assert torch.equal(gf.flat_param[o: o + sz].cpu().reshape(-1, 2), inp["phixy_prior"].float())


def test_other_conditioning_falls_back_to_the_traced_step():
    from velocycle_b200 import ppl as pyro
    from velocycle_b200.ppl import poutine
    from velocycle_b200.svi import GraphedSVI

    z, inp = load("case_stereo")
    mp = _mp(inp, "velocity")
    pyro.clear_param_store()
    pyro.set_rng_seed(1)
    cond = {"logβg": torch.full((mp.Ng, 1), 2.0, device="cuda")}
    g = GraphedSVI(poutine.condition(mp.model_fn, data=cond), poutine.block(mp.guide_fn, hide=["logβg"]), dict(ARGS), mp,
                   use_graph=False)
    assert np.isfinite(g.step()) and g._fast is None


def test_clipped_adam_propagates_nan_like_torch_clamp():
    """torch.clamp_ (Pyro's ClippedAdam, the eager path) lets a NaN gradient through; fminf / fmaxf would turn it into
    -clip and a silently drifting parameter."""
    from velocycle_b200 import _lib

    lib = _lib.load()
    n = 1000
    p = torch.ones(n, device="cuda")
    g = torch.full((n,), 0.5, device="cuda")
    g[7] = float("nan")
    g[8] = 1e9
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(1, dtype=torch.int64, device="cuda")
    rc = lib.vcb_clipped_adam(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, step.data_ptr(), 0.03, 1.0, 0.8, 0.99,
                              1e-8, 10.0, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.isnan(p[7]) and torch.isfinite(p[8]) and int(step.item()) == 1
    assert torch.isfinite(torch.cat([p[:7], p[8:]])).all()


def test_packed_count_sidecar_follows_the_design_not_the_address():
    """Metaparameters from the reference's own preprocessing carry no packed counts: they are built on first use and kept on
    the count tensor.  Another design over the same matrix must not reuse them (the old cache was keyed on the
    data pointer alone)."""
    import collections

    from velocycle_b200.likelihood import packed_counts_for

    S = torch.randint(0, 5, (40, 300), device="cuda").float()  # logical (Ng, Nc)
    U = torch.randint(0, 3, (40, 300), device="cuda").float()
    MP = collections.namedtuple("MP", "S U Db D")
    ids1 = torch.arange(300, device="cuda") % 2
    ids2 = torch.arange(300, device="cuda") % 3
    hot = lambda ids, n: torch.nn.functional.one_hot(ids, n).T[:, None, :].float()
    mp1 = MP(S, U, hot(ids1, 2), None)
    mp2 = MP(S, U, hot(ids2, 3), None)
    a = packed_counts_for(mp1, need_U=True)
    assert packed_counts_for(mp1, need_U=True) is a and packed_counts_for(mp1, need_U=False) is a
    b = packed_counts_for(mp2, need_U=True)
    assert b is not a and int(b.batch_id.max()) == 2 and int(a.batch_id.max()) == 1
