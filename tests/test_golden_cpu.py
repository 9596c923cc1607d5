"""CPU: pin the oracle (and the host-side guides / ppl runtime) to the golden vectors that were produced by
executing the reference's own source (tests/golden/generate_golden.py)."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
CASES = ["case_small", "case_stereo", "case_multi"]
KINDS = ["phase", "phase_nodnu", "velocity", "velocity_lrmn"]


def load(case):
    z = np.load(os.path.join(GOLD, case + ".npz"), allow_pickle=False)
    inp = {k[3:]: torch.as_tensor(z[k].astype(np.int64) if z[k].dtype == np.uint16 else z[k]) for k in z.files
           if k.startswith("in/")}
    return z, inp


def section(z, kind, prefix):
    p = f"{kind}/{prefix}/"
    return {k[len(p):]: torch.as_tensor(z[k]) for k in z.files if k.startswith(p)}


def test_basis_and_packing_match_reference_utils():
    from oracle.likelihood import fourier_basis, pack_direction
    from velocycle_b200.utils import pack_direction as pd2, torch_fourier_basis, unpack_direction

    z = np.load(os.path.join(GOLD, "basis.npz"))
    phi, xy = torch.as_tensor(z["phi"]), torch.as_tensor(z["xy"])
    for H in range(6):
        for der in (0, 1):
            ref = torch.as_tensor(z[f"H{H}_der{der}"])
            assert ref.shape == (phi.numel(), 2 * H + 1)
            assert torch.allclose(fourier_basis(phi, H, der), ref, atol=2e-6, rtol=0)
            assert torch.allclose(torch_fourier_basis(phi, H, der), ref, atol=2e-6, rtol=0)
    assert torch.equal(pack_direction(xy), torch.as_tensor(z["pack_direction"]))
    assert torch.equal(pd2(xy), torch.as_tensor(z["pack_direction"]))
    assert torch.allclose(unpack_direction(phi[:16]), torch.as_tensor(z["unpack_direction"]))


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("kind", KINDS)
def test_oracle_logjoint_matches_reference_model(case, kind):
    """Closed-form fp64 oracle vs the reference model source run in fp32: per-site log-probs, the log-joint and
    its gradient w.r.t. every latent."""
    from oracle.svi import model_logjoint

    z, inp = load(case)
    draws = section(z, kind, "draw")
    okind = "phase" if kind.startswith("phase") else kind
    out = model_logjoint(okind, inp, draws, with_delta_nu=(kind != "phase_nodnu"))
    lps = section(z, kind, "model_lp")
    for site, ref in lps.items():
        got = float(out[f"lp/{site}"])
        assert abs(got - float(ref)) <= 2e-5 * max(1.0, abs(float(ref))), (site, got, float(ref))
    assert abs(float(out["logjoint"]) - float(z[f"{kind}/model_logjoint"])) <= 2e-5 * abs(float(out["logjoint"]))
    grads = section(z, kind, "dlogjoint")
    for site, ref in grads.items():
        got = out[f"d/{site}"]
        err = float((got - ref.double()).abs().max() / (ref.double().abs().max() + 1e-30))
        # fp64 closed form vs the reference source run in fp32: 1e-5.  Exception: with U in the model, the sites
        # that feed a = nu.zeta' omega + gamma inherit the relu(a)+1e-5 kink -- an observed kU > 0 at a ~ 0+ has
        # dL/da = kU/m ~ 1e5 kU, so fp32 rounding of `a` alone moves those gradients by up to ~1e-2 (seen only in
        # case_multi, where the draws come from the wide priors).  The smooth sites of the same run stay at 1e-7.
        kink = kind.startswith("velocity") and site in ("logγg", "ν", "ϕxy", "νω")
        assert err <= (2e-2 if kink else 1e-5), (site, err)


@pytest.mark.parametrize("case", ["case_small", "case_stereo"])
@pytest.mark.parametrize("kind", KINDS)
def test_guides_consume_rng_like_the_reference(case, kind):
    """Our guides, same CPU seed: identical parameter names / initial values and bit-identical draws."""
    from velocycle_b200 import ppl as pyro
    from velocycle_b200.ppl import poutine
    from velocycle_b200.preprocessing import make_phase_metaparams, make_velocity_metaparams

    z, inp = load(case)
    seeds = {"phase": 11, "phase_nodnu": 21, "velocity": 31, "velocity_lrmn": 41}
    seed = seeds[kind] + CASES.index(case)
    common = dict(batch_id=inp["batch_id"], Nb=int(inp["Nb"]), count_factor=inp["cf"], device="cpu")
    if kind.startswith("phase"):
        mp = make_phase_metaparams(inp["S"], inp["U"], inp["mu_nu"], inp["sd_nu"], inp["phixy_prior"],
                                   with_delta_nu=(kind == "phase"), **common)
    else:
        mp = make_velocity_metaparams(inp["S"], inp["U"], inp["mu_nu"], inp["sd_nu"], inp["phixy_prior"],
                                      inp["mu_nw"], inp["sd_nw"], cond_id=inp["cond_id"], Nx=int(inp["Nx"]),
                                      model_type="lrmn" if kind.endswith("lrmn") else "normal", **common)
    pyro.clear_param_store()
    pyro.set_rng_seed(seed)
    tr = poutine.trace(mp.guide_fn).get_trace(mp)
    draws = section(z, kind, "draw")
    for site, ref in draws.items():
        got = tr.nodes[site]["value"].detach()
        assert got.shape == ref.shape, (site, got.shape, ref.shape)
        assert torch.equal(got, ref), site
    params = section(z, kind, "param")
    store = pyro.get_param_store()
    assert set(store.keys()) == set(params.keys())
    for name, ref in params.items():
        if name == "cov_factor":  # initialised from the RNG stream itself: equality is part of the check
            pass
        assert torch.allclose(store.get_unconstrained(name).detach(), ref, atol=0, rtol=0), name
    tr.compute_log_prob()
    for site, ref in section(z, kind, "guide_lp").items():
        assert abs(float(tr.nodes[site]["log_prob_sum"]) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref)))


def test_models_fail_loudly_without_cuda_counts():
    from velocycle_b200 import _lib
    from velocycle_b200 import ppl as pyro
    from velocycle_b200.preprocessing import make_phase_metaparams

    z, inp = load("case_small")
    mp = make_phase_metaparams(inp["S"], inp["U"], inp["mu_nu"], inp["sd_nu"], inp["phixy_prior"],
                               batch_id=inp["batch_id"], Nb=int(inp["Nb"]), count_factor=inp["cf"], device="cpu")
    pyro.clear_param_store()
    with pytest.raises(_lib.VcbError):
        mp.model_fn(mp)


def test_site_shapes_reproduce_appendix_b():
    """The recorded format_shapes() instance (Nc=1849, Ng=76, H=1, Hw=0; Stereo_seq_BrainRG.ipynb cells 84/96)."""
    z, _ = load("case_stereo")
    shapes = {s.split("|")[0]: s.split("|")[1:] for s in z["phase/site_shapes"]}
    assert shapes["ν"][:2] == ["(76, 1)", "(3,)"]
    assert shapes["Δν"][0] == "(1, 76, 1)"
    assert shapes["ϕxy"][:2] == ["(1849,)", "(2,)"]
    assert shapes["shape_inv"][0] == "(76, 1)"
    assert shapes["S"][0] == "(1, 1, 76, 1849)" and shapes["S"][2] == "(76, 1849)"
    v = {s.split("|")[0]: s.split("|")[1:] for s in z["velocity/site_shapes"]}
    assert v["logγg"][0] == "(76, 1)" and v["logβg"][0] == "(76, 1)"
    assert v["Δν"][0] == "(1, 1, 1, 76, 1)"
    assert v["νω"][0] == "(1, 1, 1, 1)"
    assert v["S"][0] == "(1, 1, 76, 1849)" and v["U"][0] == "(1, 1, 76, 1849)"


def test_clipped_adam_matches_float64_restatement():
    from oracle.svi import clipped_adam_reference
    from velocycle_b200.ppl.optim import ClippedAdam

    torch.manual_seed(0)
    p = torch.randn(50, requires_grad=True)
    opt = ClippedAdam({"lr": 0.03, "lrd": 0.999, "betas": (0.8, 0.99)})
    ref_p, m, v = p.detach().double().clone(), torch.zeros(50, dtype=torch.float64), torch.zeros(50, dtype=torch.float64)
    for step in range(1, 6):
        g = torch.randn(50) * 20  # exercises the elementwise clamp at +-10
        p.grad = g.clone()
        opt([p])
        ref_p, m, v = clipped_adam_reference(ref_p, g, m, v, step, 0.03, 0.999, (0.8, 0.99))
        assert torch.allclose(p.detach().double(), ref_p, atol=1e-6)
