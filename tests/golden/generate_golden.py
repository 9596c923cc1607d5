#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by EXECUTING THE REFERENCE'S OWN SOURCE.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/generate_golden.py

What is real and what is restated
---------------------------------
* Real reference code executed here: ``velocycle/utils.py`` (Fourier basis, angle packing) and the model /
  guide functions ``phase_latent_variable_model``, ``phase_latent_variable_guide``,
  ``velocity_latent_variable_model[_LRMN]``, ``velocity_latent_variable_guide[_LRMN]`` from the released
  ``build/lib/velocycle`` tree (the working-tree ``phase_inference_model.py`` does not parse: SURVEY 0.2).
* Restated: the Pyro runtime.  ``pyro-ppl`` is not installable offline, so ``velocycle_b200.ppl`` (Pyro's
  semantics restated) is registered as ``pyro`` before importing the reference; matplotlib / IPython, which
  the reference imports eagerly for plotting, are replaced by empty stubs.

Each ``case_*.npz`` stores the inputs (metaparameters as the reference's preprocessing lays them out), the
guide's RNG draws under a fixed CPU seed, every site's ``log_prob_sum``, the Trace_ELBO loss, the gradient of
the loss w.r.t. every (unconstrained) parameter and the gradient of the model log-joint w.r.t. the latent
values.  ``basis.npz`` stores the real ``torch_fourier_basis`` / ``pack_direction`` outputs.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from collections import namedtuple

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/build/lib"
sys.path.insert(0, ROOT)


def install_stubs():
    from velocycle_b200 import ppl
    from velocycle_b200.ppl import distributions, infer, optim, poutine

    sys.modules["pyro"] = ppl
    sys.modules["pyro.distributions"] = distributions
    sys.modules["pyro.poutine"] = poutine
    sys.modules["pyro.infer"] = infer
    sys.modules["pyro.optim"] = optim
    ag = types.ModuleType("pyro.infer.autoguide")
    for n in ("AutoNormal", "AutoDiagonalNormal", "AutoDelta", "AutoGuideList"):
        setattr(ag, n, getattr(infer.autoguide, n))
    ag.init_to_mean = infer.autoguide.init_to_mean
    ag.init_to_median = infer.autoguide.init_to_median
    sys.modules["pyro.infer.autoguide"] = ag
    infer.autoguide_module = ag
    for name in ("matplotlib", "matplotlib.pyplot", "IPython", "IPython.display"):
        m = types.ModuleType(name)
        sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["IPython.display"].clear_output = lambda *a, **k: None
    sys.modules["IPython"].display = sys.modules["IPython.display"]


def import_reference():
    install_stubs()
    sys.path.insert(0, REF)
    return importlib.import_module("velocycle")


# ----------------------------------------------------------------------------------------------------------
# Metaparameters laid out as preprocessing.py:168-203 (phase) and :270-322 (velocity) lay them out
# ----------------------------------------------------------------------------------------------------------
def make_inputs(Nc, Ng, H, Hw, Nb, Nx, seed):
    from velocycle_b200.synthetic import make_synthetic

    d = make_synthetic(Nc, Ng, H=H, Hw=Hw, Nb=Nb, Nx=Nx, seed=seed, device="cpu", sorted_batches=(Nb * Nx < 6))
    g = torch.Generator().manual_seed(seed + 1000)
    K = 2 * H + 1
    inp = dict(
        S=d.S[:, :Ng].to(torch.int64), U=d.U[:, :Ng].to(torch.int64),  # (Nc,Ng) as anndata layers hold them
        batch_id=d.batch_id.long(), cond_id=d.cond_id.long(), cf=d.cf,
        mu_nu=(d.nu + 0.3 * torch.randn(Ng, K, generator=g)),  # Cycle prior means, (Ng,K)
        sd_nu=(0.5 + torch.rand(Ng, K, generator=g)),
        phixy_prior=torch.stack([torch.cos(d.phi), torch.sin(d.phi)], -1) + 0.2 * torch.randn(Nc, 2, generator=g),
        mu_nw=d.nu_omega.clone(), sd_nw=torch.full_like(d.nu_omega, 0.05) + 0.1 * (torch.arange(2 * Hw + 1) == 0),
        H=H, Hw=Hw, Nb=Nb, Nx=Nx,
    )
    return inp


def phase_mp(vc, inp, with_delta_nu=True):
    S = inp["S"]
    Nc, Ng = S.shape
    Db = torch.nn.functional.one_hot(inp["batch_id"], inp["Nb"])  # (Nc,Nb) int64 like make_design_matrix
    metapars = dict(
        Ng=Ng, Nc=Nc, Nb=inp["Nb"],
        Db=Db.T[:, None, :].float(),
        μνg=inp["mu_nu"][:, None, :].clone(), σνg=inp["sd_nu"][:, None, :].clone(),
        ϕxy_prior=inp["phixy_prior"].clone(),
        gene_selection_model="all",
        model_fn=vc.phase_inference_model.phase_latent_variable_model,
        guide_fn=vc.phase_inference_guide.phase_latent_variable_guide,
        num_harmonics_S=inp["H"], basis_kind="fourier", noisemodel="NegativeBinomial",
        gamma_alpha=torch.tensor(1.0), gamma_beta=torch.tensor(2.0), device=torch.device("cpu"),
        kwargsζ=dict(num_harmonics=inp["H"]), σgc=torch.tensor(0.5), with_delta_nu=with_delta_nu,
        μΔν=torch.tensor(0.0), σΔν=torch.tensor(0.5),
        count_factor=inp["cf"][None, None, None, :].clone(),
        S=S.T.float(), U=inp["U"].T.float(),
    )
    return namedtuple("MetaparContainer", list(metapars.keys()))(**metapars)


def velocity_mp(vc, inp, model_type="normal", with_delta_nu=True):
    S = inp["S"]
    Nc, Ng = S.shape
    D = torch.nn.functional.one_hot(inp["cond_id"], inp["Nx"]).float()
    Db = torch.nn.functional.one_hot(inp["batch_id"], inp["Nb"]).float()
    if model_type == "lrmn":
        model_fn = vc.velocity_inference_model.velocity_latent_variable_model_LRMN
        guide_fn = vc.velocity_inference_guide.velocity_latent_variable_guide_LRMN
    else:
        model_fn = vc.velocity_inference_model.velocity_latent_variable_model
        guide_fn = vc.velocity_inference_guide.velocity_latent_variable_guide
    rep = lambda v: torch.tensor(float(v)).repeat([Ng, 1])
    metapars = dict(
        Ng=Ng, Nc=Nc, Nhω=2 * inp["Hw"] + 1, Nb=inp["Nb"], Nx=inp["Nx"],
        D=D.T[:, None, None, :].clone(), Db=Db.T[:, None, None, None, :].clone(),
        gene_selection_model="all", model_fn=model_fn, guide_fn=guide_fn, with_delta_nu=with_delta_nu,
        μΔν=torch.tensor(0.0), σΔν=torch.tensor(0.1),
        μγ=rep(0.0), σγ=rep(0.5), μβ=rep(2.0), σβ=rep(3.0),
        μνω=inp["mu_nw"][:, :, None, None].clone(), σνω=inp["sd_nw"][:, :, None, None].clone(),
        μνg=inp["mu_nu"][:, None, :].clone(), σνg=inp["sd_nu"][:, None, :].clone(),
        ϕxy_prior=inp["phixy_prior"].clone(), basis_kind="fourier", num_harmonics=inp["H"],
        noisemodel="NegativeBinomial", gamma_alpha=torch.tensor(1.0), gamma_beta=torch.tensor(2.0),
        count_factor=inp["cf"][None, None, None, :].clone(),
        kwargsζ=dict(num_harmonics=inp["H"]), kwargsζ_dϕ=dict(num_harmonics=inp["H"]),
        kwargsζω=dict(num_harmonics=inp["Hw"]),
        S=S.T.float(), U=inp["U"].T.float(), device=torch.device("cpu"), model_type=model_type,
        rho_mean=torch.tensor(4.0), rho_std=torch.tensor(1.0), rho_scale=torch.tensor(1.0), rho_rank=torch.tensor(5),
    )
    return namedtuple("MetaparContainer", list(metapars.keys()))(**metapars)


# ----------------------------------------------------------------------------------------------------------
def run_case(mp, seed, tag):
    """One Trace_ELBO evaluation of the reference model/guide pair: loss, parameter grads, latent grads."""
    from velocycle_b200 import ppl as pyro
    from velocycle_b200.ppl import poutine
    from velocycle_b200.ppl.infer import Trace_ELBO

    pyro.clear_param_store()
    pyro.set_rng_seed(seed)
    model, guide = mp.model_fn, mp.guide_fn
    out = {}
    # (1) the SVI loss and its parameter gradients
    with poutine.trace(param_only=True) as cap:
        loss = Trace_ELBO(num_particles=1).loss_and_grads(model, guide, mp)
    out["loss"] = np.float64(loss)
    store = pyro.get_param_store()
    for name in cap.trace.nodes:
        p = store.get_unconstrained(name)
        out[f"param/{name}"] = p.detach().numpy().copy()
        out[f"grad/{name}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().numpy().copy()
    # (2) same seed again: record the draws and the per-site log-probs
    pyro.set_rng_seed(seed)
    for p in store._params.values():
        p.grad = None
    guide_trace = poutine.trace(guide).get_trace(mp)
    model_trace = poutine.trace(poutine.replay(model, trace=guide_trace)).get_trace(mp)
    model_trace.compute_log_prob()
    guide_trace.compute_log_prob()
    shapes = []
    for name, site in model_trace.nodes.items():
        if site["type"] != "sample":
            continue
        fn = site["fn"]
        shapes.append(f"{name}|{tuple(fn.batch_shape)}|{tuple(fn.event_shape)}|{tuple(site['value'].shape)}")
        if site["infer"].get("_deterministic"):
            continue
        out[f"model_lp/{name}"] = np.float64(site["log_prob_sum"].item())
        if not site["is_observed"]:
            out[f"draw/{name}"] = site["value"].detach().numpy().copy()
    for name, site in guide_trace.nodes.items():
        if site["type"] == "sample":
            out[f"guide_lp/{name}"] = np.float64(site["log_prob_sum"].item())
    out["site_shapes"] = np.array(shapes)
    # (3) gradient of the model log-joint w.r.t. the latent values (conditioned model, no guide)
    leaves = {n: torch.tensor(out[f"draw/{n}"]).requires_grad_(True) for n in
              [k[5:] for k in out if k.startswith("draw/")]}
    tr = poutine.trace(poutine.condition(model, data=leaves)).get_trace(mp)
    tr.compute_log_prob()
    total = sum(s["log_prob_sum"] for s in tr.nodes.values() if s["type"] == "sample")
    total.backward()
    out["model_logjoint"] = np.float64(total.item())
    for n, t in leaves.items():
        out[f"dlogjoint/{n}"] = (t.grad if t.grad is not None else torch.zeros_like(t)).numpy().copy()
    print(f"  {tag}: loss {loss:.6f}  model log-joint {total.item():.6f}  sites {len(shapes)}")
    return out


def save_case(path, inp, results):
    flat = {}
    for k, v in inp.items():
        if isinstance(v, torch.Tensor):
            a = v.numpy()
            if k in ("S", "U"):
                a = a.astype(np.uint16)
                assert (a == v.numpy()).all()
            flat[f"in/{k}"] = a
        else:
            flat[f"in/{k}"] = np.int64(v)
    for tag, res in results.items():
        for k, v in res.items():
            flat[f"{tag}/{k}"] = v
    np.savez_compressed(path, **flat)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def main():
    vc = import_reference()
    print("reference imported from", os.path.dirname(vc.__file__))
    ref_utils = vc.utils
    # ---- basis / packing: the real utils.py ----------------------------------------------------------
    g = torch.Generator().manual_seed(7)
    phi = (torch.rand(257, generator=g) * 2 - 1) * np.pi * 1.5
    xy = torch.randn(64, 2, generator=g) * torch.tensor([1.0, 3.0])
    basis = {"phi": phi.numpy(), "xy": xy.numpy(), "pack_direction": ref_utils.pack_direction(xy).numpy(),
             "unpack_direction": ref_utils.unpack_direction(phi[:16]).numpy()}
    for H in range(0, 6):
        for der in (0, 1):
            basis[f"H{H}_der{der}"] = ref_utils.torch_fourier_basis(phi, num_harmonics=H, der=der).numpy()
    np.savez_compressed(os.path.join(HERE, "basis.npz"), **basis)
    print("wrote basis.npz")

    cases = {
        "case_small": (11, 7, 2, 1, 3, 2),      # SURVEY Appendix A probe shape
        "case_stereo": (1849, 76, 1, 0, 1, 1),  # the recorded format_shapes() instance
        "case_multi": (1024, 256, 3, 1, 16, 2),  # many batches, two conditions, H=3
    }
    for i, (name, (Nc, Ng, H, Hw, Nb, Nx)) in enumerate(cases.items()):
        print(name, (Nc, Ng, H, Hw, Nb, Nx))
        inp = make_inputs(Nc, Ng, H, Hw, Nb, Nx, seed=100 + i)
        results = {
            "phase": run_case(phase_mp(vc, inp, with_delta_nu=True), 11 + i, "phase"),
            "phase_nodnu": run_case(phase_mp(vc, inp, with_delta_nu=False), 21 + i, "phase (no Δν)"),
            "velocity": run_case(velocity_mp(vc, inp, "normal"), 31 + i, "velocity"),
            "velocity_lrmn": run_case(velocity_mp(vc, inp, "lrmn"), 41 + i, "velocity LRMN"),
        }
        save_case(os.path.join(HERE, name + ".npz"), inp, results)


if __name__ == "__main__":
    main()
