"""Velocity model (mean-field and LRMN variants) with the fused B200 likelihood, and its fit driver.

Drop-in for ``velocycle/velocity_inference_model.py``: ``velocity_latent_variable_model(mp)`` (``:304-388``) and
``velocity_latent_variable_model_LRMN(mp)`` (``:390-471``) keep their sample sites ("logγg", "logβg", "ν", "Δν",
"ϕxy", "νω", "shape_inv", "S", "U"; LRMN adds "rho_real"), deterministic sites ("γg", "ϕ", "ζ", "ζ_dϕ", "ζω",
"ω") and plates.  The likelihood block ``:359-386`` (ElogS, omega, ElogU = -log beta + log(relu(nu.zeta' omega
+ gamma) + 1e-5) + ElogS, and the two GammaPoisson sites) is one call of the fused CUDA op that streams S and
U once.  Reference quirks kept on purpose: the Delta-nu prior is Normal(0, 0.01) regardless of ``mp.σΔν``
(``:332``); the basis order is [1, sin, cos, ...].  ``ElogS`` / ``ElogU`` are not materialised per step.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from ._const import const
from .fused import fused_cycle_nb
from .likelihood import FusedCountLikelihood, count_sites_enabled, packed_counts_for, without_count_sites
from .ppl import backend
from .utils import pack_direction, torch_basis
from .velocity_inference_guide import velocity_latent_variable_guide, velocity_latent_variable_guide_LRMN  # noqa: F401

__all__ = ["velocity_latent_variable_model", "velocity_latent_variable_model_LRMN", "VelocityFitModel"]


def _velocity_model(mp, lrmn: bool):
    pyro, dist, _, _, _ = backend.get()
    dev = mp.device
    if mp.noisemodel != "NegativeBinomial":
        raise ValueError(f"{mp.noisemodel} not allowed: the B200 path implements the NegativeBinomial noise model")
    if mp.basis_kind != "fourier":
        raise ValueError(f"kind={mp.basis_kind!r} is not a valid entry use `fourier`")
    kw = {} if lrmn else {"device": dev}
    cells = pyro.plate("cells", mp.Nc, dim=-1, **kw)
    genes = pyro.plate("genes", mp.Ng, dim=-2, **kw)
    harmonics = pyro.plate("harmonics", mp.Nhω, dim=-3, **kw)
    conditions = pyro.plate("conditions", mp.Nx, dim=-4, **kw)
    batches = pyro.plate("batches", mp.Nb, dim=-5, **kw)

    dnu = None
    with genes:
        loggamma = pyro.sample("logγg", dist.Normal(mp.μγ.to(dev), mp.σγ.to(dev)))
        logbeta = pyro.sample("logβg", dist.Normal(mp.μβ.to(dev), mp.σβ.to(dev)))
        if lrmn:
            pyro.sample("rho_real", dist.Normal(mp.rho_mean, mp.rho_std))
        gamma = torch.exp(loggamma)
        pyro.deterministic("γg", gamma)
        nu = pyro.sample("ν", dist.Normal(mp.μνg.to(dev), mp.σνg.to(dev)).to_event(1))
        if mp.with_delta_nu:
            with batches:
                dnu = pyro.sample("Δν", dist.Normal(const(0.0, dev), const(0.01, dev)))
    with cells:
        phixy = pyro.sample("ϕxy", dist.Normal(mp.φxy_prior.to(dev), const(1.0, dev)).to_event(1))
    phi = pack_direction(phixy)
    pyro.deterministic("ϕ", phi)
    pyro.deterministic("ζ", torch_basis(phi, der=0, kind=mp.basis_kind, **mp.kwargsζ))
    pyro.deterministic("ζ_dϕ", torch_basis(phi, der=1, kind=mp.basis_kind, **mp.kwargsζ_dφ))
    with harmonics, conditions:
        nu_omega = pyro.sample("νω", dist.Normal(mp.μνω.to(dev), mp.σνω.to(dev)))
    zeta_omega = torch_basis(phi, der=0, kind=mp.basis_kind, **mp.kwargsζω).T
    pyro.deterministic("ζω", zeta_omega)

    with genes:
        shape_inv = pyro.sample("shape_inv", dist.Gamma(mp.gamma_alpha.to(dev), mp.gamma_beta.to(dev)))
    counts = packed_counts_for(mp, need_U=True)
    nuw = nu_omega.reshape(mp.Nx, mp.Nhω)
    # per-cell angular speed for the record (tiny); the kernel recomputes it from nu_omega so that its
    # gradient joins the single fused backward
    if counts.cond_id is None:
        cid = torch.zeros(mp.Nc, dtype=torch.long, device=dev)
    else:  # ids in the caller's cell order (PackedCounts may hold the rows sorted by batch)
        cid = (counts.cond_id if counts.perm is None else counts.cond_id[counts.inv_perm]).long()
    pyro.deterministic("ω", (nuw.detach()[cid] * zeta_omega.detach().T).sum(-1).unsqueeze(0))
    if not count_sites_enabled():  # posterior draws of latent / deterministic sites: no pass over the counts
        return
    lp_S, lp_U = fused_cycle_nb(
        counts, phi.reshape(-1), mp.count_factor.reshape(-1), nu.reshape(mp.Ng, -1),
        None if dnu is None else dnu.reshape(mp.Nb, mp.Ng), shape_inv.reshape(-1),
        logbeta.reshape(-1), gamma.reshape(-1), nuw,
    )
    with genes:
        pyro.sample("S", FusedCountLikelihood(lp_S, "S"), obs=mp.S)
        pyro.sample("U", FusedCountLikelihood(lp_U, "U"), obs=mp.U)


def velocity_latent_variable_model(mp, init_loc_fn=None):
    return _velocity_model(mp, lrmn=False)


def velocity_latent_variable_model_LRMN(mp, init_loc_fn=None):
    return _velocity_model(mp, lrmn=True)


class VelocityFitModel:
    """Fit driver with the reference's constructor and ``fit`` signature (``velocity_inference_model.py:32-153``)."""

    max_dense_elements = 200_000_000  # (Ng x Nc) above which fit() does not materialise ElogS / ElogU (+ the averaged pair)

    def __init__(self, metaparams, condition_on={}, early_exit=False, get_posterior=True, num_samples=500, n_per_bin=50):
        _, _, poutine, _, _ = backend.get()
        if len(condition_on) == 0:
            self.model, self.guide = metaparams.model_fn, metaparams.guide_fn
        else:
            self.model = poutine.condition(metaparams.model_fn, data=condition_on)
            self.guide = poutine.block(metaparams.guide_fn, hide=list(condition_on.keys()))
        self.posterior = None
        self.condition = condition_on
        self.condition_on = list(condition_on.keys())
        self.metaparams = metaparams
        self.early_exit = early_exit
        self.get_posterior = get_posterior
        self.num_samples = num_samples
        self.n_per_bin = n_per_bin

    def _return_sites(self):
        rs = ["logγg", "logβg", "νω", "γg", "ν", "ϕxy", "ϕ", "ζ", "ζ_dϕ", "ζω", "ω", "shape_inv"]
        if self.metaparams.with_delta_nu:
            rs.insert(5, "Δν")
        if self.metaparams.model_type == "lrmn":
            rs.append("rho_real")
        return rs

    def fit(self, optimizer, loss=None, num_steps=1000, intermediate_output_step_size=500, store_output=False,
            verbose=True):
        pyro, _, _, infer, _ = backend.get()
        mp = self.metaparams
        if (mp.Ng < 50) & (mp.Nc < 500):
            print("USER WARNING: the number of genes is below the recommended number for reliable velocity-learning.")
        if (mp.Ng < 350) & (mp.Nc < 50):
            print("USER WARNING: the number of cells is below the recommended number for reliable velocity-learning.")
        from .svi import agree_across_ranks, stepper_for

        svi_step = stepper_for(self, self.model, self.guide, optimizer, loss, mp)
        losses, intermediate_output = [], []
        early_exit_bool = False
        for step in range(num_steps):
            step_loss = svi_step()
            losses.append(step_loss)
            if store_output and step % intermediate_output_step_size == 0:
                logging.info("Elbo loss: {}".format(step_loss))
                intermediate_output.append(self.sample_posterior(num_samples=self.n_per_bin, rs=self._return_sites()))
            if verbose and step > 5 and step % 40 == 0:
                logging.info("step %d ELBO loss %.6g", step, step_loss)
            if early_exit_bool:
                if agree_across_ranks(bool(np.abs(np.mean(losses[-100:]) - np.mean(losses[-10:])) < 5), self.metaparams):
                    break
            elif step > 200 and self.early_exit:
                early_exit_bool = True
        self.losses = losses
        self.phis_pyro = pyro.param("ϕxy_locs").detach().squeeze().cpu().numpy().T
        self.fourier_coef = pyro.param("ν_locs").detach().squeeze().cpu().numpy().T
        self.fourier_coef_sd = pyro.param("ν_scales").detach().squeeze().cpu().numpy().T
        self.disp_pyro = pyro.param("shape_inv_locs").detach().squeeze().cpu().numpy().T
        if mp.with_delta_nu:
            self.delta_nus = pyro.param("Δν_locs").detach().unsqueeze(-3).unsqueeze(-4).float().cpu().numpy()
        if mp.model_type != "lrmn":
            self.log_gammas = pyro.param("logγg_locs").detach().squeeze().cpu().numpy().T
            self.velocity_coef = pyro.param("νω_locs").detach().unsqueeze(-3).unsqueeze(-4).float().cpu().numpy()
            self.velocity_coef_sd = pyro.param("νω_scales").detach().unsqueeze(-3).unsqueeze(-4).float().cpu().numpy()
        self.log_betas = pyro.param("logβg_locs").detach().squeeze().cpu().numpy().T
        # the estimates as new container objects (velocity_inference_model.py:160-186); names come from the priors
        from .angularspeed import AngularSpeed
        from .cycle import Cycle
        from .phases import Phases

        genes = getattr(getattr(mp, "cycle_prior", None), "genes", None)
        cells = getattr(getattr(getattr(mp, "phase_prior", None), "phi_xy", None), "columns", None)
        conds = getattr(getattr(mp, "speed_prior", None), "conditions", None)
        self.cycle_pyro = Cycle.from_array(np.atleast_2d(self.fourier_coef), np.atleast_2d(self.fourier_coef_sd), genes)
        self.cycle_pyro.set_log_betas(self.log_betas)
        self.cycle_pyro.set_disp_pyro(self.disp_pyro)
        self.phase_pyro = Phases.from_array(self.phis_pyro, cell_names=cells)
        if mp.model_type != "lrmn":
            self.cycle_pyro.set_log_gammas(self.log_gammas)
            self.speed_pyro = AngularSpeed.from_array(condition_names=conds, means_array=self.velocity_coef.squeeze(),
                                                      stds_array=self.velocity_coef_sd.squeeze(), Nhω=mp.Nhω)
        if self.get_posterior and self.num_samples > 0:
            nbins = int(np.ceil(self.num_samples / self.n_per_bin))
            bins = [self.sample_posterior(num_samples=self.n_per_bin, rs=self._return_sites()) for _ in range(nbins)]
            self.posterior = {k: torch.vstack([b[k] for b in bins]) for k in bins[0]}
            if mp.model_type == "lrmn":
                self.log_gammas = self.posterior["logγg"].mean(0).squeeze().numpy().T
                self.velocity_coef = self.posterior["νω"].mean(0).float().numpy()
                self.velocity_coef_sd = self.posterior["νω"].std(0).float().numpy()
                self.cycle_pyro.set_log_gammas(self.log_gammas)
                self.speed_pyro = AngularSpeed.from_array(condition_names=conds, means_array=self.velocity_coef.squeeze(),
                                                          stds_array=self.velocity_coef_sd.squeeze(), Nhω=mp.Nhω)
            # expected log counts at the fitted parameters / posterior means (velocity_inference_model.py:232-260), on the CPU
            # like the reference; skipped above max_dense_elements: each of the four is a dense (Ng, Nc) matrix
            if mp.Ng * mp.Nc <= self.max_dense_elements:
                from .likelihood import packed_counts_for
                from .posterior import expected_log_counts_summary

                pc = packed_counts_for(mp, need_U=True)
                dev = torch.device(mp.device)  # evaluated where the data live (the reference: on the CPU), returned on the CPU
                unsort = (lambda t: t) if pc.perm is None else (lambda t: t[pc.inv_perm])  # ids in the caller's cell order
                nu = pyro.param("ν_locs").detach().to(dev).reshape(mp.Ng, -1)
                dnu = pyro.param("Δν_locs").detach().to(dev).reshape(mp.Nb, mp.Ng) if mp.with_delta_nu else None
                cid = unsort(pc.cond_id).to(dev) if pc.cond_id is not None else torch.zeros(mp.Nc, dtype=torch.int32, device=dev)
                velo = dict(nu_omega=self.posterior["νω"].mean(0).reshape(mp.Nx, mp.Nhω).to(dev), cond_id=cid,
                            gamma=self.posterior["γg"].mean(0).reshape(-1).to(dev),
                            logbeta=self.posterior["logβg"].mean(0).reshape(-1).to(dev))
                phis = torch.as_tensor(self.phase_pyro.phis, dtype=torch.float32).to(dev)
                summ = expected_log_counts_summary(nu, phis, mp.count_factor.detach().to(dev), dnu,
                                                   unsort(pc.batch_id).to(dev) if mp.with_delta_nu else None, velocity=velo)
                self.posterior.update({k: v.cpu() for k, v in summ.items()})
        if store_output:
            return intermediate_output

    def sample_posterior(self, num_samples=1, rs=None, mp=None, take_mean=True):
        _, _, _, infer, _ = backend.get()
        mp = self.metaparams if mp is None else mp
        fast = self._batched_posterior(mp, num_samples, rs)
        if fast is not None:
            return fast
        pred = infer.Predictive(self.model, guide=self.guide, num_samples=num_samples,
                                return_sites=() if rs is None else rs)
        if rs is not None and not ({"S", "U"} & set(rs)):
            with without_count_sites():
                out = pred(mp)
        else:
            out = pred(mp)
        return {k: v.cpu() for k, v in out.items()}

    def _batched_posterior(self, mp, num_samples, rs):
        """The requested sites for all draws at once (``fastposterior.batched_posterior``) when model and guide are the
        package's own (possibly conditioned the tutorial way) and only latent / deterministic sites are asked for."""
        from . import ppl as shim
        from .faststep import model_code
        from .fastposterior import batched_posterior

        pyro, _, _, _, _ = backend.get()
        if rs is None or ({"S", "U"} & set(rs)) or pyro is not shim or torch.device(mp.device).type != "cuda":
            return None
        found = model_code(self.model, self.guide, mp)
        if found is None:
            return None
        with without_count_sites():  # output shapes as Predictive pads them (one count-free trace, RNG state untouched)
            pad = shim.infer.predictive_padding(self.model, self.guide, mp, device=mp.device)
        out = batched_posterior(mp, found[0], found[1], num_samples, rs, counts=packed_counts_for(mp, need_U=found[0] != 0),
                                pad=pad)
        return {k: v.cpu() for k, v in out.items()}

    def _check_model(self, m, *args):
        pyro, _, poutine, _, _ = backend.get()
        pyro.clear_param_store()
        trace = poutine.trace(m).get_trace(*args)
        print(trace.format_shapes())
        return trace

    def check_model(self):
        return self._check_model(self.model, self.metaparams)

    def check_guide(self):
        return self._check_model(self.guide, self.metaparams)
