"""Phase (manifold-learning) model with the fused B200 likelihood, and its fit driver.

Drop-in for ``velocycle/phase_inference_model.py``: same ``model(mp)`` signature, site names ("ν", "Δν", "ϕxy",
"shape_inv", "S"; deterministic "ϕ", "ζ"), priors and plates (``:343-395``).  The negative-binomial block
(``:368-393``: Fourier basis -> ElogS einsums -> GammaPoisson.log_prob over the (Ng,Nc) matrix) is one call of
the fused CUDA op; ``ElogS`` is therefore not materialised as a deterministic site (nothing in the fit driver
reads it; ``expected_log_counts`` computes it on request).
"""
from __future__ import annotations

import collections
import logging

import numpy as np
import torch

from ._const import const
from .fused import fused_cycle_nb
from .likelihood import FusedCountLikelihood, count_sites_enabled, packed_counts_for, without_count_sites
from .phase_inference_guide import phase_latent_variable_guide
from .ppl import backend
from .utils import pack_direction, torch_fourier_basis

__all__ = ["phase_latent_variable_model", "PhaseFitModel", "expected_log_counts"]


def phase_latent_variable_model(mp):
    pyro, dist, _, _, _ = backend.get()
    dev = mp.device
    if mp.noisemodel != "NegativeBinomial":
        raise ValueError(f"{mp.noisemodel} not allowed: the B200 path implements the NegativeBinomial noise model")
    cells = pyro.plate("cells", mp.Nc, dim=-1, device=dev)
    genes = pyro.plate("genes", mp.Ng, dim=-2, device=dev)
    batches = pyro.plate("batches", mp.Nb, dim=-3, device=dev)

    dnu = None
    with genes:
        nu = pyro.sample("ν", dist.Normal(mp.μνg.to(dev), mp.σνg.to(dev)).to_event(1))
        if mp.with_delta_nu:
            with batches:
                dnu = pyro.sample("Δν", dist.Normal(const(0.0, dev), mp.σΔν.to(dev)))
    with cells:
        phixy = pyro.sample("ϕxy", dist.Normal(mp.φxy_prior.to(dev), const(1.0, dev)).to_event(1))
    phi = pack_direction(phixy)
    pyro.deterministic("ϕ", phi)
    pyro.deterministic("ζ", torch_fourier_basis(phi.squeeze(), num_harmonics=mp.num_harmonics_S, der=0))

    with genes:
        shape_inv = pyro.sample("shape_inv", dist.Gamma(mp.gamma_alpha.to(dev), mp.gamma_beta.to(dev)))
    if not count_sites_enabled():  # posterior draws of latent / deterministic sites: no pass over the counts
        return
    counts = packed_counts_for(mp, need_U=False)
    lp_S, _ = fused_cycle_nb(
        counts, phi.reshape(-1), mp.count_factor.reshape(-1), nu.reshape(mp.Ng, -1),
        None if dnu is None else dnu.reshape(mp.Nb, mp.Ng), shape_inv.reshape(-1),
    )
    with genes:
        pyro.sample("S", FusedCountLikelihood(lp_S, "S"), obs=mp.S)


def expected_log_counts(mp, nu, phi, dnu=None):
    """ElogS (Ng,Nc) on request (the reference stores it as a deterministic site every step)."""
    zeta = torch_fourier_basis(phi.reshape(-1), num_harmonics=(nu.shape[-1] - 1) // 2, der=0)
    out = nu.reshape(mp.Ng, -1) @ zeta.T + mp.count_factor.reshape(1, -1)
    if dnu is not None:
        pc = packed_counts_for(mp, need_U=False)
        bid = (pc.batch_id if pc.perm is None else pc.batch_id[pc.inv_perm]).long()  # the caller's cell order
        out = out + dnu.reshape(mp.Nb, mp.Ng)[bid].T
    return out


class PhaseFitModel:
    """Fit driver with the reference's constructor and ``fit`` signature (``phase_inference_model.py:81-187``).

    ``condition_on`` turns the named sites into observed constants (``poutine.condition``) and hides them from the
    guide (``poutine.block``) exactly like the reference; the SVI loop, loss list and early-exit rule are the same.
    Plotting is not part of this package: ``verbose`` only logs.
    """

    max_dense_elements = 200_000_000  # (Ng x Nc) above which fit() does not materialise ElogS / ElogS2

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

    def fit(self, optimizer, loss=None, num_steps=1000, intermediate_output_step_size=100, store_output=False,
            verbose=True):
        pyro, _, _, infer, _ = backend.get()
        from .svi import agree_across_ranks, stepper_for

        svi_step = stepper_for(self, self.model, self.guide, optimizer, loss, self.metaparams)
        losses, intermediate_output = [], []
        early_exit_bool = False
        for step in range(num_steps):
            step_loss = svi_step()
            losses.append(step_loss)
            if store_output and step % intermediate_output_step_size == 0:
                intermediate_output.append(self.sample_posterior(num_samples=50))
                logging.info("Elbo loss: {}".format(step_loss))
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
        if self.metaparams.with_delta_nu:
            self.delta_nus = pyro.param("Δν_locs").detach().unsqueeze(-3).unsqueeze(-4).float().cpu().numpy()
        # the estimates as new container objects (phase_inference_model.py:186-194); gene / cell names come from the priors
        from .cycle import Cycle
        from .phases import Phases

        mp = self.metaparams
        genes = getattr(getattr(mp, "cycle_prior", None), "genes", None)
        cells = getattr(getattr(getattr(mp, "phase_prior", None), "phi_xy", None), "columns", None)
        self.cycle_pyro = Cycle.from_array(np.atleast_2d(self.fourier_coef), np.atleast_2d(self.fourier_coef_sd), genes)
        self.cycle_pyro.set_disp_pyro(self.disp_pyro)
        self.phase_pyro = Phases.from_array(self.phis_pyro, cell_names=cells)
        if self.get_posterior and self.num_samples > 0:
            nbins = int(np.ceil(self.num_samples / self.n_per_bin))
            rs = ["ν", "ϕxy", "ϕ", "ζ", "shape_inv"] + (["Δν"] if self.metaparams.with_delta_nu else [])
            bins = [self.sample_posterior(num_samples=self.n_per_bin, rs=rs) for _ in range(nbins)]
            self.posterior = {k: torch.vstack([b[k] for b in bins]) for k in bins[0]}
            # expected log counts at the fitted parameters (phase_inference_model.py:241-256), on the CPU like the reference;
            # skipped above max_dense_elements: each is a dense (Ng, Nc) matrix
            if mp.Ng * mp.Nc <= self.max_dense_elements:
                from .posterior import expected_log_counts_summary

                dev = torch.device(mp.device)  # evaluated where the data live (the reference: on the CPU), returned on the CPU
                nu = pyro.param("ν_locs").detach().to(dev).reshape(mp.Ng, -1)
                dnu = pyro.param("Δν_locs").detach().to(dev).reshape(mp.Nb, mp.Ng) if mp.with_delta_nu else None
                bid = self._batch_ids(dev) if mp.with_delta_nu else None
                phis = torch.as_tensor(self.phase_pyro.phis, dtype=torch.float32).to(dev)
                summ = expected_log_counts_summary(nu, phis, mp.count_factor.detach().to(dev), dnu, bid)
                self.posterior.update({k: v.cpu() for k, v in summ.items()})
        if store_output:
            return intermediate_output

    def _batch_ids(self, dev):
        """Batch id of every cell in the CALLER's cell order (PackedCounts may hold the rows sorted by batch)."""
        pc = packed_counts_for(self.metaparams, need_U=False)
        bid = pc.batch_id if pc.perm is None else pc.batch_id[pc.inv_perm]
        return bid.to(dev)

    def sample_posterior(self, num_samples=1, rs=None, mp=None):
        _, _, _, infer, _ = backend.get()
        mp = self.metaparams if mp is None else mp
        fast = self._batched_posterior(mp, num_samples, rs)
        if fast is not None:
            return fast
        pred = infer.Predictive(self.model, guide=self.guide, num_samples=num_samples,
                                return_sites=() if rs is None else rs)
        if rs is not None and "S" not in rs:
            with without_count_sites():
                out = pred(mp)
        else:
            out = pred(mp)
        return {k: v.cpu() for k, v in out.items()}

    def _batched_posterior(self, mp, num_samples, rs):
        """The requested sites for all draws at once (``fastposterior.batched_posterior``) when model and guide are the
        package's own (possibly conditioned) and only latent / deterministic sites are asked for."""
        from . import ppl as shim
        from .faststep import model_code
        from .fastposterior import batched_posterior

        pyro, _, _, _, _ = backend.get()
        if rs is None or "S" in rs or pyro is not shim or torch.device(mp.device).type != "cuda":
            return None
        found = model_code(self.model, self.guide, mp)
        if found is None:
            return None
        with without_count_sites():  # output shapes as Predictive pads them (one count-free trace, RNG state untouched)
            pad = shim.infer.predictive_padding(self.model, self.guide, mp, device=mp.device)
        out = batched_posterior(mp, found[0], found[1], num_samples, rs, pad=pad)
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
