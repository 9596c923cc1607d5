"""Metaparameter builders: anndata-facing API of ``velocycle/preprocessing.py:103-323`` plus tensor-level
builders that never create dense host copies.

``preprocess_for_phase_estimation`` / ``preprocess_for_velocity_estimation`` keep the reference's signatures and
return the same flat ``MetaparContainer`` namedtuple (same field names, shapes and dtypes), with two additions:
``packed_counts`` (device-resident cell-major float32 counts padded to 16-byte rows, int32 batch / condition
ids and the count spectra the kernels need) and ``model_fn`` / ``guide_fn`` pointing at this package's fused
model functions.  ``anndata`` itself is not imported: any object with ``.layers[...]`` (dense or scipy-sparse,
cells x genes), ``.obs`` and ``.var`` works.  Deviations forced by library drift since the reference's pins:
sparse ``.A`` (removed in scipy 1.14) is replaced by ``.toarray()``.
"""
from __future__ import annotations

from collections import namedtuple
from typing import Optional

import numpy as np
import torch

from .fused import PackedCounts
from .phase_inference_guide import phase_latent_variable_guide
from .phase_inference_model import phase_latent_variable_model
from .velocity_inference_guide import velocity_latent_variable_guide, velocity_latent_variable_guide_LRMN
from .velocity_inference_model import velocity_latent_variable_model, velocity_latent_variable_model_LRMN

__all__ = [
    "make_design_matrix", "make_phase_metaparams", "make_velocity_metaparams", "filter_shared_genes",
    "preprocess_for_phase_estimation", "preprocess_for_velocity_estimation", "count_factor_from_totals",
]


def _container(d: dict):
    return namedtuple("MetaparContainer", list(d.keys()))(**d)


def _dense(layer) -> np.ndarray:
    if hasattr(layer, "toarray"):
        layer = layer.toarray()
    return np.asarray(layer)


def make_design_matrix(anndata, ids="batch") -> torch.Tensor:
    """One-hot (Nc, n_levels) int64 design matrix in order of first appearance (``preprocessing.py:65-93``)."""
    if ids not in anndata.obs.columns:
        raise ValueError(f"{ids=} is not a valid entry anndata.obs")
    levels: dict = {}
    codes = np.array([levels.setdefault(v, len(levels)) for v in np.asarray(anndata.obs[ids])])
    return torch.nn.functional.one_hot(torch.as_tensor(codes), len(levels)).to(torch.int64)


def filter_shared_genes(cycle, data, filter_type: str = "intersection"):
    """Restrict a Cycle prior and an AnnData-like object to a common gene set, SORTED by gene name
    (``preprocessing.py:20-63``).  ``intersection``: genes present in both; ``union``: every gene of the data (all genes of the
    Cycle must be in the data), the Cycle extended with uninformative entries (means 0, stds 10) for the new ones."""
    from .cycle import Cycle, reorder

    cycle_genes, data_genes = set(cycle.genes), set(data.var.index)
    if filter_type == "intersection":
        keep = np.array(sorted(cycle_genes & data_genes))
        return Cycle.from_array(means_array=cycle.means[keep], stds_array=cycle.stds[keep]), data[:, keep].copy()
    if filter_type == "union":
        if cycle_genes - data_genes:
            raise Exception("Gene features detected in Cycle object cannot be found in AnnData object")
        keep = np.array(sorted(cycle_genes | data_genes))
        new_cycle = Cycle.from_array(means_array=cycle.means, stds_array=cycle.stds)
        extra = np.array(sorted(data_genes - cycle_genes))
        if extra.size:
            new_cycle.extend(gene_names=extra)
        return reorder(new_cycle, keep), data[:, keep].copy()
    raise Exception(f"{filter_type=} is not a supported gene filtering behavior")


def count_factor_from_totals(S_totals: torch.Tensor) -> torch.Tensor:
    """log(UMI_c / mean UMI) over ALL cells (``preprocessing.py:149-152``); compute before sharding."""
    t = S_totals.float()
    return torch.log(t / t.mean())


def _is_sparse(M) -> bool:
    return hasattr(M, "tocsr") and not isinstance(M, torch.Tensor)


def _pack(S_cells_by_genes, device) -> torch.Tensor:
    """(Nc, Ng) counts -> the packed float32 device matrix.  A scipy.sparse matrix (the anndata layers) goes to a CUDA device
    as CSR and is scattered there (``PackedCounts.pack_csr``); dense input is copied and widened."""
    if _is_sparse(S_cells_by_genes):
        if torch.device(device).type == "cuda":
            return PackedCounts.pack_csr(S_cells_by_genes, device)
        S_cells_by_genes = torch.as_tensor(S_cells_by_genes.toarray())
    return PackedCounts.pack_matrix(torch.as_tensor(S_cells_by_genes).to(device), layout="cells_by_genes")


def _counts_layer(layer):
    """An anndata layer as the builders take it: sparse stays sparse (no dense host copy), anything else becomes int64."""
    if _is_sparse(layer):
        return layer
    return torch.as_tensor(np.asarray(layer).astype(np.int64))


def make_phase_metaparams(
    S, U, mu_nu, sd_nu, phixy_prior, batch_id=None, Nb=1, count_factor=None, n_harmonics=None,
    with_delta_nu=True, μΔν=0.0, σΔν=0.5, gamma_alpha=1.0, gamma_beta=2.0, device="cuda",
    cycle_prior=None, phase_prior=None, spectrum=True, shard=None,
):
    """Tensor-level builder.  S, U: (Nc, Ng) counts (any dtype/device); mu_nu, sd_nu: (Ng, K); phixy_prior (Nc,2)."""
    device = torch.device(device)
    Nc, Ng = S.shape
    K = mu_nu.shape[1]
    H = (K - 1) // 2 if n_harmonics is None else n_harmonics
    Sp = _pack(S, device)
    Up = None if U is None else _pack(U, device)
    bid = torch.zeros(Nc, dtype=torch.int32) if batch_id is None else batch_id.to(torch.int32)
    counts = None
    if device.type == "cuda":  # on a CPU device only the (host-side) guides can run; the models raise
        counts = PackedCounts(Sp, Up, Ng, bid, None, spectrum=False)
        if spectrum:
            from .fused import CountSpectrum

            counts.spec_S = CountSpectrum.build(Sp, Nc, Ng, Sp.shape[1])  # phase needs the S spectrum only
    if count_factor is None:
        count_factor = count_factor_from_totals(Sp[:, :Ng].sum(1))
    f = lambda v: torch.as_tensor(v, dtype=torch.float32, device=device)
    Db = torch.nn.functional.one_hot(bid.long(), Nb).T[:, None, :].float().to(device)
    d = dict(
        Ng=Ng, Nc=Nc, Nb=Nb, Db=Db, cycle_prior=cycle_prior, phase_prior=phase_prior,
        μνg=f(mu_nu)[:, None, :], σνg=f(sd_nu)[:, None, :], ϕxy_prior=f(phixy_prior),
        gene_selection_model="all", model_fn=phase_latent_variable_model, guide_fn=phase_latent_variable_guide,
        num_harmonics_S=H, basis_kind="fourier", noisemodel="NegativeBinomial",
        gamma_alpha=f(gamma_alpha), gamma_beta=f(gamma_beta), device=device, kwargsζ=dict(num_harmonics=H),
        σgc=f(0.5), with_delta_nu=with_delta_nu, μΔν=f(μΔν), σΔν=f(σΔν),
        count_factor=f(count_factor).reshape(1, 1, 1, Nc),
        S=Sp[:, :Ng].T, U=None if Up is None else Up[:, :Ng].T,  # logical (Ng,Nc) views of the packed buffers
        packed_counts=counts, shard=shard,
    )
    if counts is not None:
        counts.shard = shard
    return _container(d)


def make_velocity_metaparams(
    S, U, mu_nu, sd_nu, phixy_prior, mu_nu_omega, sd_nu_omega, batch_id=None, cond_id=None, Nb=1, Nx=1,
    count_factor=None, n_harmonics=None, ω_n_harmonics=None, with_delta_nu=True, model_type="lrmn",
    μγ=0.0, σγ=0.5, μβ=2.0, σβ=3.0, μΔν=0.0, σΔν=0.1, gamma_alpha=1.0, gamma_beta=2.0,
    rho_mean=4.0, rho_std=1.0, rho_scale=1.0, rho_rank=5, device="cuda",
    cycle_prior=None, phase_prior=None, speed_prior=None, shard=None,
):
    """Tensor-level builder.  mu_nu_omega, sd_nu_omega: (Nx, Kw)."""
    device = torch.device(device)
    Nc, Ng = S.shape
    K = mu_nu.shape[1]
    H = (K - 1) // 2 if n_harmonics is None else n_harmonics
    Kw = mu_nu_omega.shape[1]
    Hw = (Kw - 1) // 2 if ω_n_harmonics is None else ω_n_harmonics
    Sp, Up = _pack(S, device), _pack(U, device)
    bid = torch.zeros(Nc, dtype=torch.int32) if batch_id is None else batch_id.to(torch.int32)
    cid = torch.zeros(Nc, dtype=torch.int32) if cond_id is None else cond_id.to(torch.int32)
    counts = PackedCounts(Sp, Up, Ng, bid, cid, spectrum=True) if device.type == "cuda" else None
    if count_factor is None:
        count_factor = count_factor_from_totals(Sp[:, :Ng].sum(1))
    f = lambda v: torch.as_tensor(v, dtype=torch.float32, device=device)
    rep = lambda v: f(v).reshape(-1)[:1].repeat([Ng, 1]) if f(v).numel() == 1 else f(v).reshape(Ng, 1)
    if model_type == "lrmn":
        model_fn, guide_fn = velocity_latent_variable_model_LRMN, velocity_latent_variable_guide_LRMN
    else:
        model_fn, guide_fn = velocity_latent_variable_model, velocity_latent_variable_guide
    d = dict(
        Ng=Ng, Nc=Nc, Nhω=Kw, Nb=Nb, Nx=Nx,
        D=torch.nn.functional.one_hot(cid.long(), Nx).T[:, None, None, :].float().to(device),
        Db=torch.nn.functional.one_hot(bid.long(), Nb).T[:, None, None, None, :].float().to(device),
        cycle_prior=cycle_prior, phase_prior=phase_prior, speed_prior=speed_prior,
        gene_selection_model="all", model_fn=model_fn, guide_fn=guide_fn, with_delta_nu=with_delta_nu,
        μΔν=f(μΔν), σΔν=f(σΔν), μγ=rep(μγ), σγ=rep(σγ), μβ=rep(μβ), σβ=rep(σβ),
        μνω=f(mu_nu_omega)[:, :, None, None], σνω=f(sd_nu_omega)[:, :, None, None],
        μνg=f(mu_nu)[:, None, :], σνg=f(sd_nu)[:, None, :], ϕxy_prior=f(phixy_prior),
        basis_kind="fourier", num_harmonics=H, noisemodel="NegativeBinomial",
        gamma_alpha=f(gamma_alpha), gamma_beta=f(gamma_beta),
        count_factor=f(count_factor).reshape(1, 1, 1, Nc),
        kwargsζ=dict(num_harmonics=H), kwargsζ_dϕ=dict(num_harmonics=H), kwargsζω=dict(num_harmonics=Hw),
        σₛgc=f(0.1), σᵤgc=f(0.1), S=Sp[:, :Ng].T, U=Up[:, :Ng].T, device=device, model_type=model_type,
        rho_mean=f(rho_mean), rho_std=f(rho_std), rho_scale=f(rho_scale), rho_rank=torch.tensor(int(rho_rank)),
        packed_counts=counts, shard=shard,
    )
    if counts is not None:
        counts.shard = shard
    return _container(d)


# ------------------------------------------------------------------------------------------------------
# anndata-facing wrappers with the reference's signatures
# ------------------------------------------------------------------------------------------------------
def _ids_from_design(design_mtx) -> torch.Tensor:
    return torch.as_tensor(np.asarray(design_mtx)).argmax(-1).to(torch.int32)


def preprocess_for_phase_estimation(
    anndata, cycle_obj, phase_obj, design_mtx, n_harmonics: int = 2, gene_selection_model: str = "all",
    normalize: bool = False, behavior: str = "intersection", noisemodel="NegativeBinomial",
    with_delta_nu: bool = True, condition_on={}, μΔν=torch.tensor(0).float(), σΔν=torch.tensor(0.5).float(),
    gamma_alpha=torch.tensor(1.0).float(), gamma_beta=torch.tensor(2.0).float(), beta0=0.10, beta1=0.90,
    device=torch.device("cuda"),
):
    """``velocycle/preprocessing.py:103-205`` for the NegativeBinomial model.  ``cycle_obj`` / ``phase_obj`` need
    ``means_tensor`` / ``stds_tensor`` (K x Ng) and ``phi_xy_tensor`` (2 x Nc) like the reference's containers."""
    if gene_selection_model != "all":
        raise ValueError(f"{gene_selection_model=} is not a valid model")
    if noisemodel != "NegativeBinomial" or normalize:
        raise ValueError("the B200 path implements the NegativeBinomial noise model on raw integer counts")
    S = _counts_layer(anndata.layers["spliced"])
    U = _counts_layer(anndata.layers["unspliced"])
    mp = make_phase_metaparams(
        S, U, cycle_obj.means_tensor.T, cycle_obj.stds_tensor.T, phase_obj.phi_xy_tensor.T,
        batch_id=_ids_from_design(design_mtx), Nb=int(np.asarray(design_mtx).shape[-1]), n_harmonics=n_harmonics,
        with_delta_nu=with_delta_nu, μΔν=μΔν, σΔν=σΔν, gamma_alpha=gamma_alpha, gamma_beta=gamma_beta, device=device,
        cycle_prior=cycle_obj, phase_prior=phase_obj,
    )
    # the remaining fields of the reference's container (preprocessing.py:168-203), laid out as the reference lays them out;
    # the dense (Ng, Nc) logS / logU matrices (used by the Lognormal variants and by plotting only) are not materialised
    d = mp._asdict()
    d["count_factor"] = mp.count_factor.reshape(1, 1, mp.Nc)
    d["condition"] = np.array(list(condition_on.keys()))
    d["beta0"], d["beta1"] = torch.tensor(beta0).to(device), torch.tensor(beta1).to(device)
    return _container(d)


def preprocess_for_velocity_estimation(
    anndata, cycle_obj, phase_obj, speed_obj, condition_design_mtx, batch_design_mtx, device=torch.device("cuda"),
    gene_selection_model: str = "all", null_cycle_obj=None, n_harmonics: int = 2, norm_size: int = 1000,
    with_delta_nu: bool = True, count_factor=0, count_factorU=0, ω_n_harmonics: int = 1, normalize: bool = False,
    behavior: str = "intersection", noisemodel="NegativeBinomial", condition_on={},
    μγ=torch.tensor(0.0).float(), σγ=torch.tensor(0.5).float(), μβ=torch.tensor(2.0).float(),
    σβ=torch.tensor(3.0).float(), μΔν=torch.tensor(0).float(), σΔν=torch.tensor(0.1).float(),
    gamma_alpha=torch.tensor(1.0).float(), gamma_beta=torch.tensor(2.0).float(), model_type: str = "lrmn",
    rho_mean=torch.tensor(4.0), rho_std=torch.tensor(1.0), rho_scale=torch.tensor(1.0), rho_rank=torch.tensor(5),
):
    """``velocycle/preprocessing.py:207-323`` for the NegativeBinomial model (``count_factor`` is the tensor the
    phase stage produced, as in the tutorials)."""
    if gene_selection_model != "all" and model_type != "lrmn":
        raise ValueError(f"{gene_selection_model=} is not a valid model")
    if noisemodel != "NegativeBinomial" or normalize:
        raise ValueError("the B200 path implements the NegativeBinomial noise model on raw integer counts")
    if hasattr(anndata, "var") and getattr(cycle_obj, "means", None) is not None and cycle_obj.means.columns is not None:
        cycle_obj, anndata = filter_shared_genes(cycle_obj, anndata, filter_type=behavior)   # preprocessing.py:246
    S = _counts_layer(anndata.layers["spliced"])
    U = _counts_layer(anndata.layers["unspliced"])
    cf = count_factor if isinstance(count_factor, torch.Tensor) else None
    mp = make_velocity_metaparams(
        S, U, cycle_obj.means_tensor.T, cycle_obj.stds_tensor.T, phase_obj.phi_xy_tensor.T,
        speed_obj.means_tensor.T, speed_obj.stds_tensor.T,
        batch_id=_ids_from_design(batch_design_mtx), cond_id=_ids_from_design(condition_design_mtx),
        Nb=int(np.asarray(batch_design_mtx).shape[-1]), Nx=int(np.asarray(condition_design_mtx).shape[-1]),
        count_factor=cf, n_harmonics=n_harmonics, ω_n_harmonics=ω_n_harmonics, with_delta_nu=with_delta_nu,
        model_type=model_type, μγ=μγ, σγ=σγ, μβ=μβ, σβ=σβ, μΔν=μΔν, σΔν=σΔν, gamma_alpha=gamma_alpha,
        gamma_beta=gamma_beta, rho_mean=rho_mean, rho_std=rho_std, rho_scale=rho_scale, rho_rank=rho_rank,
        device=device, cycle_prior=cycle_obj, phase_prior=phase_obj, speed_prior=speed_obj,
    )
    # the remaining fields of the reference's container (preprocessing.py:270-322); logS / logU as in the phase stage
    d = mp._asdict()
    if cf is not None:
        d["count_factor"] = cf.clone().detach().to(device)          # passed through in the caller's shape, like the reference
    d["ν"] = cycle_obj.means_tensor.T.unsqueeze(-2).to(device)
    d["condition"] = np.array(list(condition_on.keys()))
    return _container(d)
