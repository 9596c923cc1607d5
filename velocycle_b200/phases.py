"""``Phases``: per-cell phase estimates as (phi_x, phi_y) direction vectors -- the container the preprocessing reads
(``phi_xy_tensor``: ``preprocessing.py:130``) and the fit drivers fill (``phase_inference_model.py:199``).  Same attributes,
methods and CSV format as ``velocycle/phases.py`` for everything on or next to the SVI path, including the grid-search maximum-likelihood prior (``from_cycle_mle``) and
``max_corr`` and the PCA heuristic (plotting excluded).
"""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from .utils import pack_direction, torch_fourier_basis, unpack_direction

__all__ = ["Phases"]

_ROWS = ["phi_x", "phi_y"]


class Phases:
    def __init__(self):
        self.phi_xy = None   # pd.DataFrame (2, Nc): rows phi_x, phi_y; columns = cell names
        self.pcs = None
        self.omegas = None

    def __len__(self) -> int:
        return self.shape[-1]

    @property
    def shape(self):
        return self.phi_xy.shape

    def set_phixy(self, new_phixy) -> None:
        if isinstance(new_phixy, pd.DataFrame):
            self.phi_xy = new_phixy
            return
        if isinstance(new_phixy, torch.Tensor):
            new_phixy = new_phixy.detach().cpu().numpy()
        if not isinstance(new_phixy, np.ndarray):
            raise Exception("Error: invalid type for new_phixy")
        self.phi_xy = pd.DataFrame(new_phixy, index=self.phi_xy.index, columns=self.phi_xy.columns)

    def set_omegas(self, new_omegas) -> None:
        self.omegas = new_omegas

    # ---- views ----------------------------------------------------------------------------------------------
    @property
    def phi_xy_tensor(self) -> torch.Tensor:
        return torch.tensor(self.phi_xy.values.astype(np.float32))

    @property
    def phis(self) -> torch.Tensor:
        """Angles in [0, 2 pi) (``phases.py:176-186``)."""
        phis = pack_direction(self.phi_xy_tensor.T)
        phis[phis < 0] = phis[phis < 0] + 2 * np.pi
        return phis

    @property
    def directions(self) -> np.ndarray:
        return np.arctan2(self.phi_xy.values[1, :], self.phi_xy.values[0, :]) % (2 * np.pi)

    @property
    def concentrations(self) -> np.ndarray:
        return np.sqrt(np.sum(self.phi_xy.values ** 2, 0))

    @property
    def stds(self) -> np.ndarray:
        """Circular standard deviation of a von Mises with these concentrations: sqrt(1 - I1(k)/I0(k)) (``phases.py:218-234``;
        the Bessel ratio comes from the exponentially scaled torch.special functions instead of polynomial fits)."""
        k = torch.as_tensor(self.concentrations, dtype=torch.float64)
        return np.sqrt(1.0 - (torch.special.i1e(k) / torch.special.i0e(k)).numpy())

    # ---- files ----------------------------------------------------------------------------------------------
    @classmethod
    def load(cls, filepath) -> "Phases":
        out = cls()
        out.phi_xy = pd.read_csv(filepath, index_col=0)
        return out

    @classmethod
    def from_file(cls, filepath) -> "Phases":
        return cls.load(filepath)

    def save(self, pathname) -> None:
        self.phi_xy.to_csv(pathname)

    # ---- constructors ---------------------------------------------------------------------------------------
    @classmethod
    def from_array(cls, phi_xy_array, cell_names=None) -> "Phases":
        assert phi_xy_array.shape[0] == 2, "Shape of the array is incorrect"
        if cell_names is not None:
            assert len(cell_names) == phi_xy_array.shape[1]
        out = cls()
        out.phi_xy = pd.DataFrame(phi_xy_array, index=_ROWS, columns=cell_names)
        return out

    @classmethod
    def flat_prior(cls, anndata_object) -> "Phases":
        """All-zero direction vectors = no phase information (``phases.py:384-402``)."""
        out = cls()
        out.phi_xy = pd.DataFrame(np.zeros((2, anndata_object.shape[0])), index=_ROWS, columns=anndata_object.obs.index)
        return out

    @classmethod
    def from_pca_heuristic(cls, anndata_object, genes_to_use=None, concentration=1.0, layer="S_sz", small_count=1.0e-1,
                           normalize_pcs=True, zero_at_min_density=False, random_state=0, plot=False, n_components=2) -> "Phases":
        """Phase = angle in the plane of the first two principal components of log(layer + small_count)
        (``phases.py:307-383``): optional robust scaling of the PCs by their 0.5 / 99.5 percentiles around the median, and
        optional zero at the widest gap of the angle distribution.  ``plot`` is accepted and ignored (no plotting here)."""
        from sklearn.decomposition import PCA

        if layer not in anndata_object.layers:
            raise ValueError(f"{layer=} is not a valid entry anndata.obs")
        sub = anndata_object if genes_to_use is None else anndata_object[:, [g in genes_to_use for g in anndata_object.var.index]]
        L = sub.layers[layer]
        L = L.toarray() if hasattr(L, "toarray") else np.asarray(L)
        X = np.log(L + small_count)                                       # (Nc, Ng)
        pca = PCA(n_components, random_state=random_state)
        pcs = pca.fit_transform(X)
        if normalize_pcs:
            lo, hi, med = np.percentile(pcs, [0.5, 99.5, 50], 0)
            pcs = (pcs - med) / (hi - lo)
        angle = np.arctan2(pcs[:, 1], pcs[:, 0]) % (2 * np.pi)
        if zero_at_min_density:
            order = np.argsort(angle)
            start = order[np.diff(angle[order]).argmax() + 1]
            angle = (angle - angle[start]) % (2 * np.pi)
        out = cls()
        out.phi_xy = pd.DataFrame(np.vstack([np.cos(angle), np.sin(angle)]) * concentration, index=_ROWS,
                                  columns=anndata_object.obs.index)
        out.pcs, out.pca = pcs, pca
        return out

    # ---- gauge ----------------------------------------------------------------------------------------------
    def shift_zero(self, gene=None, phase=None) -> None:
        """Subtract ``phase`` from every angle; the result has unit concentration (``phases.py:404-421``)."""
        if gene is not None:
            raise Exception("Error: must phase for desired shift")
        if phase is None:
            raise Exception("Error: must specify gene or phase for desired shift")
        self.set_phixy(unpack_direction(self.phis - phase).T)

    def rotate(self, angle=None) -> None:
        if angle is None:
            raise Exception("Error: must specify angle for desired rotation")
        c, s = np.cos(angle), np.sin(angle)
        self.set_phixy(np.matmul(np.array([[c, -s], [s, c]]), self.phi_xy.values))

    def invert_direction(self) -> None:
        self.set_phixy(np.matmul(np.array([[1.0, 0.0], [0.0, -1.0]]), self.phi_xy.values))

    # ---- data-driven phases -----------------------------------------------------------------------------------
    def max_corr(self, counts, npoints: int = 100):
        """Shift in [0, 2 pi) that maximises the correlation of the (shifted, re-wrapped) phases with ``counts``
        (``phases.py:450-469``).  Returns (shift, correlation, all correlations)."""
        shifts = np.arange(0, npoints) / npoints * 2 * np.pi
        base = self.phis                                  # float32, like the reference's arithmetic
        corr = []
        for sh in shifts:
            x = base - sh
            x[x < 0] = x[x < 0] + 2 * np.pi
            corr.append(np.corrcoef(x.numpy(), counts)[0, 1])
        best = int(np.argmax(np.array(corr)))
        return shifts[best], corr[best], corr

    def from_cycle_mle(self, cycle, data, a=1, bins: int = 100, concentration: float = 10.0, noisemodel: str = "Poisson",
                       dispersion: float = 0.3, device=None, bins_per_pass: int = 8) -> None:
        """Grid-search maximum-likelihood phase per cell given a Cycle (``phases.py:471-509``): for ``bins`` phases on a
        regular grid, log P[bin, cell] = sum_g log p(S[c, g] | exp(nu_g . zeta(phi_bin) + a log n_scounts_c)) under a Poisson or
        negative-binomial (``GammaPoisson(1/dispersion, 1/(dispersion mu))``) model; every cell takes the arg-max phase with the
        given concentration.  The (bins, Ng, Nc) scan is evaluated ``bins_per_pass`` phases at a time on ``device`` (default: CPU;
        pass a CUDA device for large data) -- the reference materialises it whole."""
        if noisemodel not in ("Poisson", "NegativeBinomial"):
            raise NotImplementedError("Not implemented yet, sorry")
        dev = torch.device("cpu") if device is None else torch.device(device)
        fou = cycle.means_tensor.to(dev)                                    # (K, Ng)
        H = (fou.shape[0] - 1) // 2
        log_counts = torch.tensor(np.log(data.obs.n_scounts.values), dtype=torch.float32, device=dev)
        offset = log_counts * torch.as_tensor(a, device=dev)                # (Nc,)
        layer = data.layers["spliced"]
        layer = layer.toarray() if hasattr(layer, "toarray") else np.asarray(layer)
        k = torch.as_tensor(layer.astype(np.int64), device=dev).T.to(torch.float32)   # (Ng, Nc)
        grid = 2 * np.pi * torch.arange(0, 1, 1.0 / bins, dtype=torch.float32)
        curves = torch.matmul(torch_fourier_basis(grid, num_harmonics=H).to(dev), fou)  # (bins, Ng)
        lgk1 = torch.lgamma(k + 1.0)
        r = 1.0 / dispersion
        best_lp = torch.full((k.shape[1],), -float("inf"), device=dev)
        best_bin = torch.zeros(k.shape[1], dtype=torch.long, device=dev)
        for b0 in range(0, grid.shape[0], bins_per_pass):
            eta = curves[b0: b0 + bins_per_pass, :, None] + offset[None, None, :]       # (b, Ng, Nc)
            if noisemodel == "Poisson":
                lp = k * eta - torch.exp(eta) - lgk1
            else:  # NB with total_count r and mean mu = exp(eta): log-pmf of GammaPoisson(r, r / mu)
                mu = torch.exp(eta)
                lp = (torch.lgamma(k + r) - lgk1 - torch.lgamma(torch.as_tensor(r, device=dev))
                      + r * (np.log(r) - torch.log(r + mu)) + k * (eta - torch.log(r + mu)))
            tot = lp.sum(1)                                                             # (b, Nc)
            val, idx = tot.max(0)
            better = val > best_lp                                                      # strict: first maximum wins, like argmax
            best_lp = torch.where(better, val, best_lp)
            best_bin = torch.where(better, idx + b0, best_bin)
        phis_mle = grid[best_bin.cpu()]
        self.set_phixy((concentration * unpack_direction(phis_mle).T).numpy())
