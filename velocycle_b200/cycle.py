"""``Cycle``: the gene-wise Fourier coefficients of the cell-cycle manifold (prior and fitted), the container the
preprocessing reads (``means_tensor``, ``stds_tensor``, ``genes``: ``preprocessing.py:128-137``) and the fit drivers fill
(``phase_inference_model.py:189-205``, ``velocity_inference_model.py:156-186``).  Same attributes, methods and CSV format as
``velocycle/cycle.py`` (plotting excluded: SURVEY.md section 2); own implementation on top of ``_tables.CoefficientTables``.
"""
from __future__ import annotations

from math import atan2

import numpy as np
import pandas as pd

from ._tables import CoefficientTables, coefficient_labels

__all__ = ["Cycle", "reorder"]


class Cycle(CoefficientTables):
    _default_extension_std = 10.0

    def __init__(self):
        super().__init__()
        self.log_gammas = None  # log degradation rates (velocity fit)
        self.log_betas = None   # log splicing rates (velocity fit)
        self.disp_pyro = None   # negative-binomial dispersions shape_inv
        self.periodic = None

    def set_log_gammas(self, new_gammas) -> None:
        self.log_gammas = new_gammas

    def set_log_betas(self, new_betas) -> None:
        self.log_betas = new_betas

    def set_disp_pyro(self, new_disp_pyro) -> None:
        self.disp_pyro = new_disp_pyro

    @property
    def genes(self):
        return list(self.means.columns)

    @classmethod
    def from_array(cls, means_array, stds_array, gene_names=None) -> "Cycle":
        """(K, Ng) arrays -> tables labelled ``nu0, nu1_cos, ...`` x genes (``cycle.py:301-326``)."""
        assert means_array.shape == stds_array.shape, "Shapes of the arrays must be equal"
        if gene_names is not None:
            assert len(gene_names) == means_array.shape[1]
        rows = coefficient_labels(means_array.shape[0])
        out = cls()
        out.means = pd.DataFrame(means_array, index=rows, columns=gene_names)
        out.stds = pd.DataFrame(stds_array, index=rows, columns=gene_names)
        return out

    @classmethod
    def trivial_prior(cls, gene_names, harmonics: int = 2, means=0.0, stds=3.0) -> "Cycle":
        """Uninformative prior; for 1 or 2 harmonics the reference fixes the stds to (.1,.2,.2[,.1,.1]) (``cycle.py:341-344``)."""
        if harmonics == 1:
            stds = np.array([0.1, 0.2, 0.2])[:, None]
        if harmonics == 2:
            stds = np.array([0.1, 0.2, 0.2, 0.1, 0.1])[:, None]
        K = 2 * harmonics + 1
        rows = coefficient_labels(K)
        out = cls()
        out.means = pd.DataFrame(np.broadcast_to(means, (K, len(gene_names))).copy(), index=rows, columns=gene_names)
        out.stds = pd.DataFrame(np.broadcast_to(stds, (K, len(gene_names))).copy(), index=rows, columns=gene_names)
        return out

    # ---- gauge: where phase zero sits and which way the cycle runs -----------------------------------------------
    def shift_zero(self, gene=None, phase=None) -> None:
        """Rotate every harmonic so that ``gene`` (or ``phase``) sits at phase zero (``cycle.py:393-413``; the same rotation
        angle is applied to every harmonic, as the reference does).  Works on a fresh array and assigns it back: the
        reference's chained ``self.means[g].iloc[...] = ...`` silently does nothing under pandas copy-on-write."""
        if gene is not None:
            if gene not in self.means.keys():
                raise Exception("Error: gene not found in index")
            c, s = self.means[gene].iloc[1:3].values
            c, s = np.array([c, s]) / np.linalg.norm([c, s])
        elif phase is not None:
            c, s = np.cos(phase), np.sin(phase)
        else:
            raise Exception("Error: must specify gene or phase for desired shift")
        s = -s
        M = self.means.values.astype(float).copy()
        for i in range(1, 2 * self.harmonics + 1, 2):
            c0, s0 = M[i].copy(), M[i + 1].copy()
            M[i], M[i + 1] = c0 * c - s0 * s, c0 * s + s0 * c
        self.means = pd.DataFrame(M, index=self.means.index, columns=self.means.columns)

    def invert_direction(self) -> None:
        """Flip the sign of rows 2, 4, ... (``cycle.py:415-421``)."""
        M = self.means.values.astype(float).copy()
        M[2 * (1 + np.arange(0, self.harmonics))] *= -1.0
        self.means = pd.DataFrame(M, index=self.means.index, columns=self.means.columns)

    def check_orientation(self, gene_pair=("TOP2A", "E2F1")) -> bool:
        g1, g2 = gene_pair
        if g1 not in self.means.keys() or g2 not in self.means.keys():
            raise Exception("Error: invalid gene names")
        ang = []
        for g in (g1, g2):
            a = atan2(self.means[g].iloc[2], self.means[g].iloc[1])
            ang.append(a + 2 * np.pi if a < 0 else a)
        return (ang[1] - ang[0]) > 0


def reorder(cycle: Cycle, gene_list) -> Cycle:
    """A new Cycle with its genes in the order of ``gene_list`` (``cycle.py:449-465``)."""
    return Cycle.from_array(means_array=cycle.means[gene_list], stds_array=cycle.stds[gene_list])
