"""Shared machinery of the result containers (`Cycle`, `AngularSpeed`): a pair of pandas tables (means, stds) whose rows are
Fourier coefficients and whose columns are genes / conditions.

Row labels follow the reference's files (``cycle.py:312-314``, ``angularspeed.py:273-275``): ``nu0, nu1_cos, nu1_sin, nu2_cos,
...``.  NOTE the labels do not name the basis columns (``utils.py:420-435`` orders them [1, sin, cos, sin 2, cos 2, ...]): row
2n-1 is the sin coefficient although it is labelled ``nu{n}_cos``.  Files written by the reference carry these labels, so they
are kept verbatim; the arithmetic below (rotations, inversions) works on row positions, like the reference's.
"""
from __future__ import annotations

import copy as _copy
from typing import List, Optional, Sequence

import numpy as np
import pandas as pd
import torch

__all__ = ["coefficient_labels", "CoefficientTables"]


def coefficient_labels(n_rows: int) -> List[str]:
    """``["nu0", "nu1_cos", "nu1_sin", "nu2_cos", ...]`` for ``n_rows`` coefficients."""
    return ["nu0"] + [f"nu{i // 2 + 1}_{'sin' if i % 2 else 'cos'}" for i in range(n_rows - 1)]


def _as_frame(values, like: pd.DataFrame) -> pd.DataFrame:
    if isinstance(values, pd.DataFrame):
        return values
    if isinstance(values, torch.Tensor):
        values = values.detach().cpu().numpy()
    if isinstance(values, np.ndarray):
        return pd.DataFrame(values, index=like.index, columns=like.columns)
    raise Exception("Error: invalid type for the new coefficient table")


class CoefficientTables:
    """means / stds tables with the file format ``pd.concat([means, stds]).to_csv`` (first column = row labels)."""

    means: Optional[pd.DataFrame]
    stds: Optional[pd.DataFrame]
    _default_extension_std = 10.0

    def __init__(self):
        self.means = None
        self.stds = None

    # ---- basic protocol -------------------------------------------------------------------------------------
    def __len__(self) -> int:
        return self.shape[-1]

    def __getitem__(self, key):
        out = type(self)()
        out.means = self.means.__getitem__(key)
        out.stds = self.stds.__getitem__(key)
        return out

    @property
    def harmonics(self) -> int:
        return (self.means.shape[0] - 1) // 2

    @property
    def shape(self):
        return self.means.shape

    @property
    def means_tensor(self) -> torch.Tensor:
        return torch.tensor(self.means.values.astype(np.float32))

    @property
    def stds_tensor(self) -> torch.Tensor:
        return torch.tensor(self.stds.values.astype(np.float32))

    def set_means(self, new_means) -> None:
        self.means = _as_frame(new_means, self.means)

    def set_stds(self, new_stds) -> None:
        self.stds = _as_frame(new_stds, self.stds)

    def copy(self):
        return _copy.deepcopy(self)

    # ---- files ----------------------------------------------------------------------------------------------
    @classmethod
    def load(cls, filepath):
        both = pd.read_csv(filepath, index_col=0)
        half = both.shape[0] // 2
        out = cls()
        out.means, out.stds = both.iloc[:half, :], both.iloc[half:, :]
        return out

    @classmethod
    def from_file(cls, filepath):
        return cls.load(filepath)

    def save(self, pathname) -> None:
        pd.concat([self.means, self.stds]).to_csv(pathname)

    # ---- growing / shrinking --------------------------------------------------------------------------------
    def extend(self, gene_names: Sequence[str], means=0.0, stds=None) -> None:
        """Append columns filled with ``means`` / ``stds`` (in place).  The parameter is called ``gene_names`` in both of the
        reference's classes (``cycle.py:200``, ``angularspeed.py:157``), also where the columns are conditions."""
        stds = self._default_extension_std if stds is None else stds
        extra = self.trivial_prior(gene_names, harmonics=self.harmonics, means=means, stds=stds)
        self.means = pd.concat([self.means, extra.means], axis=1)
        self.stds = pd.concat([self.stds, extra.stds], axis=1)

    def add_harmonics(self, extra_harmonics: int = 1, means=None, stds=None) -> None:
        n0, ncol = int(self.harmonics), self.shape[1]
        m = None if means is None else np.broadcast_to(means, (2 * extra_harmonics, ncol)).copy()
        s = None if stds is None else np.broadcast_to(stds, (2 * extra_harmonics, ncol)).copy()
        for i in range(extra_harmonics):
            for j, part in enumerate(("cos", "sin")):
                label = f"nu{n0 + 1 + i}_{part}"
                self.means.loc[label] = np.zeros(ncol) if m is None else m[2 * i + j]
                self.stds.loc[label] = 10 * np.ones(ncol) if s is None else s[2 * i + j]

    def remove_harmonics(self, n: int = 1) -> None:
        """Drops the last ``n`` ROWS (the reference's behaviour, ``cycle.py:242-250``: one harmonic = two rows)."""
        self.means = self.means.iloc[:-n, :]
        self.stds = self.stds.iloc[:-n, :]
