"""A minimal stand-in for anndata.AnnData (not installable offline) with exactly what the preprocessing reads: ``layers``
(name -> (Nc, Ng) array), ``var.index`` / ``obs.index``, ``shape``, ``data[:, genes]`` column selection by name or mask, ``copy``."""
import copy as _copy

import numpy as np
import pandas as pd


class FakeAnnData:
    def __init__(self, layers, var_names, obs_names):
        self.layers = {k: np.asarray(v) for k, v in layers.items()}
        self.var = pd.DataFrame(index=pd.Index(list(var_names)))
        self.obs = pd.DataFrame(index=pd.Index(list(obs_names)))

    @property
    def shape(self):
        return (len(self.obs.index), len(self.var.index))

    def __getitem__(self, key):
        rows, cols = key
        assert rows == slice(None), "only column selection is supported"
        cols = np.asarray(cols)
        if cols.dtype == bool:
            idx = np.nonzero(cols)[0]
        else:
            pos = {g: i for i, g in enumerate(self.var.index)}
            idx = np.array([pos[g] for g in cols], dtype=int)
        return FakeAnnData({k: v[:, idx] for k, v in self.layers.items()}, self.var.index[idx], self.obs.index)

    def copy(self):
        return _copy.deepcopy(self)
