"""``AngularSpeed``: Fourier coefficients of the angular speed omega(phi) per condition -- the container the velocity
preprocessing reads (``means_tensor``, ``stds_tensor``, ``conditions``: ``preprocessing.py:229-236``) and the velocity fit
driver fills (``velocity_inference_model.py:172-177``).  Same attributes, methods and CSV format as ``velocycle/angularspeed.py``.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from ._tables import CoefficientTables, coefficient_labels

__all__ = ["AngularSpeed"]


class AngularSpeed(CoefficientTables):
    _default_extension_std = 3.0

    @property
    def conditions(self):
        return list(self.means.columns)

    @staticmethod
    def _table(array, rows, condition_names) -> pd.DataFrame:
        """(Kw, Nx) or (Nx, Kw) -> (Kw, Nx) table; a 1-row table wraps the array as is (``angularspeed.py:276-299``)."""
        df = pd.DataFrame([array]) if len(rows) == 1 else pd.DataFrame(np.asarray(array).squeeze())
        transposed = len(df.index) != len(rows)
        if transposed:
            df = df.T
        df.index = rows
        if condition_names is not None:
            df.columns = condition_names
        return df

    @classmethod
    def from_array(cls, means_array, stds_array, condition_names=None, Nhω: int = 0) -> "AngularSpeed":
        """``Nhω`` = number of coefficients per condition (2 H_omega + 1)."""
        assert means_array.shape == stds_array.shape, "Shapes of the arrays must be equal"
        rows = coefficient_labels(max(int(Nhω), 1))
        out = cls()
        out.means = cls._table(means_array, rows, condition_names)
        out.stds = cls._table(stds_array, rows, condition_names)
        return out

    @classmethod
    def trivial_prior(cls, condition_names, harmonics: int = 1, means=0.0, stds=3.0) -> "AngularSpeed":
        """Constant term N(means, stds), every higher coefficient N(0, 0.05) (``angularspeed.py:310-353``)."""
        Kw = 2 * harmonics + 1
        rows = coefficient_labels(Kw)
        mu = np.array([means] + [0.0] * (Kw - 1), dtype=np.float32)[:, None]
        sd = np.array([stds] + [0.05] * (Kw - 1), dtype=np.float32)[:, None]
        out = cls()
        out.means = pd.DataFrame(np.broadcast_to(mu, (Kw, len(condition_names))).copy(), index=rows, columns=condition_names)
        out.stds = pd.DataFrame(np.broadcast_to(sd, (Kw, len(condition_names))).copy(), index=rows, columns=condition_names)
        return out
