"""``preprocess_for_phase_estimation`` / ``preprocess_for_velocity_estimation`` against the metaparameter containers the
REFERENCE's own preprocessing produced on the same inputs (tests/golden/generate_preprocess_golden.py): every tensor field with
the reference's shape and values, the gene intersection / alphabetical ordering of the velocity stage, scalar fields.
``logS`` / ``logU`` (dense (Ng, Nc) log(count + 1) matrices for the Lognormal variants and plotting) are deliberately absent."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from _fake_anndata import FakeAnnData  # noqa: E402

from velocycle_b200.angularspeed import AngularSpeed  # noqa: E402
from velocycle_b200.cycle import Cycle  # noqa: E402
from velocycle_b200.phases import Phases  # noqa: E402
from velocycle_b200.preprocessing import (filter_shared_genes, preprocess_for_phase_estimation,  # noqa: E402
                                           preprocess_for_velocity_estimation)

ABSENT = {"logS", "logU"}


@pytest.fixture(scope="module")
def setup():
    z = np.load(os.path.join(HERE, "golden", "preprocess.npz"))
    data_genes, cycle_genes, cells = list(z["data_genes"]), list(z["cycle_genes"]), list(z["cells"])
    cyc = Cycle.from_array(z["cycle_means"], z["cycle_stds"], cycle_genes)
    ph = Phases.from_array(z["phixy"], cells)
    sp = AngularSpeed.from_array(z["speed_means"], z["speed_stds"], ["ctrl", "treated"], Nhω=3)
    ad = FakeAnnData({"spliced": z["S"], "unspliced": z["U"]}, data_genes, cells)
    return z, cyc, ph, sp, ad, torch.as_tensor(z["batch"]), torch.as_tensor(z["cond"]), cycle_genes


def _check(z, tag, mp):
    d = mp._asdict()
    n = 0
    for key in z.files:
        if not key.startswith(tag + "/"):
            continue
        f = key.split("/", 1)[1]
        if f in ABSENT:
            assert f not in d
            continue
        ref = z[key]
        if f == "cycle_prior_genes":
            assert list(ref) == list(d["cycle_prior"].genes)
        elif isinstance(d[f], torch.Tensor):
            assert tuple(d[f].shape) == tuple(ref.shape), (f, tuple(d[f].shape), ref.shape)
            assert np.allclose(d[f].cpu().numpy().astype(np.float64), ref.astype(np.float64), rtol=0, atol=1e-6), f
        else:
            assert np.array(d[f]).item() == ref.item(), f
        n += 1
    return n


def test_phase_metaparameters_match_the_reference(setup):
    z, cyc, ph, sp, ad, batch, cond, cycle_genes = setup
    mp = preprocess_for_phase_estimation(ad[:, cycle_genes], cyc, ph, batch, n_harmonics=2, device=torch.device("cpu"))
    assert _check(z, "phase", mp) >= 20
    assert mp.model_fn.__name__ == "phase_latent_variable_model" and mp.guide_fn.__name__ == "phase_latent_variable_guide"
    assert tuple(mp.count_factor.shape) == (1, 1, len(ph))


def test_velocity_metaparameters_match_the_reference(setup):
    z, cyc, ph, sp, ad, batch, cond, cycle_genes = setup
    mp0 = preprocess_for_phase_estimation(ad[:, cycle_genes], cyc, ph, batch, n_harmonics=2, device=torch.device("cpu"))
    mp = preprocess_for_velocity_estimation(ad, cyc, ph, sp, cond, batch, n_harmonics=2, ω_n_harmonics=1,
                                            count_factor=mp0.count_factor, device=torch.device("cpu"))
    assert _check(z, "velocity", mp) >= 30
    assert mp.cycle_prior.genes == sorted(cycle_genes) and mp.Ng == 5            # intersection, alphabetical
    assert mp.model_fn.__name__ == "velocity_latent_variable_model_LRMN"


def test_filter_shared_genes(setup):
    z, cyc, ph, sp, ad, batch, cond, cycle_genes = setup
    c2, d2 = filter_shared_genes(cyc, ad, "intersection")
    assert c2.genes == sorted(cycle_genes) == list(d2.var.index)
    assert np.array_equal(c2.means["TOP2A"].values, cyc.means["TOP2A"].values)
    assert np.array_equal(d2.layers["spliced"][:, 0], ad.layers["spliced"][:, list(ad.var.index).index("CCNB1")])
    c3, d3 = filter_shared_genes(cyc, ad, "union")
    assert c3.genes == sorted(ad.var.index) == list(d3.var.index) and len(c3) == 7
    # new genes get trivial_prior entries: means 0 and -- for 1 or 2 harmonics -- the fixed stds (.1,.2,.2,.1,.1) of
    # cycle.py:341-344, which override the stds=10 that extend() asks for (reference behaviour, kept)
    assert float(c3.means["ACTB"].abs().sum()) == 0.0 and list(c3.stds["GAPDH"].values) == [0.1, 0.2, 0.2, 0.1, 0.1]
    assert np.array_equal(c3.means["E2F1"].values, cyc.means["E2F1"].values)
    with pytest.raises(Exception):
        filter_shared_genes(cyc, ad[:, ["ACTB", "TOP2A"]], "union")              # a Cycle gene missing from the data
    with pytest.raises(Exception):
        filter_shared_genes(cyc, ad, "nope")
