#!/usr/bin/env python
"""Golden metaparameters of ``preprocess_for_phase_estimation`` / ``preprocess_for_velocity_estimation``, produced by
EXECUTING THE REFERENCE'S OWN ``preprocessing.py`` (released ``build/lib`` tree; Pyro / matplotlib / IPython stubbed exactly as
in ``generate_golden.py``; AnnData replaced by ``_fake_anndata.FakeAnnData``).  Build container only:

    python tests/golden/generate_preprocess_golden.py

Writes ``tests/golden/preprocess.npz``: the inputs (count layers with gene names in NON-alphabetical order and two genes the
Cycle prior does not know, priors, design matrices) and every tensor-valued field of the two metaparameter containers.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _fake_anndata import FakeAnnData  # noqa: E402
from generate_golden import import_reference  # noqa: E402


def main():
    vc = import_reference()
    from velocycle.angularspeed import AngularSpeed
    from velocycle.cycle import Cycle
    from velocycle.phases import Phases
    from velocycle.preprocessing import preprocess_for_phase_estimation, preprocess_for_velocity_estimation

    rng = np.random.default_rng(11)
    Nc, H, Hw = 17, 2, 1
    data_genes = ["ZEB1", "ACTB", "TOP2A", "E2F1", "MKI67", "GAPDH", "CCNB1"]     # not sorted; GAPDH, ACTB unknown to the prior
    cycle_genes = ["TOP2A", "ZEB1", "CCNB1", "E2F1", "MKI67"]
    cells = [f"cell{i}" for i in range(Nc)]
    S = rng.poisson(3.0, size=(Nc, len(data_genes))).astype(np.int64)
    U = rng.poisson(1.0, size=(Nc, len(data_genes))).astype(np.int64)
    K = 2 * H + 1
    cmeans, cstds = rng.normal(size=(K, len(cycle_genes))), rng.uniform(0.1, 1.0, size=(K, len(cycle_genes)))
    phixy = rng.normal(size=(2, Nc))
    smeans, sstds = np.array([[0.4, 0.3], [0.0, 0.1], [0.05, 0.0]]), np.array([[0.1, 0.1], [0.05, 0.05], [0.05, 0.05]])
    batch = torch.nn.functional.one_hot(torch.as_tensor(rng.integers(0, 3, size=Nc)), 3)
    cond = torch.nn.functional.one_hot(torch.as_tensor(rng.integers(0, 2, size=Nc)), 2)
    rec = dict(S=S, U=U, data_genes=np.array(data_genes), cycle_genes=np.array(cycle_genes), cells=np.array(cells),
               cycle_means=cmeans, cycle_stds=cstds, phixy=phixy, speed_means=smeans, speed_stds=sstds,
               batch=batch.numpy(), cond=cond.numpy())

    def record(tag, mp):
        for k, v in mp._asdict().items():
            if isinstance(v, torch.Tensor):
                rec[f"{tag}/{k}"] = v.detach().cpu().numpy()
            elif isinstance(v, (int, float, str)):
                rec[f"{tag}/{k}"] = np.array(v)
        rec[f"{tag}/cycle_prior_genes"] = np.array(list(mp.cycle_prior.genes))

    # phase stage: the data already restricted to the prior's genes, in the prior's order (as the tutorials do)
    cyc = Cycle.from_array(cmeans, cstds, cycle_genes)
    ph = Phases.from_array(phixy, cells)
    ad = FakeAnnData({"spliced": S, "unspliced": U}, data_genes, cells)[:, cycle_genes]
    mp = preprocess_for_phase_estimation(ad, cyc, ph, batch, n_harmonics=H, device=torch.device("cpu"))
    record("phase", mp)
    # velocity stage: the full data; filter_shared_genes intersects and SORTS the genes
    sp = AngularSpeed.from_array(smeans, sstds, ["ctrl", "treated"], Nhω=2 * Hw + 1)
    ad = FakeAnnData({"spliced": S, "unspliced": U}, data_genes, cells)
    mpv = preprocess_for_velocity_estimation(ad, cyc, ph, sp, cond, batch, n_harmonics=H, ω_n_harmonics=Hw,
                                             count_factor=mp.count_factor, device=torch.device("cpu"))
    record("velocity", mpv)
    np.savez(os.path.join(HERE, "preprocess.npz"), **rec)
    print("wrote preprocess.npz with", len(rec), "arrays;", sorted(k for k in rec if k.startswith("velocity/"))[:8], "...")


if __name__ == "__main__":
    main()
