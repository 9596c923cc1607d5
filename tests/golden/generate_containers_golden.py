#!/usr/bin/env python
"""Golden files of the result containers, produced by EXECUTING THE REFERENCE'S OWN CLASSES
(``/root/reference/velocycle/{cycle,phases,angularspeed}.py``; matplotlib / pyro replaced by empty stubs, as in
``generate_golden.py``).  Run in the build container only:

    python tests/golden/generate_containers_golden.py

Writes ``tests/golden/containers/``: the CSV files the reference's ``save`` methods write for seeded instances, and
``containers.npz`` with the tensors / derived quantities its properties return (means_tensor, stds_tensor, phis, directions,
concentrations, stds, trivial priors, the harmonics edits, invert_direction, rotate, shift_zero by phase).
``Cycle.shift_zero`` / ``Cycle.invert_direction`` are recorded through a plain-numpy restatement of the reference's loops,
because its chained pandas assignment is a silent no-op under the pandas version installed here.
"""
import copy
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "containers")
REF = "/root/reference/velocycle"


def load_reference():
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from velocycle_b200 import ppl
    from velocycle_b200.ppl import distributions

    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["pyro"] = ppl                      # Pyro is not installable offline: its restatement (Poisson, GammaPoisson)
    sys.modules["pyro.distributions"] = distributions
    pkg = types.ModuleType("velocycle")
    pkg.__path__ = [REF]
    sys.modules["velocycle"] = pkg
    mods = {}
    for m in ("utils", "cycle", "phases", "angularspeed"):
        spec = importlib.util.spec_from_file_location(f"velocycle.{m}", os.path.join(REF, m + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"velocycle.{m}"] = mod
        spec.loader.exec_module(mod)
        mods[m] = mod
    return mods


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    Cycle, Phases, AngularSpeed = ref["cycle"].Cycle, ref["phases"].Phases, ref["angularspeed"].AngularSpeed
    rng = np.random.default_rng(7)
    genes = [f"G{i}" for i in range(6)]
    cells = [f"cell_{i}" for i in range(9)]
    conds = ["ctrl", "treated"]
    rec = {}

    means, stds = rng.normal(size=(5, 6)), rng.uniform(0.1, 1.0, size=(5, 6))
    cyc = Cycle.from_array(means, stds, genes)
    cyc.save(os.path.join(OUT, "cycle.csv"))
    rec["cycle_in_means"], rec["cycle_in_stds"] = means, stds
    rec["cycle_means_tensor"], rec["cycle_stds_tensor"] = cyc.means_tensor.numpy(), cyc.stds_tensor.numpy()
    back = Cycle.load(os.path.join(OUT, "cycle.csv"))
    rec["cycle_loaded_means"], rec["cycle_loaded_stds"] = back.means.values, back.stds.values
    tp = Cycle.trivial_prior(genes, harmonics=2)
    rec["cycle_trivial2_means"], rec["cycle_trivial2_stds"] = tp.means.values, tp.stds.values
    tp3 = Cycle.trivial_prior(genes, harmonics=3, means=0.5, stds=2.0)
    rec["cycle_trivial3_means"], rec["cycle_trivial3_stds"] = tp3.means.values, tp3.stds.values
    ext = copy.deepcopy(cyc)  # (the reference's Cycle.copy() raises NameError: `copy` is never imported in cycle.py)
    ext.extend(["X1", "X2"])
    ext.add_harmonics(1)
    rec["cycle_edit_means"], rec["cycle_edit_stds"] = ext.means.values, ext.stds.values
    rec["cycle_edit_rows"] = np.array(list(ext.means.index))
    rec["cycle_edit_cols"] = np.array(list(ext.means.columns))
    ext.remove_harmonics(2)
    rec["cycle_removed_rows"] = np.array(list(ext.means.index))
    # check_orientation (cycle.py:423-447) indexes a string-labelled Series with `[2]`, a KeyError under this pandas: restated
    def orient(g1, g2):
        a = [np.arctan2(means[2, genes.index(g)], means[1, genes.index(g)]) for g in (g1, g2)]
        a = [x + 2 * np.pi if x < 0 else x for x in a]
        return (a[1] - a[0]) > 0
    rec["cycle_orientation"] = np.array([orient("G0", "G1"), orient("G2", "G5")])
    # the reference's rotation / inversion loops (cycle.py:407-421), restated on arrays
    M = means.copy()
    c, s = M[1:3, 2] / np.linalg.norm(M[1:3, 2])
    s = -s
    for i in range(1, 5, 2):
        c0, s0 = M[i].copy(), M[i + 1].copy()
        M[i], M[i + 1] = c0 * c - s0 * s, c0 * s + s0 * c
    rec["cycle_shift_gene2"] = M
    M = means.copy()
    M[[2, 4]] *= -1
    rec["cycle_inverted"] = M

    xy = rng.normal(size=(2, 9)) * 2.0
    ph = Phases.from_array(xy, cells)
    ph.save(os.path.join(OUT, "phases.csv"))
    rec["phases_in"] = xy
    rec["phases_tensor"] = ph.phi_xy_tensor.numpy()
    rec["phases_phis"] = ph.phis.numpy()
    rec["phases_directions"], rec["phases_concentrations"], rec["phases_stds"] = ph.directions, ph.concentrations, ph.stds
    rec["phases_loaded"] = Phases.load(os.path.join(OUT, "phases.csv")).phi_xy.values
    p2 = Phases.from_array(xy.copy(), cells)
    p2.rotate(0.7)
    rec["phases_rotated"] = p2.phi_xy.values
    p2.invert_direction()
    rec["phases_rot_inv"] = p2.phi_xy.values
    p3 = Phases.from_array(xy.copy(), cells)
    p3.shift_zero(phase=1.1)
    rec["phases_shifted"] = p3.phi_xy.values

    # grid-search MLE prior and max_corr on a small seeded data set drawn around a known cycle
    from types import SimpleNamespace
    import torch
    true_phi = rng.uniform(0, 2 * np.pi, size=40)
    cyc_mle = Cycle.from_array(np.vstack([rng.normal(0.5, 0.3, size=(1, 6)), rng.normal(0, 0.8, size=(2, 6))]), np.ones((3, 6)), genes)
    zeta = np.stack([np.ones(40), np.sin(true_phi), np.cos(true_phi)], 1)
    n_sc = rng.integers(50, 150, size=40).astype(float)
    lam = np.exp(zeta @ cyc_mle.means.values + 0.5 * np.log(n_sc)[:, None])
    S_mle = rng.poisson(lam).astype(np.int64)
    data = SimpleNamespace(obs=SimpleNamespace(n_scounts=SimpleNamespace(values=n_sc)), layers={"spliced": S_mle})
    rec["mle_means"], rec["mle_n_scounts"], rec["mle_S"] = cyc_mle.means.values, n_sc, S_mle
    for nm in ("Poisson", "NegativeBinomial"):
        pm = Phases.from_array(np.zeros((2, 40)), [f"c{i}" for i in range(40)])
        pm.from_cycle_mle(cyc_mle, data, a=0.5, bins=50, concentration=7.0, noisemodel=nm, dispersion=0.4)
        rec[f"mle_phixy_{nm}"] = pm.phi_xy.values
    pc = Phases.from_array(xy.copy(), cells)
    mc_in = rng.normal(size=9)
    sh, cbest, call = pc.max_corr(mc_in, npoints=20)
    rec["maxcorr_in"], rec["maxcorr_all"] = mc_in, np.array(call)
    rec["maxcorr_shift"], rec["maxcorr_best"] = np.array(sh), np.array(cbest)

    # PCA heuristic on a seeded ring-shaped data set
    ang = rng.uniform(0, 2 * np.pi, size=60)
    load = rng.normal(size=(2, 12))
    Lz = np.exp(1.0 + 0.8 * (np.stack([np.cos(ang), np.sin(ang)], 1) @ load) + 0.1 * rng.normal(size=(60, 12)))
    ad = SimpleNamespace(layers={"S_sz": Lz}, obs=SimpleNamespace(index=[f"k{i}" for i in range(60)]))
    rec["pca_layer"] = Lz
    for tag, kw in (("plain", {}), ("gap", dict(zero_at_min_density=True, concentration=3.0)), ("raw", dict(normalize_pcs=False))):
        rec[f"pca_phixy_{tag}"] = Phases.from_pca_heuristic(ad, **kw).phi_xy.values

    am, asd = rng.normal(size=(3, 2)), rng.uniform(0.05, 0.5, size=(3, 2))
    sp = AngularSpeed.from_array(am, asd, conds, Nhω=3)
    sp.save(os.path.join(OUT, "angularspeed.csv"))
    rec["speed_in_means"], rec["speed_in_stds"] = am, asd
    rec["speed_means_tensor"], rec["speed_stds_tensor"] = sp.means_tensor.numpy(), sp.stds_tensor.numpy()
    spT = AngularSpeed.from_array(am.T.copy(), asd.T.copy(), conds, Nhω=3)       # (Nx, Kw) input is transposed
    rec["speed_T_means"] = spT.means.values
    sp1 = AngularSpeed.from_array(np.array([0.3, 0.5]), np.array([0.1, 0.2]), conds, Nhω=1)   # constant speed
    rec["speed_const_means"], rec["speed_const_stds"] = sp1.means.values, sp1.stds.values
    tps = AngularSpeed.trivial_prior(conds, harmonics=1, means=0.4, stds=0.2)
    rec["speed_trivial_means"], rec["speed_trivial_stds"] = tps.means.values, tps.stds.values
    rec["speed_loaded_means"] = AngularSpeed.load(os.path.join(OUT, "angularspeed.csv")).means.values
    np.savez(os.path.join(OUT, "containers.npz"), **rec)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
