"""Result containers (Cycle, Phases, AngularSpeed) against golden files produced by executing the reference's own classes
(tests/golden/generate_containers_golden.py): CSV formats byte for byte, tensors, priors, edits and gauge transformations."""
import os

import numpy as np
import pytest
import torch

from velocycle_b200.angularspeed import AngularSpeed
from velocycle_b200.cycle import Cycle, reorder
from velocycle_b200.phases import Phases

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "containers")
GENES = [f"G{i}" for i in range(6)]
CELLS = [f"cell_{i}" for i in range(9)]
CONDS = ["ctrl", "treated"]


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(G, "containers.npz"), allow_pickle=False)


def _same_file(tmp_path, name, obj):
    out = tmp_path / name
    obj.save(out)
    assert out.read_bytes() == open(os.path.join(G, name), "rb").read(), f"{name}: the CSV differs from the reference's"


def test_cycle_tables_files_and_priors(gold, tmp_path):
    c = Cycle.from_array(gold["cycle_in_means"], gold["cycle_in_stds"], GENES)
    assert c.genes == GENES and c.harmonics == 2 and c.shape == (5, 6) and len(c) == 6
    assert list(c.means.index) == ["nu0", "nu1_cos", "nu1_sin", "nu2_cos", "nu2_sin"]
    assert np.array_equal(c.means_tensor.numpy(), gold["cycle_means_tensor"])
    assert np.array_equal(c.stds_tensor.numpy(), gold["cycle_stds_tensor"])
    _same_file(tmp_path, "cycle.csv", c)
    back = Cycle.from_file(os.path.join(G, "cycle.csv"))
    assert np.array_equal(back.means.values, gold["cycle_loaded_means"]) and np.array_equal(back.stds.values, gold["cycle_loaded_stds"])
    assert back.genes == GENES
    t2 = Cycle.trivial_prior(GENES, harmonics=2)
    assert np.array_equal(t2.means.values, gold["cycle_trivial2_means"]) and np.array_equal(t2.stds.values, gold["cycle_trivial2_stds"])
    t3 = Cycle.trivial_prior(GENES, harmonics=3, means=0.5, stds=2.0)
    assert np.array_equal(t3.means.values, gold["cycle_trivial3_means"]) and np.array_equal(t3.stds.values, gold["cycle_trivial3_stds"])
    sub = c[["G1", "G4"]]
    assert sub.genes == ["G1", "G4"] and np.array_equal(sub.means.values, gold["cycle_in_means"][:, [1, 4]])


def test_cycle_edits_and_setters(gold):
    c = Cycle.from_array(gold["cycle_in_means"], gold["cycle_in_stds"], GENES)
    e = c.copy()
    e.extend(["X1", "X2"])
    e.add_harmonics(1)
    assert np.array_equal(e.means.values, gold["cycle_edit_means"]) and np.array_equal(e.stds.values, gold["cycle_edit_stds"])
    assert list(e.means.index) == list(gold["cycle_edit_rows"]) and list(e.means.columns) == list(gold["cycle_edit_cols"])
    e.remove_harmonics(2)
    assert list(e.means.index) == list(gold["cycle_removed_rows"])
    assert np.array_equal(c.means.values, gold["cycle_in_means"])          # copy() is deep
    c.set_means(torch.zeros(5, 6))
    c.set_stds(np.ones((5, 6)))
    assert float(c.means.values.sum()) == 0.0 and c.genes == GENES and float(c.stds.values.mean()) == 1.0
    with pytest.raises(Exception):
        c.set_means([1, 2, 3])
    c.set_log_gammas(np.arange(6.0))
    c.set_log_betas(np.arange(6.0) + 1)
    c.set_disp_pyro(np.arange(6.0) + 2)
    assert c.log_gammas[5] == 5 and c.log_betas[0] == 1 and c.disp_pyro[0] == 2 and c.periodic is None


def test_cycle_gauge(gold):
    c = Cycle.from_array(gold["cycle_in_means"].copy(), gold["cycle_in_stds"], GENES)
    assert [bool(c.check_orientation(["G0", "G1"])), bool(c.check_orientation(["G2", "G5"]))] == list(gold["cycle_orientation"])
    with pytest.raises(Exception):
        c.check_orientation(["G0", "nope"])
    s = c.copy()
    s.shift_zero(gene="G2")
    assert np.allclose(s.means.values, gold["cycle_shift_gene2"], rtol=0, atol=1e-15)
    assert abs(s.means["G2"].iloc[2]) < 1e-12                              # the gene now sits at phase zero
    p = c.copy()
    p.shift_zero(phase=0.0)
    assert np.allclose(p.means.values, gold["cycle_in_means"])
    with pytest.raises(Exception):
        c.copy().shift_zero()
    with pytest.raises(Exception):
        c.copy().shift_zero(gene="nope")
    i = c.copy()
    i.invert_direction()
    assert np.array_equal(i.means.values, gold["cycle_inverted"])
    r = reorder(c, ["G3", "G0"])
    assert r.genes == ["G3", "G0"] and np.array_equal(r.means.values, gold["cycle_in_means"][:, [3, 0]])


def test_phases(gold, tmp_path):
    p = Phases.from_array(gold["phases_in"], CELLS)
    assert p.shape == (2, 9) and len(p) == 9 and list(p.phi_xy.index) == ["phi_x", "phi_y"]
    assert np.array_equal(p.phi_xy_tensor.numpy(), gold["phases_tensor"])
    assert np.array_equal(p.phis.numpy(), gold["phases_phis"])
    assert np.array_equal(p.directions, gold["phases_directions"]) and np.array_equal(p.concentrations, gold["phases_concentrations"])
    assert np.allclose(p.stds, gold["phases_stds"], rtol=0, atol=2e-6)      # (polynomial Bessel fits in the reference)
    _same_file(tmp_path, "phases.csv", p)
    assert np.array_equal(Phases.load(os.path.join(G, "phases.csv")).phi_xy.values, gold["phases_loaded"])
    q = Phases.from_array(gold["phases_in"].copy(), CELLS)
    q.rotate(0.7)
    assert np.array_equal(q.phi_xy.values, gold["phases_rotated"])
    q.invert_direction()
    assert np.array_equal(q.phi_xy.values, gold["phases_rot_inv"])
    s = Phases.from_array(gold["phases_in"].copy(), CELLS)
    s.shift_zero(phase=1.1)
    assert np.array_equal(s.phi_xy.values, gold["phases_shifted"]) and list(s.phi_xy.columns) == CELLS
    for bad in (dict(gene="G0"), dict()):
        with pytest.raises(Exception):
            s.shift_zero(**bad)
    with pytest.raises(Exception):
        s.rotate()
    with pytest.raises(AssertionError):
        Phases.from_array(np.zeros((3, 4)))

    class _Ad:  # the two attributes flat_prior reads from an AnnData
        shape = (4, 10)

        class obs:
            index = ["a", "b", "c", "d"]

    f = Phases.flat_prior(_Ad)
    assert f.shape == (2, 4) and float(np.abs(f.phi_xy.values).sum()) == 0.0 and list(f.phi_xy.columns) == ["a", "b", "c", "d"]
    s.set_omegas(np.ones(9))
    assert s.omegas.sum() == 9


def test_angular_speed(gold, tmp_path):
    a = AngularSpeed.from_array(gold["speed_in_means"], gold["speed_in_stds"], CONDS, Nhω=3)
    assert a.conditions == CONDS and a.harmonics == 1 and a.shape == (3, 2)
    assert np.array_equal(a.means_tensor.numpy(), gold["speed_means_tensor"]) and np.array_equal(a.stds_tensor.numpy(), gold["speed_stds_tensor"])
    _same_file(tmp_path, "angularspeed.csv", a)
    assert np.array_equal(AngularSpeed.load(os.path.join(G, "angularspeed.csv")).means.values, gold["speed_loaded_means"])
    aT = AngularSpeed.from_array(gold["speed_in_means"].T.copy(), gold["speed_in_stds"].T.copy(), CONDS, Nhω=3)
    assert np.array_equal(aT.means.values, gold["speed_T_means"]) and aT.conditions == CONDS
    c = AngularSpeed.from_array(np.array([0.3, 0.5]), np.array([0.1, 0.2]), CONDS, Nhω=1)
    assert np.array_equal(c.means.values, gold["speed_const_means"]) and np.array_equal(c.stds.values, gold["speed_const_stds"])
    t = AngularSpeed.trivial_prior(CONDS, harmonics=1, means=0.4, stds=0.2)
    assert np.array_equal(t.means.values, gold["speed_trivial_means"]) and np.array_equal(t.stds.values, gold["speed_trivial_stds"])
    t.extend(["third"])
    assert t.conditions == CONDS + ["third"] and float(t.stds["third"].iloc[0]) == 3.0


def test_phases_from_cycle_mle_and_max_corr(gold):
    """The grid-search maximum-likelihood prior (phases.py:471-509) and max_corr against the reference's own run."""
    from types import SimpleNamespace

    cyc = Cycle.from_array(gold["mle_means"], np.ones((3, 6)), GENES)
    data = SimpleNamespace(obs=SimpleNamespace(n_scounts=SimpleNamespace(values=gold["mle_n_scounts"])),
                           layers={"spliced": gold["mle_S"]})
    for nm in ("Poisson", "NegativeBinomial"):
        p = Phases.from_array(np.zeros((2, 40)), [f"c{i}" for i in range(40)])
        p.from_cycle_mle(cyc, data, a=0.5, bins=50, concentration=7.0, noisemodel=nm, dispersion=0.4, bins_per_pass=7)
        assert np.array_equal(p.phi_xy.values, gold[f"mle_phixy_{nm}"]), nm          # same arg-max bin for every cell
        assert np.allclose(p.concentrations, 7.0, atol=1e-5)
    import scipy.sparse as sp
    data_sp = SimpleNamespace(obs=data.obs, layers={"spliced": sp.csr_matrix(gold["mle_S"])})   # sparse layer
    q = Phases.from_array(np.zeros((2, 40)), [f"c{i}" for i in range(40)])
    q.from_cycle_mle(cyc, data_sp, a=0.5, bins=50, concentration=7.0)
    assert np.array_equal(q.phi_xy.values, gold["mle_phixy_Poisson"])
    with pytest.raises(NotImplementedError):
        q.from_cycle_mle(cyc, data, noisemodel="Lognormal")
    m = Phases.from_array(gold["phases_in"].copy(), CELLS)
    shift, best, allc = m.max_corr(gold["maxcorr_in"], npoints=20)
    assert shift == float(gold["maxcorr_shift"]) and np.allclose(allc, gold["maxcorr_all"], rtol=0, atol=1e-12)
    assert np.isclose(best, float(gold["maxcorr_best"]), rtol=0, atol=1e-12)


def test_phases_from_pca_heuristic(gold):
    from types import SimpleNamespace

    ad = SimpleNamespace(layers={"S_sz": gold["pca_layer"]}, obs=SimpleNamespace(index=[f"k{i}" for i in range(60)]))
    for tag, kw in (("plain", {}), ("gap", dict(zero_at_min_density=True, concentration=3.0)), ("raw", dict(normalize_pcs=False))):
        p = Phases.from_pca_heuristic(ad, **kw)
        assert np.allclose(p.phi_xy.values, gold[f"pca_phixy_{tag}"], rtol=0, atol=1e-12), tag
        assert p.pcs.shape == (60, 2) and list(p.phi_xy.columns)[:2] == ["k0", "k1"]
    with pytest.raises(ValueError):
        Phases.from_pca_heuristic(ad, layer="nope")


def test_circular_corrcoef():
    from velocycle_b200.utils import circular_corrcoef

    rng = np.random.default_rng(0)
    a = rng.uniform(0, 2 * np.pi, 500)
    assert abs(circular_corrcoef(a, a) - 1.0) < 1e-12
    assert abs(circular_corrcoef(a, (a + 1.234) % (2 * np.pi)) - 1.0) < 1e-12        # invariant to a common rotation
    assert circular_corrcoef(a, rng.uniform(0, 2 * np.pi, 500)) < 0.15
    ref = np.abs(np.mean(np.exp(1j * a) * np.conj(np.exp(1j * (a + 0.3 * rng.normal(size=500))))))   # the reference's expression
    rng = np.random.default_rng(0); a = rng.uniform(0, 2 * np.pi, 500); b = a + 0.3 * np.random.default_rng(5).normal(size=500)
    assert abs(circular_corrcoef(a, b) - np.abs(np.mean(np.exp(1j * a) * np.conj(np.exp(1j * b))))) < 1e-12
    with pytest.raises(AssertionError):
        circular_corrcoef(a, a[:10])
