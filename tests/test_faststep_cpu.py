"""CPU: which (model, guide) pairs the fused SVI step recognises, and that nothing falls back to a CPU evaluation."""
import collections

import pytest
import torch


def _mp(**kw):
    MP = collections.namedtuple("MP", "noisemodel with_delta_nu device")
    return MP(kw.get("noisemodel", "NegativeBinomial"), kw.get("with_delta_nu", True), kw.get("device", "cpu"))


def test_model_code_recognises_the_standard_pairs_and_the_tutorial_conditioning():
    from velocycle_b200 import phase_inference_guide as pg, phase_inference_model as pm
    from velocycle_b200 import velocity_inference_guide as vg, velocity_inference_model as vm
    from velocycle_b200.faststep import model_code
    from velocycle_b200.ppl import poutine

    mp = _mp()
    assert model_code(pm.phase_latent_variable_model, pg.phase_latent_variable_guide, mp) == (0, {})
    assert model_code(vm.velocity_latent_variable_model, vg.velocity_latent_variable_guide, mp) == (1, {})
    assert model_code(vm.velocity_latent_variable_model_LRMN, vg.velocity_latent_variable_guide_LRMN, mp) == (2, {})
    # mismatched pair, other noise model, user function
    assert model_code(vm.velocity_latent_variable_model, vg.velocity_latent_variable_guide_LRMN, mp) is None
    assert model_code(pm.phase_latent_variable_model, pg.phase_latent_variable_guide, _mp(noisemodel="Poisson")) is None
    assert model_code(lambda mp: None, pg.phase_latent_variable_guide, mp) is None
    # the fit drivers' wrapping for condition_on (phase_inference_model.py:110-115)
    cond = {"ϕxy": torch.zeros(3, 2), "ν": torch.zeros(4, 1, 3), "shape_inv": torch.ones(4, 1), "Δν": torch.zeros(1, 1, 1, 4, 1)}
    m = poutine.condition(vm.velocity_latent_variable_model_LRMN, data=cond)
    g = poutine.block(vm.velocity_latent_variable_guide_LRMN, hide=list(cond))
    code, c = model_code(m, g, mp)
    assert code == 2 and set(c) == set(cond)
    # a guide that hides something else than the model conditions on, or an unsupported site: the traced step
    assert model_code(m, poutine.block(vm.velocity_latent_variable_guide_LRMN, hide=["ν"]), mp) is None
    m2 = poutine.condition(vm.velocity_latent_variable_model, data={"logβg": torch.zeros(4, 1)})
    assert model_code(m2, poutine.block(vm.velocity_latent_variable_guide, hide=["logβg"]), mp) is None
    assert model_code(poutine.condition(pm.phase_latent_variable_model, data={"Δν": torch.zeros(1, 4, 1)}),
                      poutine.block(pg.phase_latent_variable_guide, hide=["Δν"]), _mp(with_delta_nu=False)) is None


def test_graphed_svi_refuses_a_cpu_device():
    from velocycle_b200 import _lib
    from velocycle_b200.svi import GraphedSVI

    with pytest.raises(_lib.VcbError):
        GraphedSVI(lambda mp: None, lambda mp: None, {"lr": 0.01}, _mp())
