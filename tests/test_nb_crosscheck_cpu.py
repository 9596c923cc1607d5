"""The restated ``GammaPoisson.log_prob`` against an independent implementation.

Pyro is not installable here (third-party, ``pyro-ppl==1.8.6`` in the reference's ``requirements.txt:105``), so its
``GammaPoisson.log_prob`` (``pyro/distributions/conjugate.py``) is restated twice in this repo: ``oracle.likelihood.
gamma_poisson_log_prob`` (the checker) and ``velocycle_b200.ppl.distributions.GammaPoisson`` (used when the reference's
model source is executed to produce the golden fixtures).  Both are pinned here against
``torch.distributions.NegativeBinomial`` -- the same pmf in another parameterisation, written by other people -- over the
whole range the path sees: r = 1/shape_inv in [0.1, 1000], counts 0..500, means from 1e-3 to 1e3.
GammaPoisson(concentration=r, rate=r/mu)  ==  NegativeBinomial(total_count=r, probs=mu/(r+mu)).
"""
import pytest
import torch

from oracle.likelihood import gamma_poisson_log_prob
from velocycle_b200.ppl.distributions import GammaPoisson


def _grid():
    r = torch.logspace(-1, 3, 41, dtype=torch.float64)[:, None, None]
    mu = torch.logspace(-3, 3, 25, dtype=torch.float64)[None, :, None]
    k = torch.cat([torch.arange(0, 40), torch.arange(40, 501, 20)]).to(torch.float64)[None, None, :]
    return torch.broadcast_tensors(r, mu, k)


def test_oracle_gamma_poisson_matches_torch_negative_binomial():
    r, mu, k = _grid()
    ref = torch.distributions.NegativeBinomial(total_count=r, probs=mu / (r + mu), validate_args=False).log_prob(k)
    got = gamma_poisson_log_prob(r, r / mu, k)
    assert float((got - ref).abs().max()) < 1e-10
    assert float(((got - ref).abs() / ref.abs().clamp_min(1.0)).max()) < 1e-11


def test_ppl_gamma_poisson_matches_torch_negative_binomial():
    r, mu, k = _grid()
    ref = torch.distributions.NegativeBinomial(total_count=r, probs=mu / (r + mu), validate_args=False).log_prob(k)
    got = GammaPoisson(r, r / mu).log_prob(k)
    assert float((got - ref).abs().max()) < 1e-10


def test_gamma_poisson_gradients_match_negative_binomial():
    """d/d eta and d/d shape_inv through the reference's parameterisation (1/shape_inv, 1/(shape_inv exp(eta)))."""
    torch.manual_seed(0)
    eta = torch.randn(200, dtype=torch.float64, requires_grad=True)
    si = (torch.rand(200, dtype=torch.float64) * 3 + 0.05).requires_grad_(True)
    k = torch.poisson(torch.exp(eta.detach()) * 2)
    lp = gamma_poisson_log_prob(1.0 / si, 1.0 / (si * torch.exp(eta)), k).sum()
    g_eta, g_si = torch.autograd.grad(lp, (eta, si))
    eta2, si2 = eta.detach().clone().requires_grad_(True), si.detach().clone().requires_grad_(True)
    r2, mu2 = 1.0 / si2, torch.exp(eta2)
    lp2 = torch.distributions.NegativeBinomial(total_count=r2, probs=mu2 / (r2 + mu2), validate_args=False).log_prob(k).sum()
    h_eta, h_si = torch.autograd.grad(lp2, (eta2, si2))
    assert float((g_eta - h_eta).abs().max() / h_eta.abs().max()) < 1e-10
    assert float((g_si - h_si).abs().max() / h_si.abs().max()) < 1e-9


@pytest.mark.parametrize("dtype", [torch.float32])
def test_fp32_evaluation_is_fp32_accurate(dtype):
    """The reference evaluates in fp32: its own error against fp64 is what the 1e-4 parity tolerance has to absorb."""
    r, mu, k = _grid()
    ref = gamma_poisson_log_prob(r, r / mu, k)
    got = gamma_poisson_log_prob(r.to(dtype), (r / mu).to(dtype), k.to(dtype)).double()
    rel = ((got - ref).abs() / ref.abs().clamp_min(1.0)).max()
    assert float(rel) < 2e-3  # lgamma cancellation at r ~ 1000, k ~ 500 costs the fp32 op chain about 1e-3 relative
