"""Post-fit posterior summaries (velocycle_b200/posterior.py) against the reference's own einsum expressions
(velocity_inference_model.py:236-258, phase_inference_model.py:245-253) on tensors laid out as the reference lays them out."""
import torch

from velocycle_b200.posterior import expected_log_counts_summary
from velocycle_b200.utils import torch_fourier_basis


def test_expected_log_counts_match_the_reference_expressions():
    g = torch.Generator().manual_seed(0)
    Ng, Nc, H, Hw, Nb, Nx = 7, 23, 2, 1, 3, 2
    K, Kw = 2 * H + 1, 2 * Hw + 1
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    nu_ref = rn(Ng, 1, K)                    # pyro.param("ν_locs")
    dnu_ref = 0.1 * rn(Nb, 1, 1, Ng, 1)      # pyro.param("Δν_locs") of the velocity guide
    phis = torch.rand(Nc, generator=g, dtype=torch.float64) * 6.283
    bid = torch.randint(0, Nb, (Nc,), generator=g)
    cid = torch.randint(0, Nx, (Nc,), generator=g)
    Db = torch.nn.functional.one_hot(bid, Nb).T[:, None, None, None, :].double()      # (Nb,1,1,1,Nc)
    D = torch.nn.functional.one_hot(cid, Nx).T[:, None, None, :].double()             # (Nx,1,1,Nc)
    cf = 0.2 * rn(1, 1, 1, Nc)
    nuw = rn(Nx, Kw, 1, 1)                   # posterior mean of "νω"
    gamma, logbeta = rn(Ng).exp(), rn(Ng)
    # ---- the reference's expressions, verbatim einsum strings ----
    ζ = torch_fourier_basis(phis, num_harmonics=H, der=0)
    ζ_dϕ = torch_fourier_basis(phis, num_harmonics=H, der=1)
    ζω = torch_fourier_basis(phis, num_harmonics=Hw, der=0).T
    ref = {}
    for tag, c in (("", cf), ("2", torch.full_like(cf, float(cf.mean())))):
        ElogS = torch.einsum("...gch,ch->gc", nu_ref, ζ) + torch.einsum("bxhgc,bxhgc->gc", Db, dnu_ref) + c
        ω = torch.einsum("...xhgc,hc,xhgc->gc", [nuw, ζω, D])
        ElogU = -logbeta.unsqueeze(-1) + torch.log(torch.relu(torch.einsum("gch,ch->gc", nu_ref, ζ_dϕ) * ω + gamma.unsqueeze(-1)) + 1e-5) + ElogS
        ref["ElogS" + tag], ref["ElogU" + tag] = ElogS.squeeze(), ElogU.squeeze()
    got = expected_log_counts_summary(nu_ref.reshape(Ng, K), phis, cf.reshape(-1), dnu_ref.reshape(Nb, Ng), bid,
                                      velocity=dict(nu_omega=nuw.reshape(Nx, Kw), cond_id=cid, gamma=gamma, logbeta=logbeta))
    assert set(got) == {"ElogS", "ElogS2", "ElogU", "ElogU2"}
    for k in ref:
        assert got[k].shape == (Ng, Nc)
        assert torch.allclose(got[k], ref[k], rtol=0, atol=1e-12), k
    # phase model: no velocity part, no batches
    got = expected_log_counts_summary(nu_ref.reshape(Ng, K), phis, cf.reshape(-1))
    assert set(got) == {"ElogS", "ElogS2"}
    assert torch.allclose(got["ElogS"], (torch.einsum("...gch,ch->gc", nu_ref, ζ) + cf).squeeze(), rtol=0, atol=1e-12)
