"""GPU parity: the fused CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerance (BASELINE.json north_star): ELBO and all gradients within 1e-4 relative in fp32.  Gradients are
compared normwise per tensor (max |diff| <= tol * max |ref|) against the float64 oracle; inputs are
generated away from the optimum so that the gradient norms are not cancellation noise.
"""
import math

import pytest
import torch

from oracle.likelihood import analytic_gradients, fused_reference, fused_reference_chunked

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _problem(d, velocity, with_dnu=True):
    gamma = torch.exp(d.loggamma)
    p = dict(S=d.S[:, : d.Ng], phi=d.phi, cf=d.cf, batch_id=d.batch_id, nu=d.nu, shape_inv=d.shape_inv)
    if with_dnu:
        p["dnu"] = d.dnu
    if velocity:
        p.update(U=d.U[:, : d.Ng], cond_id=d.cond_id, logbeta=d.logbeta, gamma=gamma, nu_omega=d.nu_omega)
    return p


def _run(d, velocity, with_dnu=True, inline=False, grad=True, perturb=0.3, seed=0, relu_margin=0.05, tcgen05=False):
    """``relu_margin``: after perturbing, raise gamma_g so that a = d*omega + gamma >= margin everywhere.
    At the kink of relu(a)+1e-5 an observed kU > 0 makes dL/da = kU/m ~ 1e5*kU, so ANY two fp32 evaluations
    (the reference's own included) differ there by ~1e-7*|d omega|/1e-5 = 1% -- parity to 1e-4 is only defined
    away from it.  The relu-dead branch (a < 0) is covered by test_relu_dead_elements_and_zero_counts."""
    from velocycle_b200.fused import PackedCounts, fused_elbo_grad
    from velocycle_b200.synthetic import fourier_rows

    g = torch.Generator(device="cpu").manual_seed(seed)
    # evaluate away from the generating parameters
    dev = d.S.device
    def jit(t, s=perturb):
        return t + s * torch.randn(t.shape, generator=g).to(dev)
    d.nu = jit(d.nu)
    d.phi = jit(d.phi)
    d.logbeta = jit(d.logbeta)
    d.loggamma = jit(d.loggamma)
    d.shape_inv = d.shape_inv * torch.exp(jit(torch.zeros_like(d.shape_inv)))
    d.nu_omega = jit(d.nu_omega, 0.1)
    if velocity and relu_margin is not None:
        H, Hw = (d.nu.shape[1] - 1) // 2, (d.nu_omega.shape[1] - 1) // 2
        dd = fourier_rows(d.phi, H, 1) @ d.nu.T
        om = (fourier_rows(d.phi, Hw, 0) * d.nu_omega[d.cond_id.long()]).sum(-1)
        need = (-(dd * om[:, None])).amax(0).clamp_min(0.0) + relu_margin
        d.loggamma = torch.maximum(d.loggamma, torch.log(need))
    counts = PackedCounts(d.S, d.U if velocity else None, d.Ng, d.batch_id, d.cond_id, spectrum=not inline)
    p = _problem(d, velocity, with_dnu)
    out = fused_elbo_grad(
        counts, p["phi"], p["cf"], p["nu"], p.get("dnu"), p["shape_inv"],
        p.get("logbeta"), p.get("gamma"), p.get("nu_omega"), grad=grad, inline_lgamma=inline, want_d_omega=True,
        tcgen05=tcgen05,
    )
    torch.cuda.synchronize()
    ref = fused_reference(p, dtype=torch.float64, grad=grad)
    ref.update({"_ref32": fused_reference(p, dtype=torch.float32, grad=grad)})
    return out, ref, p


def _compare(out, ref, tol=TOL, ref32=None):
    ref = dict(ref)
    ref32 = ref.pop("_ref32", ref32)
    """Normwise error against the fp64 oracle: <= 1e-4 for every tensor.  One exception: d/dshape_inv =
    -r^2 (psi - L - dnu0 / r) is a difference of sums ~1e3 times larger than itself, so at a few hundred cells fp32
    evaluation of the reference op chain itself is further than 1e-4 from fp64; there the bound is twice the
    reference's own fp32 error."""
    checked = 0
    for k, v in ref.items():
        if k in ("total", "omega") or k not in out:
            continue
        got = out[k].double().cpu().reshape(v.shape)
        assert torch.isfinite(got).all(), k
        err = float((got - v).abs().max() / (v.abs().max() + 1e-30))
        bound = tol
        if k == "d_shape_inv" and ref32 is not None and k in ref32:  # the one documented exception (cancellation)
            e32 = float((ref32[k].double().reshape(v.shape) - v).abs().max() / (v.abs().max() + 1e-30))
            bound = max(tol, 2.0 * e32)
        assert err <= bound, f"{k}: normwise rel err {err:.3e} > {bound:.3e}"
        checked += 1
    tot = out["lp_S"].double().sum().item() + (out["lp_U"].double().sum().item() if "lp_U" in out else 0.0)
    assert abs(tot - float(ref["total"])) <= tol * abs(float(ref["total"]))
    return checked


SHAPES = [
    # Nc, Ng, H, Hw, Nb, Nx
    (11, 7, 2, 1, 3, 2),        # the SURVEY Appendix-A probe shape
    (257, 203, 3, 1, 2, 2),     # ragged: Ng % 4 != 0, Nc % stage != 0
    (1849, 76, 1, 0, 1, 1),     # the Appendix-B (Stereo-seq notebook) shape
    (600, 1918, 1, 1, 1, 1),    # "Large" gene set width, one 512-thread tile with padding lanes
    (300, 2500, 3, 1, 4, 2),    # several gene tiles -> cross-tile cell partials
    (64, 128, 0, 0, 1, 1),      # H = 0: constant-only basis
    (130, 40, 4, 2, 2, 3),
    (97, 33, 5, 5, 1, 2),       # maximum harmonics compiled in
    (203, 50, 2, 5, 2, 2),      # round-2 path (H <= 3) with more angular-speed harmonics than the table recurrence holds
    (41, 36, 3, 3, 1, 3),
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("velocity", [False, True])
def test_fused_matches_oracle(shape, velocity):
    from velocycle_b200.synthetic import make_synthetic

    Nc, Ng, H, Hw, Nb, Nx = shape
    d = make_synthetic(Nc, Ng, H=H, Hw=Hw, Nb=Nb, Nx=Nx, seed=3, device="cuda", sorted_batches=True)
    out, ref, _ = _run(d, velocity)
    n = _compare(out, ref)
    assert n >= (9 if velocity else 6)


# Shapes whose CTAs walk many more 8-cell groups than either shared-memory ring holds (count ring: 6 groups; table
# ring: 6-16 groups, FEWER slots than warps for the velocity model -> the two-phases-ahead case of the mbarrier
# parity waits), with batch boundaries inside CTAs and a ragged last group.
RING_SHAPES = [
    # Nc, Ng, H, Hw, Nb, Nx, sorted
    (20003, 2000, 3, 1, 1, 1, True),    # BASELINE gene count: 4 gene tiles x 37 cell splits, ~68 groups per CTA
    (30001, 512, 1, 1, 3, 2, True),     # one gene tile, 148 cell splits, batch boundaries inside CTAs
    (9000, 40, 2, 1, 4, 2, False),      # 2-warp CTAs (8 per SM), every group mixes batches
]


@pytest.mark.parametrize("shape", RING_SHAPES)
@pytest.mark.parametrize("velocity", [False, True])
def test_long_streams_match_oracle(shape, velocity):
    from velocycle_b200.synthetic import make_synthetic

    Nc, Ng, H, Hw, Nb, Nx, sorted_b = shape
    d = make_synthetic(Nc, Ng, H=H, Hw=Hw, Nb=Nb, Nx=Nx, seed=13, device="cuda", sorted_batches=sorted_b)
    out, ref, _ = _run(d, velocity)
    _compare(out, ref)


def test_results_are_deterministic(monkeypatch):
    """No atomics on the hot path (sorted batches): two evaluations agree bit for bit, also when the scratch
    workspace starts out as NaNs (nothing reads workspace bytes it has not written)."""
    from velocycle_b200.fused import PackedCounts, fused_elbo_grad
    from velocycle_b200.synthetic import make_synthetic

    d = make_synthetic(20000, 700, H=3, Hw=1, Nb=1, Nx=2, seed=5, device="cuda")
    counts = PackedCounts(d.S, d.U, d.Ng, d.batch_id, d.cond_id)
    args = (counts, d.phi, d.cf, d.nu, None, d.shape_inv, d.logbeta, torch.exp(d.loggamma), d.nu_omega)
    a = {k: v.clone() for k, v in fused_elbo_grad(*args, grad=True).items()}
    monkeypatch.setenv("VCB_DEBUG_POISON_WS", "1")
    b = fused_elbo_grad(*args, grad=True)
    for k in a:
        if not k.startswith("_"):  # ("_workspace" is scratch kept alive with the result)
            assert torch.equal(a[k], b[k]), k


def _perturbed(Nc, Ng, Nb, Nx, seed, sorted_batches=True):
    """Synthetic data evaluated away from the generating parameters, gamma lifted off the relu kink (see _run)."""
    from velocycle_b200.synthetic import fourier_rows, make_synthetic

    d = make_synthetic(Nc, Ng, H=3, Hw=1, Nb=Nb, Nx=Nx, seed=seed, device="cuda", stats=False, sorted_batches=sorted_batches)
    g = torch.Generator(device="cpu").manual_seed(seed)
    jit = lambda t, s=0.3: t + s * torch.randn(t.shape, generator=g).to(t.device)
    d.nu, d.phi, d.logbeta, d.loggamma = jit(d.nu), jit(d.phi), jit(d.logbeta), jit(d.loggamma)
    d.shape_inv = d.shape_inv * torch.exp(jit(torch.zeros_like(d.shape_inv)))
    d.nu_omega = jit(d.nu_omega, 0.1)
    dd = fourier_rows(d.phi, 3, 1) @ d.nu.T
    om = (fourier_rows(d.phi, 1, 0) * d.nu_omega[d.cond_id.long()]).sum(-1)
    d.loggamma = torch.maximum(d.loggamma, torch.log((-(dd * om[:, None])).amax(0).clamp_min(0.0) + 0.05))
    return d


def _against_chunked_oracle(d, with_dnu):
    from velocycle_b200.fused import PackedCounts, fused_elbo_grad

    counts = PackedCounts(d.S, d.U, d.Ng, d.batch_id, d.cond_id)
    p = _problem(d, True, with_dnu)
    out = fused_elbo_grad(counts, p["phi"], p["cf"], p["nu"], p.get("dnu"), p["shape_inv"], p["logbeta"], p["gamma"],
                          p["nu_omega"], grad=True, want_d_omega=True)
    torch.cuda.synchronize()
    ref = fused_reference_chunked(p, dtype=torch.float64)
    for k, v in ref.items():
        if k in ("total", "omega") or k not in out:
            continue
        got = out[k].double().cpu().reshape(v.shape)
        err = float((got - v).abs().max() / (v.abs().max() + 1e-30))
        # d/dshape_inv = -r^2 (psi - L - dnu0/r) is a difference of sums ~1e3 times larger than itself: fp32 partial sums
        # over 1e5 cells leave it ~1e-3 accurate in ANY fp32 evaluation (the small-shape tests bound it by the reference's
        # own fp32 error); every other tensor: the 1e-4 contract
        assert err <= (2e-3 if k == "d_shape_inv" else TOL), f"{k}: normwise rel err {err:.3e}"
    tot = out["lp_S"].double().sum().item() + out["lp_U"].double().sum().item()
    assert abs(tot - float(ref["total"])) <= TOL * abs(float(ref["total"]))
    return counts


def test_config_c3_100k_x_2k_matches_oracle():
    """BASELINE config C3 (100 000 cells x 2 000 genes, H = 3, velocity): the kernel against the fp64 op chain evaluated in
    cell blocks -- 4 gene tiles x 37 cell splits, ~170 ring stages per CTA."""
    _against_chunked_oracle(_perturbed(100_000, 2000, 1, 1, seed=17), with_dnu=False)


def test_config_c5_tiling_16_batches_matches_oracle():
    """BASELINE config C5's tiling (5 000 genes = 10 gene tiles, 16 batches with per-batch gene offsets, 2 conditions) on
    20 000 cells handed over in SHUFFLED batch order: PackedCounts sorts the rows by batch once, the kernel switches offsets
    at stage boundaries inside CTAs, results come back in the caller's cell order -- and twice the same bits."""
    from velocycle_b200.fused import fused_elbo_grad

    d = _perturbed(20_000, 5000, 16, 2, seed=23, sorted_batches=False)
    counts = _against_chunked_oracle(d, with_dnu=True)
    assert counts.perm is not None
    p = _problem(d, True, True)
    args = (counts, p["phi"], p["cf"], p["nu"], p["dnu"], p["shape_inv"], p["logbeta"], p["gamma"], p["nu_omega"])
    a = {k: v.clone() for k, v in fused_elbo_grad(*args, grad=True).items() if not k.startswith("_")}
    b = fused_elbo_grad(*args, grad=True)
    for k in a:
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("velocity", [False, True])
def test_inline_lgamma_matches_oracle(velocity):
    from velocycle_b200.synthetic import make_synthetic

    d = make_synthetic(300, 150, H=2, Hw=1, Nb=2, Nx=2, seed=5, device="cuda")
    out, ref, _ = _run(d, velocity, inline=True)
    _compare(out, ref)


@pytest.mark.parametrize("velocity", [False, True])
def test_unsorted_batches_and_no_dnu(velocity):
    from velocycle_b200.synthetic import make_synthetic

    d = make_synthetic(500, 90, H=2, Hw=1, Nb=5, Nx=3, seed=11, device="cuda", sorted_batches=False)
    out, ref, _ = _run(d, velocity)
    _compare(out, ref)
    d = make_synthetic(500, 90, H=2, Hw=1, Nb=1, Nx=1, seed=12, device="cuda")
    out, ref, _ = _run(d, velocity, with_dnu=False)
    assert "d_dnu" not in out
    _compare(out, ref)


def test_forward_only_matches_oracle():
    from velocycle_b200.synthetic import make_synthetic

    d = make_synthetic(333, 210, H=3, Hw=1, Nb=2, Nx=2, seed=2, device="cuda")
    out, ref, _ = _run(d, True, grad=False)
    assert "d_nu" not in out
    _compare(out, ref)


def test_relu_dead_elements_and_zero_counts():
    """a = d*omega + gamma <= 0 for many elements (subgradient 0 there), and an all-zero gene / cell."""
    from velocycle_b200.synthetic import make_synthetic

    d = make_synthetic(400, 64, H=2, Hw=1, Nb=1, Nx=1, seed=9, device="cuda")
    d.nu[:, 1:] *= 6.0           # large derivative amplitudes
    d.loggamma -= 3.0            # small gamma: relu goes dead on half the circle
    d.nu_omega[:, 0] = 1.0
    d.S[:, 5] = 0
    d.U[:, 5] = 0
    d.S[17, :] = 0
    d.U[17, :] = 0
    out, ref, p = _run(d, True, perturb=0.0, relu_margin=None)
    a = analytic_gradients(p)
    assert float((a["d_gamma"] - ref["d_gamma"]).abs().max()) < 1e-9
    _compare(out, ref)


def test_large_counts_and_dispersion_range():
    from velocycle_b200.synthetic import make_synthetic

    d = make_synthetic(256, 48, H=1, Hw=0, Nb=1, Nx=1, seed=4, device="cuda")
    d.S[:, :8] = torch.round(d.S[:, :8] * 40 + 100)   # counts in the hundreds
    d.nu[:8, 0] += 5.0
    d.shape_inv[:16] = torch.logspace(-3, 1, 16, device="cuda")  # r from 1000 to 0.1
    out, ref, _ = _run(d, True, perturb=0.05)
    _compare(out, ref)


def test_matches_fp32_reference_chain():
    """Against the reference op chain evaluated in fp32 (what the reference itself would return)."""
    from velocycle_b200.synthetic import make_synthetic

    d = make_synthetic(512, 256, H=3, Hw=1, Nb=2, Nx=2, seed=21, device="cuda")
    out, ref64, p = _run(d, True)
    ref32 = fused_reference(p, dtype=torch.float32, grad=True)
    for k in ("d_nu", "d_phi", "d_logbeta", "d_gamma", "d_shape_inv", "d_nu_omega", "lp_S", "lp_U"):
        got = out[k].double().cpu().reshape(ref64[k].shape)
        e_ours = float((got - ref64[k]).abs().max() / ref64[k].abs().max())
        e_ref32 = float((ref32[k].double() - ref64[k]).abs().max() / ref64[k].abs().max())
        # the fused path must be at least as close to the fp64 truth as 1e-4, and the fp32 chain within the same band
        assert e_ours <= TOL, (k, e_ours)
        assert e_ref32 <= 10 * TOL, (k, e_ref32)


def test_at_the_relu_kink_no_worse_than_the_fp32_reference_chain():
    """Elements with a = d*omega + gamma in (0, 0.05): dL/da = kU / m reaches 1e5 kU at m = relu(a) + 1e-5, so fp32
    rounding of `a` alone moves the gamma / nu / phi / nu_omega gradients by up to per cent in ANY fp32 evaluation.  Here
    gamma is set so that thousands of elements sit in that band (none lifted away), and the kernel must be as close to
    the fp64 truth as the reference's own fp32 op chain is (factor 3: both errors are a few sign-flip sized rounding
    events, i.e. noisy), tensor by tensor; tensors the kink does not touch keep the 1e-4 contract."""
    from velocycle_b200.synthetic import fourier_rows, make_synthetic

    d = make_synthetic(512, 128, H=2, Hw=1, Nb=1, Nx=1, seed=31, device="cuda")
    dd = fourier_rows(d.phi, 2, 1) @ d.nu.T
    om = (fourier_rows(d.phi, 1, 0) * d.nu_omega[d.cond_id.long()]).sum(-1)
    a_wo_gamma = dd * om[:, None]
    # per gene: gamma = a quantile of -d*omega, so that a = d*omega + gamma straddles zero for that gene
    d.loggamma = torch.log((-a_wo_gamma).quantile(0.7, dim=0).clamp_min(1e-3))
    a = a_wo_gamma + torch.exp(d.loggamma)[None, :]
    n_band = int(((a > 0) & (a < 0.05)).sum())
    assert n_band > 500, n_band
    out, ref64, p = _run(d, True, perturb=0.0, relu_margin=None)
    ref32 = ref64.pop("_ref32")
    for k in ("d_nu", "d_phi", "d_gamma", "d_nu_omega", "d_omega", "d_logbeta", "d_shape_inv", "lp_S", "lp_U"):
        v = ref64[k]
        got = out[k].double().cpu().reshape(v.shape)
        e_ours = float((got - v).abs().max() / v.abs().max())
        e_ref32 = float((ref32[k].double().reshape(v.shape) - v).abs().max() / v.abs().max())
        assert e_ours <= max(TOL, 3.0 * e_ref32), (k, e_ours, e_ref32)


def test_autograd_function_scales_and_poisons():
    from velocycle_b200.fused import PackedCounts, fused_cycle_nb
    from velocycle_b200.synthetic import make_synthetic

    d = make_synthetic(200, 60, H=2, Hw=1, Nb=2, Nx=2, seed=8, device="cuda")
    d.loggamma += 1.5  # keep a = d*omega + gamma away from the relu kink (see _run)
    counts = PackedCounts(d.S, d.U, d.Ng, d.batch_id, d.cond_id)
    leaves = [t.clone().requires_grad_(True) for t in (d.phi, d.nu, d.dnu, d.shape_inv, d.logbeta, d.loggamma, d.nu_omega)]
    phi, nu, dnu, sinv, lb, lg, nw = leaves
    lpS, lpU = fused_cycle_nb(counts, phi, d.cf, nu, dnu, sinv, lb, torch.exp(lg), nw)
    loss = -(lpS.sum() + lpU.sum())
    loss.backward()
    p = dict(S=d.S[:, : d.Ng], U=d.U[:, : d.Ng], phi=d.phi, cf=d.cf, batch_id=d.batch_id, cond_id=d.cond_id, nu=d.nu,
             dnu=d.dnu, shape_inv=d.shape_inv, logbeta=d.logbeta, gamma=torch.exp(d.loggamma), nu_omega=d.nu_omega)
    ref = fused_reference(p)
    assert abs(loss.item() + float(ref["total"])) <= TOL * abs(float(ref["total"]))
    for leaf, key in zip(leaves, ["d_phi", "d_nu", "d_dnu", "d_shape_inv", "d_logbeta", None, "d_nu_omega"]):
        if key is None:
            want = -(ref["d_gamma"] * torch.exp(d.loggamma.double().cpu()))  # chain rule through exp stays in torch
        else:
            want = -ref[key]
        got = leaf.grad.double().cpu().reshape(want.shape)
        assert float((got - want).abs().max() / want.abs().max()) <= TOL
    # unequal upstream weights are refused loudly (NaN), never silently wrong
    leaves2 = [t.clone().requires_grad_(True) for t in (d.phi, d.nu)]
    lpS, lpU = fused_cycle_nb(counts, leaves2[0], d.cf, leaves2[1], d.dnu, d.shape_inv, d.logbeta, torch.exp(d.loggamma), d.nu_omega)
    (lpS.sum() + 2.0 * lpU.sum()).backward()
    assert torch.isnan(leaves2[1].grad).all()


def test_histogram_rejects_non_integer_counts():
    from velocycle_b200 import _lib
    from velocycle_b200.fused import PackedCounts
    from velocycle_b200.synthetic import make_synthetic

    d = make_synthetic(64, 32, H=1, Hw=0, seed=1, device="cuda")
    d.S[3, 4] = 0.5
    with pytest.raises(_lib.VcbError):
        PackedCounts(d.S, None, d.Ng)


def test_abi_argument_errors():
    import ctypes as C
    from velocycle_b200 import _lib

    lib = _lib.load()
    p = _lib.VcbProblem()
    assert lib.vcb_phase_fwd_bwd(C.byref(p), None, 0, None) < 0
    p.Nc, p.Ng, p.ld, p.H = 4, 6, 6, 1   # ld not a multiple of 4
    t = torch.zeros(64, device="cuda")
    p.S = p.phi = p.nu = p.shape_inv = p.lp_S = t.data_ptr()
    p.flags = _lib.VCB_FLAG_LGAMMA_INLINE
    assert lib.vcb_phase_fwd_bwd(C.byref(p), t.data_ptr(), 256, None) == -3
    p.ld = 8
    p.H = 9
    assert lib.vcb_phase_fwd_bwd(C.byref(p), t.data_ptr(), 256, None) == -4
    p.H = 1
    assert lib.vcb_phase_fwd_bwd(C.byref(p), t.data_ptr(), 16, None) == -5
    assert b"workspace" in lib.vcb_strerror(-5)
