"""The operand-table kernels (velocycle_b200/csrc/vcb.cu: vcb_cell_tables4_kernel / vcb_cell_tables_kernel) evaluate one
``sincosf`` per cell and obtain sin / cos of the higher harmonics by the angle-addition recurrence
``s_n = s_{n-1} c_1 + c_{n-1} s_1``, ``c_n = c_{n-1} c_1 - s_{n-1} s_1`` (two fused multiply-adds per step) instead of the reference's
direct ``torch.sin(n * phi)`` (``utils.py:400-437``).  This emulates the recurrence in float32 exactly as the kernel orders it
and bounds its distance from the float64 values: a few float32 ulp at n = 5, three orders of magnitude inside the 1e-4 contract."""
import numpy as np


def fma(a, b, c):
    """float32 fused multiply-add (one rounding), emulated in float64: exact for float32 operands up to double rounding."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def test_harmonic_recurrence_stays_within_a_few_ulp():
    rng = np.random.default_rng(0)
    phi = rng.uniform(-np.pi, np.pi, 200_000).astype(np.float32)
    s1 = np.sin(phi.astype(np.float64)).astype(np.float32)  # sincosf: correctly rounded to ~1 ulp
    c1 = np.cos(phi.astype(np.float64)).astype(np.float32)
    s, c = s1.copy(), c1.copy()
    worst = 0.0
    for n in range(2, 6):
        s_new = fma(s, c1, (c * s1).astype(np.float32))
        c_new = fma(c, c1, (-(s * s1)).astype(np.float32))
        s, c = s_new, c_new
        err = max(np.abs(s - np.sin(n * phi.astype(np.float64))).max(), np.abs(c - np.cos(n * phi.astype(np.float64))).max())
        worst = max(worst, err)
        assert err <= (n + 1) * 2.0 ** -23, (n, err)   # values are in [-1, 1]: one ulp there is 2^-24 .. 2^-23
    assert worst < 1e-6
