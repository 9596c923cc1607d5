"""ctypes binding of ``libvcb.so`` (C ABI declared in ``include/vcb.h``).

The library is built in-tree by ``__graft_entry__.build()`` (plain ``nvcc -shared``); there is no
fallback: if it is missing, importing anything that needs it raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvcb.so")

VCB_FLAG_GRAD = 1
VCB_FLAG_LGAMMA_INLINE = 2
VCB_FLAG_TCGEN05 = 4
VCB_FLAG_LEGACY_STREAM = 8
VCB_MAX_HARMONICS = 5
VCB_COUNTS_U8, VCB_COUNTS_U16, VCB_COUNTS_I32 = 1, 2, 4
VCB_CSR_F32, VCB_CSR_I32, VCB_CSR_F64, VCB_CSR_I64 = 0, 1, 2, 3
VCB_COUNTS_B2, VCB_COUNTS_B4 = 16, 32  # HostCounts-only format tags: 2 / 4 bits per entry (vcb_expand_counts_packed)
VCB_COUNTS_B2N = 64  # 2-bit codes -> nibble stream -> byte stream (vcb_expand_counts_twolevel)
VCB_PACKED_BLOCK_WORDS = 256

c_float_p = C.c_void_p  # device pointers travel as integers


class VcbSpectrum(C.Structure):
    _fields_ = [("off", C.c_void_p), ("val", C.c_void_p), ("mult", C.c_void_p), ("lgk1", C.c_void_p)]


class VcbProblem(C.Structure):
    _fields_ = [
        ("Nc", C.c_int64), ("Ng", C.c_int64), ("ld", C.c_int64),
        ("H", C.c_int32), ("Hw", C.c_int32), ("Nb", C.c_int32), ("Nx", C.c_int32),
        ("flags", C.c_uint32), ("reserved", C.c_uint32),
        ("S", C.c_void_p), ("U", C.c_void_p),
        ("phi", C.c_void_p), ("cf", C.c_void_p), ("batch_id", C.c_void_p), ("cond_id", C.c_void_p),
        ("nu", C.c_void_p), ("dnu", C.c_void_p), ("shape_inv", C.c_void_p),
        ("logbeta", C.c_void_p), ("gamma", C.c_void_p), ("nu_omega", C.c_void_p),
        ("spec_S", VcbSpectrum), ("spec_U", VcbSpectrum),
        ("lp_S", C.c_void_p), ("lp_U", C.c_void_p),
        ("d_nu", C.c_void_p), ("d_dnu", C.c_void_p), ("d_shape_inv", C.c_void_p),
        ("d_logbeta", C.c_void_p), ("d_gamma", C.c_void_p), ("d_nu_omega", C.c_void_p),
        ("d_phi", C.c_void_p), ("d_cf", C.c_void_p), ("d_omega", C.c_void_p),
        ("ev_stream_begin", C.c_void_p), ("ev_stream_end", C.c_void_p),
    ]


class VcbSvi(C.Structure):
    """``vcb_svi_t`` of include/vcb.h (field for field)."""

    _fields_ = (
        [("Nc", C.c_int64), ("Ng", C.c_int64)]
        + [(n, C.c_int32) for n in ("H", "Hw", "Nb", "Nx", "rank", "model")]
        + [("param", C.c_void_p), ("grad", C.c_void_p)]
        + [(n, C.c_int64) for n in ("o_nu_locs", "o_nu_scales", "o_dnu_locs", "o_phixy_locs", "o_shape_inv_locs",
                                    "o_logbeta_locs", "o_logbeta_scales", "o_loggamma_locs", "o_loggamma_scales",
                                    "o_nuw_locs", "o_nuw_scales", "o_loc", "o_cov_factor", "o_cov_diag", "o_rho_real_loc")]
        + [(n, C.c_void_p) for n in ("eps_nu", "eps_loggamma", "eps_logbeta", "eps_nuw", "eps_phixy", "eps_W", "eps_D",
                                     "mu_nu", "sd_nu", "mu_loggamma", "sd_loggamma", "mu_logbeta", "sd_logbeta",
                                     "mu_nuw", "sd_nuw", "phixy_prior", "cell_row",
                                     "cond_nu", "cond_dnu", "cond_shape_inv", "cond_phixy")]
        + [(n, C.c_float) for n in ("sd_dnu", "gamma_alpha", "gamma_beta", "rho_mean", "rho_std", "rho_scale")]
        + [(n, C.c_void_p) for n in ("nu", "dnu", "shape_inv", "loggamma", "gamma", "logbeta", "nu_omega", "phixy", "phi",
                                     "lp_S", "lp_U", "d_nu", "d_dnu", "d_shape_inv", "d_logbeta", "d_gamma", "d_nu_omega",
                                     "d_phi", "cell_partials", "gene_partials", "lik_partials", "cell_lp", "loss")]
    )


VCB_MAX_RANKS = 16


class VcbComm(C.Structure):
    """``vcb_comm_t`` of include/vcb.h."""

    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("slot_floats", C.c_int64),
                ("slots", C.c_void_p * VCB_MAX_RANKS), ("flags", C.c_void_p * VCB_MAX_RANKS), ("epoch", C.c_void_p)]


EXPORTS = (
    "vcb_version",
    "vcb_strerror",
    "vcb_workspace_bytes",
    "vcb_phase_fwd_bwd",
    "vcb_velocity_fwd_bwd",
    "vcb_count_histogram",
    "vcb_expand_counts",
    "vcb_csr_to_counts",
    "vcb_expand_counts_packed",
    "vcb_expand_counts_twolevel",
    "vcb_clipped_adam",
    "vcb_svi_partials",
    "vcb_svi_sample",
    "vcb_svi_backward",
    "vcb_allreduce_sum",
)

_lib: Optional[C.CDLL] = None


class VcbError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libvcb.so once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VcbError(
            f"{LIB_PATH} not found: the CUDA library has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` at the repo root (needs nvcc). "
            "velocycle_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    lib.vcb_version.restype = C.c_int
    lib.vcb_strerror.restype = C.c_char_p
    lib.vcb_strerror.argtypes = [C.c_int]
    lib.vcb_workspace_bytes.restype = C.c_size_t
    lib.vcb_workspace_bytes.argtypes = [C.POINTER(VcbProblem)]
    for name in ("vcb_phase_fwd_bwd", "vcb_velocity_fwd_bwd"):
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(VcbProblem), C.c_void_p, C.c_size_t, C.c_void_p]
    lib.vcb_count_histogram.restype = C.c_int
    lib.vcb_count_histogram.argtypes = [
        C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
    ]
    lib.vcb_expand_counts_packed.restype = C.c_int
    lib.vcb_expand_counts_packed.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_int64, C.c_void_p]
    lib.vcb_expand_counts_twolevel.restype = C.c_int
    lib.vcb_expand_counts_twolevel.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.vcb_csr_to_counts.restype = C.c_int
    lib.vcb_csr_to_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vcb_expand_counts.restype = C.c_int
    lib.vcb_expand_counts.argtypes = [
        C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
    ]
    lib.vcb_clipped_adam.restype = C.c_int
    lib.vcb_clipped_adam.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
        C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p,
    ]
    lib.vcb_svi_partials.restype = C.c_int
    lib.vcb_svi_partials.argtypes = [C.c_int64, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    for name in ("vcb_svi_sample", "vcb_svi_backward"):
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(VcbSvi), C.c_void_p]
    lib.vcb_allreduce_sum.restype = C.c_int
    lib.vcb_allreduce_sum.argtypes = [C.POINTER(VcbComm), C.c_void_p, C.c_int64, C.c_void_p]
    _lib = lib
    return lib


def check(code: int, what: str = "libvcb") -> None:
    if code != 0:
        msg = load().vcb_strerror(code).decode()
        raise VcbError(f"{what} failed with code {code}: {msg}")
