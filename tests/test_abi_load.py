"""CPU: the C-ABI library builds, loads and exports every symbol include/vcb.h declares."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g

    g.build()
    from velocycle_b200 import _lib

    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "vcb.h")).read()
    declared = set(re.findall(r"\b(vcb_[a-z_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.vcb_version() == 210
    assert lib.vcb_strerror(0) == b"ok"
    assert b"NULL" in lib.vcb_strerror(-1)


def test_problem_struct_matches_header_layout():
    from velocycle_b200 import _lib

    # 3 int64 + 6 int32/uint32 + 12 pointers + 2*4 spectrum pointers + 11 output pointers + 2 event handles
    assert C.sizeof(_lib.VcbProblem) == 3 * 8 + 6 * 4 + 12 * 8 + 8 * 8 + 11 * 8 + 2 * 8
    assert C.sizeof(_lib.VcbSpectrum) == 32
    # vcb_svi_t: 2 int64 + 6 int32 + 2 pointers + 15 int64 offsets + 21 pointers + 6 floats + 23 pointers
    assert C.sizeof(_lib.VcbSvi) == 2 * 8 + 6 * 4 + 2 * 8 + 15 * 8 + 21 * 8 + 6 * 4 + 23 * 8


def test_argument_validation_needs_no_gpu():
    from velocycle_b200 import _lib

    lib = _lib.load()
    p = _lib.VcbProblem()
    assert lib.vcb_workspace_bytes(C.byref(p)) == 0
    assert lib.vcb_phase_fwd_bwd(None, None, 0, None) == -1
    p.Nc, p.Ng, p.ld, p.H = 10, 8, 8, 3
    assert lib.vcb_workspace_bytes(C.byref(p)) > 0
    assert lib.vcb_phase_fwd_bwd(C.byref(p), None, 0, None) == -1  # required pointers missing


def test_product_refuses_cpu_tensors():
    import pytest
    import torch
    from velocycle_b200 import _lib
    from velocycle_b200.fused import PackedCounts

    with pytest.raises(_lib.VcbError):
        PackedCounts(torch.zeros(4, 4), None, 4)
