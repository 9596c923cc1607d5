"""CSR -> device counts (``vcb_csr_to_counts`` / ``PackedCounts.pack_csr``): the sparse anndata layers go to the device
as CSR and are scattered into the float32 cell-major layout there, replacing the dense host copy of
preprocessing.py:138-143 / 243-249.  Integer work: the result must equal scipy's ``toarray()`` bit for bit."""
from types import SimpleNamespace

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from velocycle_b200 import _lib
from velocycle_b200.fused import PackedCounts
from velocycle_b200.preprocessing import _counts_layer, _pack


def _random_csr(Nc, Ng, density, dtype, seed=0, duplicates=False, shuffle=False):
    rng = np.random.default_rng(seed)
    M = sp.random(Nc, Ng, density=density, format="csr", random_state=rng,
                  data_rvs=lambda n: rng.integers(1, 300, size=n).astype(np.float64))
    M = M.astype(dtype)
    if duplicates:  # non-canonical CSR: the same (row, gene) twice -- toarray() sums them
        coo = M.tocoo()
        rows = np.concatenate([coo.row, coo.row[:50]])
        cols = np.concatenate([coo.col, coo.col[:50]])
        vals = np.concatenate([coo.data, coo.data[:50]])
        order = np.argsort(rows, kind="stable")
        indptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=Nc))])
        M = sp.csr_matrix((vals[order], cols[order], indptr), shape=(Nc, Ng))
    if shuffle:  # gene ids not sorted inside a row
        M = M.copy()
        for r in range(Nc):
            a, b = M.indptr[r], M.indptr[r + 1]
            p = rng.permutation(b - a)
            M.indices[a:b], M.data[a:b] = M.indices[a:b][p], M.data[a:b][p]
    return M


def test_sparse_layers_stay_sparse_and_cpu_devices_densify():
    M = _random_csr(40, 13, 0.3, np.float32)
    assert _counts_layer(M) is M                      # no dense host copy is made for a sparse layer
    dense = _counts_layer(M.toarray())
    assert dense.dtype == torch.int64
    out = _pack(M, "cpu")                             # (a CPU device serves the host-side guides only)
    assert out.shape == (40, 16) and torch.equal(out[:, :13], torch.as_tensor(M.toarray()).float())
    assert torch.count_nonzero(out[:, 13:]) == 0
    with pytest.raises(_lib.VcbError):
        PackedCounts.pack_csr(M, "cpu")               # the product path never falls back to the CPU


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.int32, np.float64, np.int64])
@pytest.mark.parametrize("shape", [(777, 50), (64, 2000), (5, 3), (1000, 1918)])
def test_csr_scatter_matches_toarray(dtype, shape):
    Nc, Ng = shape
    M = _random_csr(Nc, Ng, 0.2, dtype, seed=Nc + Ng, duplicates=(Nc >= 64), shuffle=True)
    out = PackedCounts.pack_csr(M, "cuda")
    ld = (Ng + 3) // 4 * 4
    assert out.shape == (Nc, ld) and out.dtype == torch.float32
    ref = torch.zeros(Nc, ld)
    ref[:, :Ng] = torch.as_tensor(np.asarray(M.toarray(), dtype=np.float64)).float()
    assert torch.equal(out.cpu(), ref)
    assert torch.equal(out.cpu(), PackedCounts.pack_matrix(torch.as_tensor(M.toarray()), "cells_by_genes"))


@pytest.mark.gpu
def test_csr_edge_cases():
    empty = sp.csr_matrix((33, 10), dtype=np.float32)            # nnz = 0
    assert torch.count_nonzero(PackedCounts.pack_csr(empty, "cuda")) == 0
    M = _random_csr(50, 20, 0.3, np.float32)
    M.data[7] = 2.5                                              # not a count
    with pytest.raises(_lib.VcbError):
        PackedCounts.pack_csr(M, "cuda")
    M = _random_csr(50, 20, 0.3, np.float32)
    bad = sp.csr_matrix((M.data, np.where(np.arange(M.nnz) == 3, 25, M.indices), M.indptr), shape=(50, 30))
    bad._shape = (50, 20)                                        # a gene id past Ng
    with pytest.raises(_lib.VcbError):
        PackedCounts.pack_csr(bad, "cuda")
    csc = _random_csr(60, 24, 0.2, np.int64).tocsc()             # any scipy format is converted to CSR first
    assert torch.equal(PackedCounts.pack_csr(csc, "cuda").cpu()[:, :24], torch.as_tensor(csc.toarray()).float())


@pytest.mark.gpu
def test_preprocess_with_sparse_layers_equals_dense_layers():
    """The anndata-facing wrapper builds the same metaparameters from sparse layers (CSR on the GPU) as from dense ones."""
    from velocycle_b200.preprocessing import preprocess_for_velocity_estimation

    Nc, Ng, H = 300, 37, 1
    S = _random_csr(Nc, Ng, 0.4, np.float32, seed=1)
    U = _random_csr(Nc, Ng, 0.15, np.float32, seed=2)
    g = torch.Generator().manual_seed(0)
    cyc = SimpleNamespace(means_tensor=torch.randn(2 * H + 1, Ng, generator=g), stds_tensor=torch.full((2 * H + 1, Ng), 0.5))
    ph = SimpleNamespace(phi_xy_tensor=torch.randn(2, Nc, generator=g))
    spd = SimpleNamespace(means_tensor=torch.tensor([[0.4], [0.0], [0.0]]), stds_tensor=torch.tensor([[0.1], [0.05], [0.05]]))
    D1 = torch.ones(Nc, 1, dtype=torch.int64)
    mps = []
    for layers in ({"spliced": S, "unspliced": U}, {"spliced": S.toarray(), "unspliced": U.toarray()}):
        ad = SimpleNamespace(layers=layers)
        mps.append(preprocess_for_velocity_estimation(ad, cyc, ph, spd, D1, D1, n_harmonics=H, device=torch.device("cuda")))
    a, b = mps
    assert torch.equal(a.S, b.S) and torch.equal(a.U, b.U) and torch.equal(a.count_factor, b.count_factor)
    assert torch.equal(a.packed_counts.S, b.packed_counts.S) and torch.equal(a.packed_counts.U, b.packed_counts.U)
