"""HostCounts: the compact pinned-host staging format of a count matrix and its widening on the device
(``vcb_expand_counts``; replaces the int64 host->device copy + ``.float()`` of preprocessing.py:142-143, 193-194).
Integer work: the round trip must be bit-exact."""
import pytest
import torch

from velocycle_b200 import _lib
from velocycle_b200.fused import HostCounts


def _matrix(kind, Nc=777, ld=52, seed=0):
    g = torch.Generator().manual_seed(seed)
    if kind == "small":      # every count < 255: one byte each, no escapes
        return torch.poisson(torch.full((Nc, ld), 2.0), generator=g)
    if kind == "escapes":    # a few large counts: one byte each + (index, value) list
        M = torch.poisson(torch.full((Nc, ld), 2.0), generator=g)
        M[3, 5], M[Nc - 1, ld - 1], M[100, 0] = 255.0, 70000.0, 1093.0
        return M
    if kind == "wide":       # counts in the hundreds everywhere: two bytes each
        return torch.poisson(torch.full((Nc, ld), 400.0), generator=g)
    if kind == "huge":       # beyond 16 bits: four bytes each
        M = torch.poisson(torch.full((Nc, ld), 400.0), generator=g)
        M[0, 0] = 1.0e6
        return M
    raise ValueError(kind)


@pytest.mark.parametrize("kind,fmt", [("small", _lib.VCB_COUNTS_U8), ("escapes", _lib.VCB_COUNTS_U8),
                                      ("wide", _lib.VCB_COUNTS_U16), ("huge", _lib.VCB_COUNTS_I32)])
def test_format_choice_and_host_packing(kind, fmt):
    M = _matrix(kind)
    h = HostCounts.from_tensor(M, chunk_rows=200)  # several chunks, ragged last one
    assert h.fmt == fmt and h.shape == tuple(M.shape)
    staged = h.staged.to(torch.int64)
    if fmt == _lib.VCB_COUNTS_U8:
        esc = M >= 255
        assert torch.equal(staged[~esc], M[~esc].to(torch.int64))
        assert bool((staged[esc] == 255).all())
        if esc.any():
            order = torch.argsort(h.over_idx)
            assert torch.equal(h.over_idx[order], esc.reshape(-1).nonzero().reshape(-1))
            assert torch.equal(h.over_val[order], M.reshape(-1)[esc.reshape(-1)])
            assert h.nbytes == M.numel() + 12 * int(esc.sum())
        else:
            assert h.over_idx is None and h.nbytes == M.numel()
    else:
        assert torch.equal(staged, M.to(torch.int64))
        assert h.nbytes == M.numel() * (2 if fmt == _lib.VCB_COUNTS_U16 else 4)


def test_unrepresentable_counts_are_refused():
    M = _matrix("small")
    M[1, 1] = float(2 ** 24)
    with pytest.raises(_lib.VcbError):
        HostCounts.from_tensor(M)


def test_upload_refuses_cpu_destination():
    h = HostCounts.from_tensor(_matrix("small"))
    with pytest.raises(_lib.VcbError):
        h.upload(torch.zeros(h.shape))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["small", "escapes", "wide", "huge"])
@pytest.mark.parametrize("shape", [(777, 52), (1, 4), (4099, 2000)])
def test_device_round_trip_is_bit_exact(kind, shape):
    M = _matrix(kind, *shape, seed=3) if shape[0] > 200 else _matrix("small", *shape, seed=3)
    h = HostCounts.from_tensor(M)
    dst = torch.full(M.shape, -1.0, device="cuda")
    h.upload(dst)
    h.upload(dst)  # idempotent, reuses the device staging buffers
    torch.cuda.synchronize()
    assert torch.equal(dst.cpu(), M)


@pytest.mark.gpu
def test_from_device_tensor_and_kernel_sees_identical_counts():
    """Packing from the device-resident matrix (what bench.py's e2e leg does) and feeding the widened copy to the
    fused path gives the same log-probs and gradients, bit for bit."""
    from velocycle_b200.fused import PackedCounts, fused_elbo_grad
    from velocycle_b200.synthetic import make_synthetic

    d = make_synthetic(3000, 203, H=2, Hw=1, Nb=1, Nx=1, seed=7, device="cuda")
    d.S[17, 3] = 4321.0
    counts = PackedCounts(d.S, d.U, d.Ng, d.batch_id, d.cond_id)
    args = (d.phi, d.cf, d.nu, None, d.shape_inv, d.logbeta, torch.exp(d.loggamma), d.nu_omega)
    ref = {k: v.clone() for k, v in fused_elbo_grad(counts, *args, grad=True).items()}
    hS, hU = HostCounts.from_tensor(d.S), HostCounts.from_tensor(d.U)
    assert hS.fmt == _lib.VCB_COUNTS_U8 and hS.over_idx is not None
    S2, U2 = torch.empty_like(d.S), torch.empty_like(d.U)
    hS.upload(S2)
    hU.upload(U2)
    assert torch.equal(S2, d.S) and torch.equal(U2, d.U)
    out = fused_elbo_grad(PackedCounts(S2, U2, d.Ng, d.batch_id, d.cond_id), *args, grad=True)
    for k in ref:
        if k.startswith("_"):  # scratch buffers kept alive with the result
            continue
        assert torch.equal(ref[k], out[k]), (k, float((ref[k] - out[k]).abs().max()), float(ref[k].abs().max()))


@pytest.mark.gpu
def test_expand_counts_argument_errors():
    lib = _lib.load()
    t = torch.zeros(64, device="cuda")
    assert lib.vcb_expand_counts(None, 1, 16, t.data_ptr(), None, None, 0, None) == -1
    assert lib.vcb_expand_counts(t.data_ptr(), 3, 16, t.data_ptr(), None, None, 0, None) == -2
    assert lib.vcb_expand_counts(t.data_ptr() + 4, 1, 16, t.data_ptr(), None, None, 0, None) == -3
    assert lib.vcb_expand_counts(t.data_ptr(), 2, 16, t.data_ptr(), None, None, 5, None) == -1  # overflow list only with u8


# ---- sub-byte staging (2 / 4 bits per entry + escape side stream; vcb_expand_counts_packed) -------------------------------
def _decode_sub_byte(h):
    """Pure-torch restatement of the formats (include/vcb.h) -- the oracle of the device decoders."""
    n = h.shape[0] * h.shape[1]
    w = h.staged.to(torch.int64) & 0xFFFFFFFF
    if h.fmt == _lib.VCB_COUNTS_B2N:   # 2-bit codes -> nibble stream -> byte stream
        per_block = 16 * _lib.VCB_PACKED_BLOCK_WORDS
        codes = ((w[:, None] >> (torch.arange(16) * 2)) & 3).reshape(-1)
        cnt1 = (codes == 3).reshape(-1, per_block).sum(1)
        assert torch.equal(h.block_off, torch.cumsum(cnt1, 0) - cnt1)
        e1 = (codes == 3).nonzero().reshape(-1)
        nw = h.nibbles.to(torch.int64) & 0xFFFFFFFF
        nib = ((nw[:, None] >> (torch.arange(8) * 4)) & 15).reshape(-1)[: e1.numel()]
        cnt2 = torch.zeros(cnt1.numel(), dtype=torch.int64).index_add_(0, e1 // per_block, (nib == 15).long())
        assert torch.equal(h.block_off2, torch.cumsum(cnt2, 0) - cnt2)
        vals = (nib + 3).float()
        e2 = (nib == 15).nonzero().reshape(-1)
        vals[e2] = h.side[: e2.numel()].float()
        out = codes.clone().float()
        out[e1] = vals
        out = out[:n]
    else:
        bits = h.bits
        per, E = 32 // bits, (1 << bits) - 1
        codes = ((w[:, None] >> (torch.arange(per) * bits)) & E).reshape(-1)
        cnt = (codes == E).reshape(-1, per * _lib.VCB_PACKED_BLOCK_WORDS).sum(1)
        assert torch.equal(h.block_off, torch.cumsum(cnt, 0) - cnt)       # escapes before each block
        out = codes[:n].clone().float()
        esc = (out == E).nonzero().reshape(-1)
        out[esc] = h.side[: esc.numel()].float()                             # escape bytes in entry order
    if h.over_idx is not None:
        out[h.over_idx] = h.over_val                                         # byte 255 -> overflow list
    return out.reshape(h.shape)


def _sparse_matrix(lam, Nc, ld, seed=0):
    """Poisson counts with a few planted values on every escape boundary (3, 17, 18, 19, 254, 255, 70000)."""
    g = torch.Generator().manual_seed(seed)
    M = torch.poisson(torch.full((Nc, ld), lam), generator=g)
    if Nc > 6 and ld > 6:
        M[3, 5], M[0, 0], M[5, 1], M[Nc - 1, ld - 1] = 255.0, 70000.0, 254.0, 16.0
        M[2, 2], M[4, 4], M[6, 6], M[6, 5] = 17.0, 18.0, 19.0, 3.0
    return M


@pytest.mark.parametrize("lam,shape,fmt", [(0.4, (3000, 2000), _lib.VCB_COUNTS_B2N), (2.0, (777, 52), _lib.VCB_COUNTS_B2N),
                                           (3.7, (4100, 200), _lib.VCB_COUNTS_B4), (0.4, (1, 4), _lib.VCB_COUNTS_B2),
                                           (0.9, (70001, 8), _lib.VCB_COUNTS_B2N),   # two packing chunks: nibble carry, ragged tail
                                           (3.7, (70001, 8), _lib.VCB_COUNTS_B4),
                                           (400.0, (300, 40), _lib.VCB_COUNTS_I32)])
def test_sub_byte_format_choice_and_host_packing(lam, shape, fmt):
    M = _sparse_matrix(lam, *shape)
    h = HostCounts.from_tensor(M, sub_byte=True)
    assert h.fmt == fmt
    if fmt in (_lib.VCB_COUNTS_B2, _lib.VCB_COUNTS_B2N, _lib.VCB_COUNTS_B4):
        assert torch.equal(_decode_sub_byte(h), M)
        assert h.nbytes < 0.62 * M.numel() + 8192        # well under one byte per entry
    assert HostCounts.from_tensor(M).fmt in (_lib.VCB_COUNTS_U8, _lib.VCB_COUNTS_U16, _lib.VCB_COUNTS_I32)   # opt-in only


@pytest.mark.gpu
@pytest.mark.parametrize("lam", [0.4, 2.0, 3.7])
@pytest.mark.parametrize("shape", [(777, 52), (1, 4), (4099, 2000), (70000, 100)])
def test_sub_byte_device_round_trip_is_bit_exact(lam, shape):
    M = _sparse_matrix(lam, *shape, seed=5)
    for src in (M, M.cuda()):                              # packed from a host or a device-resident matrix
        h = HostCounts.from_tensor(src, sub_byte=True)
        dst = torch.full(M.shape, -1.0, device="cuda")
        h.upload(dst)
        h.upload(dst)
        torch.cuda.synchronize()
        assert torch.equal(dst.cpu(), M), (h.fmt, shape)


@pytest.mark.gpu
def test_expand_counts_packed_argument_errors():
    lib = _lib.load()
    t = torch.zeros(4096, device="cuda")
    o = torch.zeros(8, dtype=torch.int64, device="cuda")
    assert lib.vcb_expand_counts_packed(None, 4, t.data_ptr(), o.data_ptr(), 16, t.data_ptr(), None, None, 0, None) == -1
    assert lib.vcb_expand_counts_packed(t.data_ptr(), 3, t.data_ptr(), o.data_ptr(), 16, t.data_ptr(), None, None, 0, None) == -2
    assert lib.vcb_expand_counts_packed(t.data_ptr(), 4, t.data_ptr(), o.data_ptr(), 16, t.data_ptr() + 4, None, None, 0, None) == -3
    assert lib.vcb_expand_counts_packed(t.data_ptr(), 2, t.data_ptr(), o.data_ptr(), 16, t.data_ptr(), None, None, 3, None) == -1
    assert lib.vcb_expand_counts_twolevel(t.data_ptr(), None, t.data_ptr(), o.data_ptr(), o.data_ptr(), 16, t.data_ptr(), None, None, 0, None) == -1
    assert lib.vcb_expand_counts_twolevel(t.data_ptr(), t.data_ptr(), t.data_ptr(), o.data_ptr(), o.data_ptr(), 16, t.data_ptr() + 4, None, None, 0, None) == -3
