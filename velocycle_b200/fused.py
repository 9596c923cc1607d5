"""Host side of the fused ELBO+gradient op: packed device-resident counts, the raw C-ABI call and the
``torch.autograd.Function`` the model functions use.

Replaces, for the negative-binomial noise model, the (Ng,Nc) op chain of
``velocycle/phase_inference_model.py:368-393`` and ``velocycle/velocity_inference_model.py:344-386``
(Fourier basis -> ElogS/ElogU einsums -> ``GammaPoisson.log_prob`` -> autograd backward).
PyTorch is plumbing here (device memory, streams, autograd wiring); the arithmetic lives in
``csrc/`` behind ``include/vcb.h``.  No CPU path exists.
"""
from __future__ import annotations

import os

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import (VcbProblem, VcbSpectrum, VCB_FLAG_GRAD, VCB_FLAG_LEGACY_STREAM, VCB_FLAG_LGAMMA_INLINE,
                   VCB_FLAG_TCGEN05)
from .sharding import allreduce_flat_

__all__ = ["CountSpectrum", "HostCounts", "PackedCounts", "fused_elbo_grad", "FusedCycleNB", "fused_cycle_nb"]


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _dev_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.VcbError(f"{name} must be a CUDA tensor: velocycle_b200 has no CPU path")
    return t.detach().to(torch.float32).contiguous()


@dataclass
class CountSpectrum:
    """Per-gene histogram of one count matrix in CSR form (built once per dataset).

    ``lgamma(r+k) - lgamma(r) - lgamma(k+1)`` and ``digamma(r+k) - digamma(r)`` depend on the data only
    through how often each count value k occurs in a gene, so their sums over cells are evaluated per
    step from this spectrum (a few hundred thousand terms) instead of per cell and gene (billions).
    """

    off: torch.Tensor  # int32 [Ng+1]
    val: torch.Tensor  # float32 [nnz] distinct k > 0
    mult: torch.Tensor  # float32 [nnz] how many cells show that k
    lgk1: torch.Tensor  # float64 [Ng]  sum_c lgamma(k+1)
    max_count: int

    def as_struct(self) -> VcbSpectrum:
        return VcbSpectrum(self.off.data_ptr(), self.val.data_ptr(), self.mult.data_ptr(), self.lgk1.data_ptr())

    @staticmethod
    def build(M: torch.Tensor, Nc: int, Ng: int, ld: int) -> "CountSpectrum":
        """M: (Nc, ld) float32 CUDA, cell-major.  Raises on negative or non-integer counts."""
        lib = _lib.load()
        dev = M.device
        kmax = int(M.max().item()) if Nc > 0 else 0
        if kmax >= (1 << 20):
            raise _lib.VcbError(f"count value {kmax} too large for the histogram path; use inline=True")
        B = kmax + 2
        hist = torch.zeros((Ng, B), dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(
            lib.vcb_count_histogram(M.data_ptr(), Nc, Ng, ld, B, hist.data_ptr(), status.data_ptr(),
                                    torch.cuda.current_stream(dev).cuda_stream),
            "vcb_count_histogram",
        )
        if int(status.item()) != 0:
            raise _lib.VcbError(
                "counts must be non-negative integers (GammaPoisson.support, as Pyro's validation "
                "enforces for the reference); pass inline=True for real-valued pseudo-counts"
            )
        h = hist[:, 1:]  # k = 0 contributes nothing
        nz = (h > 0)
        per_gene = nz.sum(1)
        off = torch.zeros(Ng + 1, dtype=torch.int32, device=dev)
        off[1:] = per_gene.cumsum(0).to(torch.int32)
        idx = nz.nonzero(as_tuple=False)  # row-major: sorted by gene then k
        val = (idx[:, 1] + 1).to(torch.float32)
        mult = h[nz].to(torch.float32)
        ks = torch.arange(B, device=dev, dtype=torch.float64)
        lgk1 = (hist.to(torch.float64) * torch.lgamma(ks + 1.0)).sum(1)
        if val.numel() == 0:  # keep valid device pointers
            val = torch.zeros(1, dtype=torch.float32, device=dev)
            mult = torch.zeros(1, dtype=torch.float32, device=dev)
        return CountSpectrum(off.contiguous(), val.contiguous(), mult.contiguous(), lgk1.contiguous(), kmax)


class HostCounts:
    """Pinned-host staging of one count matrix in the narrowest exact integer format, and its upload.

    The reference ships the dense count matrix to the device as int64 and widens it there
    (``preprocessing.py:142-143, 193-194``: ``torch.tensor(S).to(device)`` ... ``.T.float()``): 8 bytes per entry
    over PCIe.  Counts are small integers, so the staging copy here holds one byte per entry (255 = escape, the
    few larger entries travel as an (index, value) list), two bytes when more than 2 % of the entries would
    escape, or four; ``upload`` copies it to the device and ``vcb_expand_counts`` widens it into the float32
    cell-major ``[Nc][ld]`` matrix the kernels stream.  Every count below 2**24 round-trips exactly.
    """

    ESCAPE = 255

    def __init__(self, staged: torch.Tensor, fmt: int, shape, over_idx=None, over_val=None, side=None, block_off=None):
        self.staged, self.fmt, self.shape = staged, fmt, tuple(shape)
        self.over_idx, self.over_val = over_idx, over_val
        self.side, self.block_off = side, block_off  # sub-byte formats: escape bytes in entry order, escapes before each block
        self.nibbles = self.block_off2 = None        # two-level 2-bit format: escape nibbles, second-level escapes before each block
        self._dev = None  # device staging buffers, created on first upload

    @property
    def nbytes(self) -> int:
        """Bytes that cross PCIe per upload."""
        n = self.staged.numel() * self.staged.element_size()
        if self.over_idx is not None:
            n += self.over_idx.numel() * 8 + self.over_val.numel() * 4
        if self.side is not None:
            n += self.side.numel() + self.block_off.numel() * 8
        if self.nibbles is not None:
            n += self.nibbles.numel() * 4 + self.block_off2.numel() * 8
        return n

    @property
    def bits(self) -> int:
        """Bits per entry of the main stream."""
        return {_lib.VCB_COUNTS_B2: 2, _lib.VCB_COUNTS_B2N: 2, _lib.VCB_COUNTS_B4: 4, _lib.VCB_COUNTS_U8: 8, _lib.VCB_COUNTS_U16: 16,
                _lib.VCB_COUNTS_I32: 32}[self.fmt]

    @staticmethod
    def choose_format(max_count: float, escape_fraction: float) -> int:
        if max_count >= 2 ** 24:
            raise _lib.VcbError("counts >= 2**24 are not exactly representable in the float32 device layout")
        if max_count < HostCounts.ESCAPE or escape_fraction <= 0.02:
            return _lib.VCB_COUNTS_U8
        return _lib.VCB_COUNTS_U16 if max_count < 2 ** 16 else _lib.VCB_COUNTS_I32

    @classmethod
    def from_tensor(cls, M: torch.Tensor, chunk_rows: int = 1 << 16, sub_byte: bool = False) -> "HostCounts":
        """``M``: (Nc, ld) non-negative integer-valued matrix (any real dtype, CPU or CUDA), already padded to the
        device pitch.  Packed chunk by chunk on the device the data lives on; the staging copy is pinned.
        ``sub_byte``: also consider the 2- and 4-bit formats (escape bytes in a side stream) and take the smallest."""
        assert M.dim() == 2
        Nc, ld = M.shape
        mx = float(M.max()) if M.numel() else 0.0
        n_esc = n3 = n15 = n18 = 0
        for r0 in range(0, Nc, chunk_rows):
            blk = M[r0: r0 + chunk_rows]
            n_esc += int((blk >= cls.ESCAPE).sum())
            if sub_byte:
                n3 += int((blk >= 3).sum())
                n15 += int((blk >= 15).sum())
                n18 += int((blk >= 18).sum())
        fmt = cls.choose_format(mx, n_esc / max(1, M.numel()))
        if sub_byte and M.numel() and ld % 4 == 0 and mx < 2 ** 24:
            n = M.numel()
            over = 12 * n_esc  # counts >= 255 travel as (int64 index, float32 value) pairs in every byte / sub-byte format
            size = {_lib.VCB_COUNTS_B2: n / 4 + n3 + over, _lib.VCB_COUNTS_B4: n / 2 + n15 + over,
                    _lib.VCB_COUNTS_B2N: n / 4 + n3 / 2 + n18 + over,
                    fmt: n * {_lib.VCB_COUNTS_U8: 1, _lib.VCB_COUNTS_U16: 2, _lib.VCB_COUNTS_I32: 4}[fmt]
                    + (over if fmt == _lib.VCB_COUNTS_U8 else 0)}
            best = min(size, key=size.get)
            if best == _lib.VCB_COUNTS_B2N:
                return cls._pack_two_level(M, n_esc, n3)
            if best in (_lib.VCB_COUNTS_B2, _lib.VCB_COUNTS_B4):
                return cls._pack_sub_byte(M, 2 if best == _lib.VCB_COUNTS_B2 else 4, best, n_esc)
        tdt = {_lib.VCB_COUNTS_U8: torch.uint8, _lib.VCB_COUNTS_U16: torch.uint16, _lib.VCB_COUNTS_I32: torch.int32}[fmt]
        pin = torch.cuda.is_available()
        staged = torch.empty((Nc, ld), dtype=tdt, pin_memory=pin)
        idx, val = [], []
        for r0 in range(0, Nc, chunk_rows):
            blk = M[r0: r0 + chunk_rows]
            if fmt == _lib.VCB_COUNTS_U8:
                esc = blk >= cls.ESCAPE
                if n_esc:
                    nz = esc.reshape(-1).nonzero().reshape(-1)
                    idx.append((nz + r0 * ld).to(torch.int64).cpu())
                    val.append(blk.reshape(-1)[nz].to(torch.float32).cpu())
                staged[r0: r0 + chunk_rows].copy_(torch.where(esc, cls.ESCAPE, blk).to(torch.uint8))
            elif fmt == _lib.VCB_COUNTS_U16:
                staged[r0: r0 + chunk_rows].copy_(blk.to(torch.int32).to(torch.uint16))
            else:
                staged[r0: r0 + chunk_rows].copy_(blk.to(torch.int32))
        over_idx = over_val = None
        if fmt == _lib.VCB_COUNTS_U8 and n_esc:
            over_idx = torch.cat(idx)
            over_val = torch.cat(val)
            if pin:
                over_idx, over_val = over_idx.pin_memory(), over_val.pin_memory()
        return cls(staged, fmt, (Nc, ld), over_idx, over_val)

    @classmethod
    def _pack_sub_byte(cls, M: torch.Tensor, bits: int, fmt: int, n_over: int) -> "HostCounts":
        """2- or 4-bit codes, 32/bits per little-endian int32 word; code 2^bits-1 = escape -> next byte of the side stream;
        side byte 255 -> (index, value) overflow list (format: include/vcb.h, vcb_expand_counts_packed)."""
        Nc, ld = M.shape
        n = Nc * ld
        per, E = 32 // bits, (1 << bits) - 1
        per_block = per * _lib.VCB_PACKED_BLOCK_WORDS
        n_blocks = (n + per_block - 1) // per_block
        pin = torch.cuda.is_available()
        codes = torch.zeros(n_blocks * _lib.VCB_PACKED_BLOCK_WORDS, dtype=torch.int32, pin_memory=pin)
        esc_per_block = torch.zeros(n_blocks, dtype=torch.int64)
        rows = max(1024, (1 << 16) // 1024 * 1024)  # rows * ld is a whole number of blocks (ld % 4 == 0)
        shifts = (torch.arange(per, dtype=torch.int64) * bits)
        side, idx, val = [], [], []
        for r0 in range(0, Nc, rows):
            flat = M[r0: r0 + rows].reshape(-1)
            e0 = r0 * ld
            pad = (-flat.numel()) % per_block
            if pad:
                flat = torch.cat([flat, flat.new_zeros(pad)])
            esc = flat >= E
            side.append(flat[esc].clamp(max=cls.ESCAPE).to(torch.uint8).cpu())
            if n_over:
                nz = (flat >= cls.ESCAPE).nonzero().reshape(-1)
                idx.append((nz + e0).to(torch.int64).cpu())
                val.append(flat[nz].to(torch.float32).cpu())
            c = flat.clamp(max=E).to(torch.int64).reshape(-1, per)
            words = (c << shifts.to(c.device)).sum(1)
            words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
            w0 = e0 // per
            codes[w0: w0 + words.numel()].copy_(words)
            b0 = e0 // per_block
            cnt = esc.reshape(-1, per_block).sum(1).to(torch.int64).cpu()
            esc_per_block[b0: b0 + cnt.numel()] = cnt
        side = torch.cat(side) if side else torch.zeros(0, dtype=torch.uint8)
        if side.numel() == 0:
            side = torch.zeros(1, dtype=torch.uint8)  # (a valid pointer for the kernel)
        block_off = torch.cumsum(esc_per_block, 0) - esc_per_block
        over_idx = torch.cat(idx) if n_over else None
        over_val = torch.cat(val) if n_over else None
        if pin:
            side, block_off = side.pin_memory(), block_off.pin_memory()
            if n_over:
                over_idx, over_val = over_idx.pin_memory(), over_val.pin_memory()
        return cls(codes, fmt, (Nc, ld), over_idx, over_val, side=side, block_off=block_off)

    @classmethod
    def _pack_two_level(cls, M: torch.Tensor, n_over: int, n3: int) -> "HostCounts":
        """2-bit codes (3 = escape) -> nibble stream (value-3; 15 = escape) -> byte stream (value; 255 -> overflow list).
        Format: include/vcb.h, vcb_expand_counts_twolevel."""
        Nc, ld = M.shape
        n = Nc * ld
        per, per_block = 16, 16 * _lib.VCB_PACKED_BLOCK_WORDS
        n_blocks = (n + per_block - 1) // per_block
        pin = torch.cuda.is_available()
        codes = torch.zeros(n_blocks * _lib.VCB_PACKED_BLOCK_WORDS, dtype=torch.int32, pin_memory=pin)
        nib_words = torch.zeros(max(1, (n3 + 7) // 8), dtype=torch.int32, pin_memory=pin)
        esc1 = torch.zeros(n_blocks, dtype=torch.int64)
        esc2 = torch.zeros(n_blocks, dtype=torch.int64)
        rows = (1 << 16) // 1024 * 1024
        dev = M.device
        sh2 = (torch.arange(16, dtype=torch.int64, device=dev) * 2)
        sh4 = (torch.arange(8, dtype=torch.int64, device=dev) * 4)
        to_i32 = lambda w: torch.where(w >= 2 ** 31, w - 2 ** 32, w).to(torch.int32)
        carry = torch.zeros(0, dtype=torch.int64, device=dev)  # nibbles that did not fill a word yet
        nib_done = 0
        side, idx, val = [], [], []
        for r0 in range(0, Nc, rows):
            flat = M[r0: r0 + rows].reshape(-1)
            e0 = r0 * ld
            pad = (-flat.numel()) % per_block
            if pad:
                flat = torch.cat([flat, flat.new_zeros(pad)])
            m1 = flat >= 3
            v1 = flat[m1]
            m2 = v1 >= 18
            side.append(v1[m2].clamp(max=cls.ESCAPE).to(torch.uint8).cpu())
            if n_over:
                nz = (flat >= cls.ESCAPE).nonzero().reshape(-1)
                idx.append((nz + e0).to(torch.int64).cpu())
                val.append(flat[nz].to(torch.float32).cpu())
            words = to_i32((flat.clamp(max=3).to(torch.int64).reshape(-1, per) << sh2).sum(1))
            w0 = e0 // per
            codes[w0: w0 + words.numel()].copy_(words)
            b0 = e0 // per_block
            c1 = m1.reshape(-1, per_block).sum(1).to(torch.int64)
            esc1[b0: b0 + c1.numel()] = c1.cpu()
            # second-level escapes per block: scatter-add the flags of the escaped entries into their blocks
            blk_of = (m1.nonzero().reshape(-1) // per_block)
            c2 = torch.zeros(c1.numel(), dtype=torch.int64, device=dev).index_add_(0, blk_of, m2.to(torch.int64))
            esc2[b0: b0 + c2.numel()] = c2.cpu()
            nibs = torch.cat([carry, (v1 - 3).clamp(max=15).to(torch.int64)])
            whole = nibs.numel() // 8 * 8
            if whole:
                nw = to_i32((nibs[:whole].reshape(-1, 8) << sh4).sum(1))
                nib_words[nib_done: nib_done + nw.numel()].copy_(nw)
                nib_done += nw.numel()
            carry = nibs[whole:]
        if carry.numel():
            last = torch.cat([carry, carry.new_zeros(8 - carry.numel())])
            nib_words[nib_done: nib_done + 1].copy_(to_i32((last.reshape(1, 8) << sh4).sum(1)))
        side = torch.cat(side) if side else torch.zeros(0, dtype=torch.uint8)
        if side.numel() == 0:
            side = torch.zeros(1, dtype=torch.uint8)
        off1 = torch.cumsum(esc1, 0) - esc1
        off2 = torch.cumsum(esc2, 0) - esc2
        over_idx = torch.cat(idx) if n_over else None
        over_val = torch.cat(val) if n_over else None
        if pin:
            side, off1, off2 = side.pin_memory(), off1.pin_memory(), off2.pin_memory()
            if n_over:
                over_idx, over_val = over_idx.pin_memory(), over_val.pin_memory()
        h = cls(codes, _lib.VCB_COUNTS_B2N, (Nc, ld), over_idx, over_val, side=side, block_off=off1)
        h.nibbles, h.block_off2 = nib_words, off2
        return h

    # ---- device side: two sets of staging buffers, so that the H2D copy of the next upload overlaps the work on the last one
    def _device_set(self, dev, k: int):
        if self._dev is None or self._dev[0][0].device != dev:
            def mk(t):
                return None if t is None else torch.empty_like(t, device=dev)
            self._dev = [[mk(self.staged), mk(self.over_idx), mk(self.over_val), mk(self.side), mk(self.block_off), None,
                          mk(self.nibbles), mk(self.block_off2)] for _ in range(2)]
            self._next, self._pending = 0, None
        return self._dev[k]

    def start_upload(self, device, stream: Optional["torch.cuda.Stream"] = None) -> None:
        """Enqueue the H2D copies of the staging buffers on ``stream`` (default: the current stream) into the spare device
        buffer set.  With a dedicated copy stream they overlap whatever the compute stream is doing (the previous SVI step);
        ``finish_upload`` makes the compute stream wait for them.  Nothing synchronises the host."""
        dev = torch.device(device)
        if dev.type != "cuda":
            raise _lib.VcbError("HostCounts uploads need a CUDA device: velocycle_b200 has no CPU path")
        k = getattr(self, "_next", 0)
        bufs = self._device_set(dev, k)
        k = self._next
        bufs = self._dev[k]
        stream = stream or torch.cuda.current_stream(dev)
        if bufs[5] is not None:
            stream.wait_event(bufs[5])  # the widening kernel that last read this set has finished
        with torch.cuda.stream(stream):
            for d, h in zip(bufs[:5] + bufs[6:8], (self.staged, self.over_idx, self.over_val, self.side, self.block_off,
                                                   self.nibbles, self.block_off2)):
                if d is not None:
                    d.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
        self._pending = (k, ev)
        self._next = 1 - k

    def finish_upload(self, dst: torch.Tensor) -> None:
        """Current stream: wait for the copies of the last ``start_upload``, widen them into ``dst`` (float32 CUDA (Nc, ld))."""
        if not dst.is_cuda:
            raise _lib.VcbError("HostCounts.upload needs a CUDA destination: velocycle_b200 has no CPU path")
        assert dst.dtype == torch.float32 and dst.is_contiguous() and tuple(dst.shape) == self.shape
        if getattr(self, "_pending", None) is None:
            raise _lib.VcbError("finish_upload without start_upload")
        dev = dst.device
        k, ev = self._pending
        self._pending = None
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ev)
        d_st, d_i, d_v, d_side, d_off, _, d_nib, d_off2 = self._dev[k]
        n_over = 0 if d_i is None else d_i.numel()
        lib = _lib.load()
        if d_nib is not None:
            _lib.check(lib.vcb_expand_counts_twolevel(d_st.data_ptr(), d_nib.data_ptr(), d_side.data_ptr(), d_off.data_ptr(),
                                                      d_off2.data_ptr(), dst.numel(), dst.data_ptr(), _ptr(d_i), _ptr(d_v), n_over,
                                                      cur.cuda_stream), "vcb_expand_counts_twolevel")
        elif d_side is not None:
            _lib.check(lib.vcb_expand_counts_packed(d_st.data_ptr(), self.bits, d_side.data_ptr(), d_off.data_ptr(), dst.numel(),
                                                    dst.data_ptr(), _ptr(d_i), _ptr(d_v), n_over, cur.cuda_stream),
                       "vcb_expand_counts_packed")
        else:
            _lib.check(lib.vcb_expand_counts(d_st.data_ptr(), self.fmt, dst.numel(), dst.data_ptr(), _ptr(d_i), _ptr(d_v),
                                             n_over, cur.cuda_stream), "vcb_expand_counts")
        done = torch.cuda.Event()
        done.record(cur)
        self._dev[k][5] = done

    def upload(self, dst: torch.Tensor) -> None:
        """H2D copy of the staging buffers + widening into ``dst`` (float32 CUDA (Nc, ld), contiguous), all on the
        current stream; nothing synchronises."""
        if not dst.is_cuda:
            raise _lib.VcbError("HostCounts.upload needs a CUDA destination: velocycle_b200 has no CPU path")
        self.start_upload(dst.device)
        self.finish_upload(dst)


class PackedCounts:
    """Device-resident spliced / unspliced counts in the layout the kernels stream.

    Cell-major float32 ``[Nc][ld]`` with ``ld = roundup(Ng, 4)`` (16-byte rows for 128-bit loads and
    TMA bulk copies), zero padding columns, int32 batch / condition ids, and the count spectra.
    ``from_model_tensors`` accepts the reference's ``mp.S`` / ``mp.U`` (logical (Ng,Nc), physically
    cell-major, ``preprocessing.py:193-194``) and the one-hot design tensors ``mp.Db`` / ``mp.D``.
    """

    def __init__(self, S: torch.Tensor, U: Optional[torch.Tensor], Ng: int,
                 batch_id: Optional[torch.Tensor] = None, cond_id: Optional[torch.Tensor] = None,
                 spectrum: bool = True):
        if not S.is_cuda:
            raise _lib.VcbError("PackedCounts needs CUDA tensors: velocycle_b200 has no CPU path")
        assert S.dim() == 2 and S.dtype == torch.float32 and S.is_contiguous()
        self.Nc, self.ld = int(S.shape[0]), int(S.shape[1])
        self.Ng = int(Ng)
        assert self.ld % 4 == 0 and self.ld >= self.Ng
        assert S.data_ptr() % 16 == 0
        self.S, self.U = S, U
        if U is not None:
            assert U.shape == S.shape and U.dtype == torch.float32 and U.is_contiguous() and U.data_ptr() % 16 == 0
        dev = S.device
        self.batch_id = None if batch_id is None else batch_id.to(device=dev, dtype=torch.int32).contiguous()
        self.cond_id = None if cond_id is None else cond_id.to(device=dev, dtype=torch.int32).contiguous()
        # The streaming kernel switches the batch offsets per 16-cell stage; a stage that mixes batches takes a masked pass
        # per batch present.  Design matrices come in arbitrary cell order (preprocessing.py:65-93), so cells are stably
        # sorted by batch ONCE here: rows of S / U and the ids are reordered (a copy, only when the input is not sorted
        # already), `perm[i]` = caller's index of the cell in row i.  Per-cell inputs / outputs of a call are permuted at the
        # wrapper (fused_elbo_grad) or inside the fused step's cell kernels; after sorting at most Nb - 1 stages are mixed.
        self.perm = self.inv_perm = None
        if self.batch_id is not None and self.Nc > 1 and bool((self.batch_id[1:] < self.batch_id[:-1]).any()):
            self.perm = torch.argsort(self.batch_id.long(), stable=True)
            self.inv_perm = torch.empty_like(self.perm)
            self.inv_perm[self.perm] = torch.arange(self.Nc, device=dev)
            self.S = self.S.index_select(0, self.perm)
            if self.U is not None:
                self.U = self.U.index_select(0, self.perm)
            self.batch_id = self.batch_id[self.perm].contiguous()
            if self.cond_id is not None:
                self.cond_id = self.cond_id[self.perm].contiguous()
        self.spec_S = self.spec_U = None
        if spectrum:
            self.build_spectra()

    def build_spectra(self) -> None:
        if self.spec_S is None:
            self.spec_S = CountSpectrum.build(self.S, self.Nc, self.Ng, self.ld)
        if self.U is not None and self.spec_U is None:
            self.spec_U = CountSpectrum.build(self.U, self.Nc, self.Ng, self.ld)

    @property
    def device(self):
        return self.S.device

    @staticmethod
    def pack_matrix(M: torch.Tensor, layout: str = "genes_by_cells") -> torch.Tensor:
        """Return a (Nc, roundup(Ng,4)) float32 contiguous copy.  ``layout`` names the LOGICAL shape of M."""
        if layout == "genes_by_cells":
            M = M.T  # logical (Nc, Ng); for the reference's mp.S this is the physical layout already
        elif layout != "cells_by_genes":
            raise ValueError(f"{layout=}")
        Nc, Ng = M.shape
        ld = (Ng + 3) // 4 * 4
        out = torch.zeros((Nc, ld), dtype=torch.float32, device=M.device)
        out[:, :Ng] = M
        return out

    @staticmethod
    def pack_csr(M, device) -> torch.Tensor:
        """(Nc cells, Ng genes) scipy.sparse matrix -> the same (Nc, roundup(Ng,4)) float32 device matrix as ``pack_matrix``,
        built ON the device by ``vcb_csr_to_counts`` from the three CSR arrays: the dense int64 host copy the reference
        makes (``preprocessing.py:138-143, 243-249``: 8 bytes x Nc x Ng of host memory and PCIe) never exists."""
        import numpy as np

        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.VcbError("pack_csr builds the count matrix on a CUDA device (there is no CPU path)")
        lib = _lib.load()
        csr = M.tocsr()
        Nc, Ng = csr.shape
        ld = (Ng + 3) // 4 * 4
        out = torch.empty((Nc, ld), dtype=torch.float32, device=device)
        if Nc == 0:
            return out
        kinds = {np.dtype(np.float32): _lib.VCB_CSR_F32, np.dtype(np.int32): _lib.VCB_CSR_I32,
                 np.dtype(np.float64): _lib.VCB_CSR_F64, np.dtype(np.int64): _lib.VCB_CSR_I64}
        data = csr.data
        if data.dtype not in kinds:
            data = data.astype(np.float64 if data.dtype.kind == "f" else np.int64)
        indptr = torch.from_numpy(np.ascontiguousarray(csr.indptr, dtype=np.int64)).to(device, non_blocking=True)
        nnz = int(csr.indptr[-1])
        indices = torch.from_numpy(np.ascontiguousarray(csr.indices[:nnz], dtype=np.int32)).to(device) if nnz else None
        values = torch.from_numpy(np.ascontiguousarray(data[:nnz])).to(device) if nnz else None
        status = torch.zeros(1, dtype=torch.int32, device=device)
        rc = lib.vcb_csr_to_counts(indptr.data_ptr(), _ptr(indices), _ptr(values), kinds[np.dtype(data.dtype)], Nc, Ng, ld,
                                   out.data_ptr(), status.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
        _lib.check(rc, "vcb_csr_to_counts")
        if int(status.item()) != 0:
            raise _lib.VcbError("the sparse count matrix holds gene ids outside [0, Ng) or values that are not integers in [0, 2^24)")
        return out

    @staticmethod
    def one_hot_to_ids(D: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """mp.Db (Nb,1,Nc) / (Nb,1,1,1,Nc) or mp.D (Nx,1,1,Nc) one-hot -> int32 ids (Nc,)."""
        if D is None:
            return None
        D2 = D.reshape(D.shape[0], -1)
        return D2.argmax(0).to(torch.int32)

    @classmethod
    def from_model_tensors(cls, S, U=None, Db=None, D=None, spectrum: bool = True) -> "PackedCounts":
        Ng = int(S.shape[0])
        Sp = cls.pack_matrix(S)
        Up = None if U is None else cls.pack_matrix(U)
        return cls(Sp, Up, Ng, cls.one_hot_to_ids(Db), cls.one_hot_to_ids(D), spectrum=spectrum)


def fused_elbo_grad(
    counts: PackedCounts,
    phi: torch.Tensor,
    cf: Optional[torch.Tensor],
    nu: torch.Tensor,
    dnu: Optional[torch.Tensor],
    shape_inv: torch.Tensor,
    logbeta: Optional[torch.Tensor] = None,
    gamma: Optional[torch.Tensor] = None,
    nu_omega: Optional[torch.Tensor] = None,
    grad: bool = True,
    inline_lgamma: bool = False,
    want_d_omega: bool = False,
    tcgen05: bool = False,
    legacy_stream: bool = False,
) -> Dict[str, torch.Tensor]:
    """One launch sequence of the fused path through the C ABI.  All tensors float32 CUDA.

    velocity model iff ``nu_omega`` is given.  Returns ``lp_S`` [Ng] (``lp_U``) and, under ``grad``, the
    gradients of ``sum(lp_S) + sum(lp_U)`` keyed like ``include/vcb.h``.
    """
    lib = _lib.load()
    dev = counts.device
    velocity = nu_omega is not None
    Nc, Ng, ld = counts.Nc, counts.Ng, counts.ld
    phi = _dev_f32(phi, "phi").reshape(-1)
    if counts.perm is not None:  # rows are sorted by batch: per-cell inputs follow
        phi = phi[counts.perm]
    nu = _dev_f32(nu, "nu").reshape(Ng, -1)
    K = nu.shape[1]
    shape_inv = _dev_f32(shape_inv, "shape_inv").reshape(-1)
    cf = None if cf is None else _dev_f32(cf, "cf").reshape(-1)
    if cf is not None and counts.perm is not None:
        cf = cf[counts.perm]
    assert phi.numel() == Nc and shape_inv.numel() == Ng and K % 2 == 1
    Nb = 0
    if dnu is not None:
        dnu = _dev_f32(dnu, "dnu").reshape(-1, Ng)
        Nb = dnu.shape[0]
        if counts.batch_id is None:
            if Nb != 1:
                raise _lib.VcbError("dnu with more than one batch needs batch ids")
            counts.batch_id = torch.zeros(Nc, dtype=torch.int32, device=dev)
    p = VcbProblem()
    p.Nc, p.Ng, p.ld = Nc, Ng, ld
    p.H, p.Nb = (K - 1) // 2, Nb
    p.flags = (VCB_FLAG_GRAD if grad else 0) | (VCB_FLAG_LGAMMA_INLINE if inline_lgamma else 0) | (VCB_FLAG_TCGEN05 if tcgen05 else 0) | (VCB_FLAG_LEGACY_STREAM if legacy_stream else 0)
    p.S = counts.S.data_ptr()
    p.phi, p.cf = phi.data_ptr(), _ptr(cf)
    p.batch_id = _ptr(counts.batch_id) if Nb > 0 else None
    p.nu, p.dnu, p.shape_inv = nu.data_ptr(), _ptr(dnu), shape_inv.data_ptr()
    out: Dict[str, torch.Tensor] = {}
    f32 = dict(dtype=torch.float32, device=dev)
    Nx = Kw = 0
    if velocity:
        if counts.U is None:
            raise _lib.VcbError("velocity model needs unspliced counts")
        logbeta = _dev_f32(logbeta, "logbeta").reshape(-1)
        gamma = _dev_f32(gamma, "gamma").reshape(-1)
        nu_omega = _dev_f32(nu_omega, "nu_omega")
        nu_omega = nu_omega.reshape(nu_omega.shape[0], -1)
        Nx, Kw = nu_omega.shape
        if Nx > 1 and counts.cond_id is None:
            raise _lib.VcbError("more than one condition needs condition ids")
        p.Hw, p.Nx = (Kw - 1) // 2, Nx
        p.U = counts.U.data_ptr()
        p.cond_id = _ptr(counts.cond_id)
        p.logbeta, p.gamma, p.nu_omega = logbeta.data_ptr(), gamma.data_ptr(), nu_omega.data_ptr()
    else:
        p.Hw, p.Nx = 0, 0
    # every gene-level / global output lives in ONE flat buffer: under cell sharding it is what the single
    # all-reduce of the step moves (sharding.py)
    sizes = [("lp_S", Ng), ("lp_U", Ng if velocity else 0)]
    if grad:
        sizes += [("d_shape_inv", Ng), ("d_logbeta", Ng if velocity else 0), ("d_gamma", Ng if velocity else 0),
                  ("d_nu", Ng * K), ("d_dnu", Nb * Ng), ("d_nu_omega", Nx * Kw if velocity else 0)]
    flat = torch.empty(sum(n for _, n in sizes), **f32)
    off = 0
    for name, n in sizes:
        if n:
            out[name] = flat[off: off + n]
            setattr(p, name, out[name].data_ptr())
        off += n
    out["gene_flat"] = flat
    if grad:
        out["d_nu"] = out["d_nu"].view(Ng, K)
        if Nb > 0:
            out["d_dnu"] = out["d_dnu"].view(Nb, Ng)
        if velocity:
            out["d_nu_omega"] = out["d_nu_omega"].view(Nx, Kw)
        out["d_phi"] = torch.empty(Nc, **f32)
        out["d_cf"] = torch.empty(Nc, **f32)
        p.d_phi, p.d_cf = out["d_phi"].data_ptr(), out["d_cf"].data_ptr()
        if velocity and want_d_omega:
            out["d_omega"] = torch.empty(Nc, **f32)
            p.d_omega = out["d_omega"].data_ptr()
    if not inline_lgamma:
        counts.build_spectra()
        p.spec_S = counts.spec_S.as_struct()
        if velocity:
            p.spec_U = counts.spec_U.as_struct()
    ev = getattr(counts, "profile_events", None)  # (begin, end) raw cudaEvent_t handles, set by bench.py
    if ev is not None:
        p.ev_stream_begin, p.ev_stream_end = ev
    ws_bytes = lib.vcb_workspace_bytes(C.byref(p))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    if os.environ.get("VCB_DEBUG_POISON_WS"):  # every byte 0xFF = NaN: an uninitialised read shows up in the outputs
        ws.fill_(255)
    fn = lib.vcb_velocity_fwd_bwd if velocity else lib.vcb_phase_fwd_bwd
    _lib.check(fn(C.byref(p), ws.data_ptr(), ws_bytes, torch.cuda.current_stream(dev).cuda_stream),
               "vcb_velocity_fwd_bwd" if velocity else "vcb_phase_fwd_bwd")
    out["_workspace"] = ws  # keep alive until the stream has consumed it
    if counts.perm is not None:  # per-cell outputs back in the caller's cell order
        for k in ("d_phi", "d_cf", "d_omega"):
            if k in out:
                out[k] = out[k][counts.inv_perm]
    return out


class FusedCycleNB(torch.autograd.Function):
    """log-likelihood of the count matrices as an autograd node with a fused forward+backward.

    ``forward`` returns the per-gene log-prob sums (lp_S, lp_U) and computes every gradient in the same
    streaming pass; ``backward`` only scales the saved gradients.  That is exact when all genes of both
    sites carry the same upstream weight, which is what ``Trace_ELBO`` produces (-1 for every site).
    Any other weighting poisons the result with NaN instead of returning a silently wrong gradient.
    """

    @staticmethod
    def forward(ctx, counts, inline_lgamma, phi, cf, nu, dnu, shape_inv, logbeta, gamma, nu_omega):
        velocity = nu_omega is not None
        needs = ctx.needs_input_grad[2:]
        grad = any(needs)
        out = fused_elbo_grad(counts, phi, cf, nu, dnu, shape_inv, logbeta, gamma, nu_omega,
                              grad=grad, inline_lgamma=inline_lgamma)
        allreduce_flat_(out["gene_flat"], getattr(counts, "shard", None))  # the step's single exchange
        ctx.velocity = velocity
        ctx.grad_computed = grad
        ctx.shapes = [None if t is None else t.shape for t in (phi, cf, nu, dnu, shape_inv, logbeta, gamma, nu_omega)]
        if grad:
            names = ["d_phi", "d_cf", "d_nu", "d_dnu", "d_shape_inv", "d_logbeta", "d_gamma", "d_nu_omega"]
            ctx.save_for_backward(*[out.get(n) if out.get(n) is not None else torch.empty(0, device=counts.device)
                                    for n in names])
        lpS = out["lp_S"]
        lpU = out["lp_U"] if velocity else torch.zeros_like(lpS)
        return lpS, lpU

    @staticmethod
    def backward(ctx, g_lpS, g_lpU):
        if not ctx.grad_computed:
            return (None,) * 10
        saved = ctx.saved_tensors
        s = g_lpS.reshape(-1)[0]
        bad = (g_lpS != s).any()
        if ctx.velocity:
            bad = bad | (g_lpU != s).any()
        scale = torch.where(bad, torch.full_like(s, float("nan")), s)
        grads = []
        for i, (t, shp) in enumerate(zip(saved, ctx.shapes)):
            if shp is None or t.numel() == 0 or not ctx.needs_input_grad[2 + i]:
                grads.append(None)
            else:
                grads.append((t * scale).reshape(shp))
        return (None, None, *grads)


def fused_cycle_nb(counts: PackedCounts, phi, cf, nu, dnu, shape_inv, logbeta=None, gamma=None, nu_omega=None,
                   inline_lgamma: bool = False):
    """Functional wrapper: returns (lp_S[Ng], lp_U[Ng])."""
    return FusedCycleNB.apply(counts, inline_lgamma, phi, cf, nu, dnu, shape_inv, logbeta, gamma, nu_omega)
