"""Nearest-code search over row-major vectors on the same sm_100a kernels as the CodeBook.

Mirrors ``GaussianDiffusion2D.gaussian_to_indices``
(/root/reference/network/vqDiffusion/submodule/diffusion_gaussian2d.py:322-347): the vectors ``(B, L, D)`` are flattened
to rows, the squared distance to every table row is ``|x|^2 + |e|^2 - 2 x.e`` in fp32 and the first minimum wins --
exactly the CodeBook's formula (codebook.py:70-82), so the same tcgen05 distance GEMM + exact fp32 re-rank decide it,
fed through the row-major entry point ``vq_argmin_rows`` of include/vq_b200.h.

The kernels are specialised for 256-wide vectors.  Narrower tables (``gaussian_dim: 96`` in configs/*.yml) are
zero-padded to 256 columns: a zero column adds exactly 0 to every dot product and norm (``fma(0, 0, p) == p``), so no
distance -- and no tie -- changes.  There is no CPU path.
"""
from __future__ import annotations

import torch

from . import _native

__all__ = ["CodeTable", "nearest_indices"]

_D = 256
_IDX = {torch.int64: 64, torch.int32: 32, torch.int16: 16, torch.uint16: 16}


def _index_bits(dtype, K: int) -> int:
    if dtype not in _IDX:
        raise ValueError(f"index dtype must be int64, int32, int16 or uint16, got {dtype}")
    if dtype == torch.int16 and K > 32768:
        raise ValueError(f"int16 indices need K <= 32768 (got K={K}); use torch.uint16 or int32")
    if dtype == torch.uint16 and K > 65536:
        raise ValueError(f"uint16 indices need K <= 65536 (got K={K})")
    return _IDX[dtype]


class CodeTable:
    """A lookup table ``(K, D <= 256)`` prepared for nearest-row queries (fp16 operand image, |e|^2, scalars).

    Build it once per table (``gaussian_lookup_table`` is a fixed buffer, diffusion_gaussian2d.py:287) and call
    :meth:`nearest` per batch; ``refresh()`` after the table changed.
    """

    def __init__(self, table: torch.Tensor):
        if table.dim() != 2 or table.shape[1] > _D:
            raise ValueError(f"table must be (K, D <= {_D}), got {tuple(table.shape)}")
        if not table.is_cuda or table.dtype != torch.float32:
            raise RuntimeError("CodeTable needs a CUDA float32 table (no CPU path)")
        self.table = table
        self.K, self.D = table.shape
        self._ws = None
        self.last_stats = None
        self.refresh()

    def refresh(self):
        t = self.table.detach()
        dev = t.device
        if self.D < _D:
            E = torch.zeros((self.K, _D), dtype=torch.float32, device=dev)
            E[:, : self.D] = t
        else:
            E = t.contiguous()
        k_pad = _native.padded_codes(self.K)
        self._E = E
        self._E_h = torch.empty((k_pad, _D), dtype=torch.float16, device=dev)
        self._e2 = torch.empty((k_pad,), dtype=torch.float32, device=dev)
        self._cb = torch.empty((4,), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = _native.lib().vq_prepare_codebook(E.data_ptr(), self.K, _D, self._E_h.data_ptr(), self._e2.data_ptr(),
                                                   self._cb.data_ptr(), int(torch.cuda.current_stream(dev).cuda_stream))
        _native.check(rc, "vq_prepare_codebook")

    @torch.no_grad()
    def nearest(self, x: torch.Tensor, dtype=torch.int64, recipe: str = "expanded") -> torch.Tensor:
        """``argmin_k |x - table[k]|^2`` for every vector along the last axis of ``x`` -> indices of shape x.shape[:-1].

        ``recipe`` names the reference's fp32 formula whose argmin (ties, rounding and all) is reproduced:
        ``"expanded"`` = ``|x|^2 + |e|^2 - 2 x.e`` (codebook.py:70-79, diffusion_gaussian2d.py:334-339), ``"diffsq"`` =
        ``sum((x - e)**2)`` (v_vq_diffusion.py:114-123)."""
        if recipe not in _native.VQ_RECIPES:
            raise ValueError(f"recipe must be one of {sorted(_native.VQ_RECIPES)}, got {recipe!r}")
        if x.shape[-1] != self.D:
            raise ValueError(f"last dimension {x.shape[-1]} != table width {self.D}")
        if not x.is_cuda or x.dtype != torch.float32 or x.device != self._E.device:
            raise RuntimeError("CodeTable.nearest needs a CUDA float32 tensor on the table's device (no CPU path)")
        bits = _index_bits(dtype, self.K)
        dev = x.device
        rows = x.reshape(-1, self.D)
        N = rows.shape[0]
        if self.D < _D:
            xr = torch.zeros((N, _D), dtype=torch.float32, device=dev)
            xr[:, : self.D] = rows
        else:
            xr = rows.contiguous()
        idx = torch.empty((N,), dtype=dtype, device=dev)
        stats = torch.empty((4,), dtype=torch.int64, device=dev)
        nbytes = _native.workspace_bytes(N, self.K, _D)
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = _native.lib().vq_argmin_rows(xr.data_ptr(), N, _D, self._E.data_ptr(), self._E_h.data_ptr(),
                                              self._e2.data_ptr(), self._cb.data_ptr(), self.K, _native.VQ_RECIPES[recipe],
                                              idx.data_ptr(), bits,
                                              stats.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                              int(torch.cuda.current_stream(dev).cuda_stream))
        _native.check(rc, "vq_argmin_rows")
        self.last_stats = stats
        return idx.reshape(x.shape[:-1])


def nearest_indices(x: torch.Tensor, table: torch.Tensor, dtype=torch.int64, recipe: str = "expanded") -> torch.Tensor:
    """One-shot form of :class:`CodeTable` (prepares the table on every call)."""
    return CodeTable(table).nearest(x, dtype=dtype, recipe=recipe)
