"""Nearest-code search over row-major vectors on the same sm_100a kernels as the CodeBook.

Mirrors the reference's three other nearest-neighbour sites, each with its own fp32 formula ("recipe") whose argmin --
ties, rounding and all -- is reproduced:

* ``"expanded"``  ``GaussianDiffusion2D.gaussian_to_indices``
  (/root/reference/network/vqDiffusion/submodule/diffusion_gaussian2d.py:322-347): ``|x|^2 + |e|^2 - 2 x.e``, the CodeBook's
  own formula (codebook.py:70-82);
* ``"diffsq"``  the search at the end of ``V_VQDiffusion.sample``
  (network/continous_vq_diffusion/v_vq_diffusion.py:114-123): ``sum((x - e) ** 2)`` over a broadcast difference;
* ``"cdist_normalized"``  ``VQGaussianDiffusion3DWrapper.gaussian_to_indices``
  (network/vqDiffusion/submodule/diffusion_gaussian3d.py:543-570): both sides ``F.normalize``-d, ``torch.cdist``, argmin.

All run the tcgen05 distance GEMM for candidates and an exact fp32 stage for the decision, through ``vq_argmin_rows`` of
include/vq_b200.h, at the table's own width (any ``D <= 512``: ``gaussian_dim`` is 96 in configs/*.yml and 512 in the 3D
wrapper's runs) -- the contraction is padded to a multiple of 64 inside the kernels, nothing is padded in memory.  There is
no CPU path.
"""
from __future__ import annotations

import torch

from . import _native

__all__ = ["CodeTable", "nearest_indices", "gaussian_to_indices"]

_MAX_D = 512
_IDX = {torch.int64: 64, torch.int32: 32, torch.int16: 16, torch.uint16: 16}


def _index_bits(dtype, K: int) -> int:
    if dtype not in _IDX:
        raise ValueError(f"index dtype must be int64, int32, int16 or uint16, got {dtype}")
    if dtype == torch.int16 and K > 32768:
        raise ValueError(f"int16 indices need K <= 32768 (got K={K}); use torch.uint16 or int32")
    if dtype == torch.uint16 and K > 65536:
        raise ValueError(f"uint16 indices need K <= 65536 (got K={K})")
    return _IDX[dtype]


def _stream(dev) -> int:
    return int(torch.cuda.current_stream(dev).cuda_stream)


class CodeTable:
    """A lookup table ``(K, D <= 512)`` prepared for nearest-row queries (fp16 operand image, |e|^2, scalars).

    Build it once per table (``gaussian_lookup_table`` is a fixed buffer, diffusion_gaussian2d.py:287 /
    diffusion_gaussian3d.py:513-515) and call :meth:`nearest` per batch; ``refresh()`` after the table changed.  The derived
    state of a recipe is built on first use: ``"cdist_normalized"`` keeps an L2-normalised fp32 copy of the table
    (the reference re-normalises its buffer on every call, diffusion_gaussian3d.py:560).
    """

    def __init__(self, table: torch.Tensor):
        if table.dim() != 2 or not (1 <= table.shape[1] <= _MAX_D):
            raise ValueError(f"table must be (K, 1 <= D <= {_MAX_D}), got {tuple(table.shape)}")
        if not table.is_cuda or table.dtype != torch.float32:
            raise RuntimeError("CodeTable needs a CUDA float32 table (no CPU path)")
        self.table = table
        self.K, self.D = table.shape
        self._ws = None
        self.last_stats = None
        self._prepared = {}
        self.refresh()

    def refresh(self):
        self._prepared = {}
        self._prepare(False)

    def _prepare(self, normalized: bool):
        st = self._prepared.get(normalized)
        if st is not None:
            return st
        t = self.table.detach().contiguous()
        dev = t.device
        lib = _native.lib()
        with torch.cuda.device(dev):
            if normalized:
                E = torch.empty_like(t)
                _native.check(lib.vq_normalize_rows(t.data_ptr(), self.K, self.D, E.data_ptr(), _stream(dev)), "vq_normalize_rows")
            else:
                E = t
            k_pad = _native.padded_codes(self.K)
            d_pad = 64 if self.D <= 64 else 128 if self.D <= 128 else 256 if self.D <= 256 else 512
            E_h = torch.empty((k_pad, d_pad), dtype=torch.float16, device=dev)
            e2 = torch.empty((k_pad,), dtype=torch.float32, device=dev)
            cb = torch.empty((4,), dtype=torch.float32, device=dev)
            rc = lib.vq_prepare_codebook(E.data_ptr(), self.K, self.D, E_h.data_ptr(), e2.data_ptr(), cb.data_ptr(), _stream(dev))
            _native.check(rc, "vq_prepare_codebook")
        st = self._prepared[normalized] = (E, E_h, e2, cb)
        return st

    @torch.no_grad()
    def nearest(self, x: torch.Tensor, dtype=torch.int64, recipe: str = "expanded") -> torch.Tensor:
        """``argmin_k distance(x, table[k])`` for every vector along the last axis of ``x`` -> indices of shape x.shape[:-1];
        ``recipe`` names the reference formula (module docstring)."""
        if recipe not in _native.VQ_RECIPES:
            raise ValueError(f"recipe must be one of {sorted(_native.VQ_RECIPES)}, got {recipe!r}")
        if x.shape[-1] != self.D:
            raise ValueError(f"last dimension {x.shape[-1]} != table width {self.D}")
        E, E_h, e2, cb = self._prepare(recipe == "cdist_normalized")
        if not x.is_cuda or x.dtype != torch.float32 or x.device != E.device:
            raise RuntimeError("CodeTable.nearest needs a CUDA float32 tensor on the table's device (no CPU path)")
        bits = _index_bits(dtype, self.K)
        dev = x.device
        rows = x.reshape(-1, self.D).contiguous()
        N = rows.shape[0]
        idx = torch.empty((N,), dtype=dtype, device=dev)
        stats = torch.empty((4,), dtype=torch.int64, device=dev)
        nbytes = _native.workspace_bytes_cached(N, self.K, self.D)
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = _native.lib().vq_argmin_rows(rows.data_ptr(), N, self.D, E.data_ptr(), E_h.data_ptr(), e2.data_ptr(), cb.data_ptr(),
                                              self.K, _native.VQ_RECIPES[recipe], idx.data_ptr(), bits, stats.data_ptr(),
                                              self._ws.data_ptr(), self._ws.numel(), _stream(dev))
        _native.check(rc, "vq_argmin_rows")
        self.last_stats = stats
        return idx.reshape(x.shape[:-1])


def nearest_indices(x: torch.Tensor, table: torch.Tensor, dtype=torch.int64, recipe: str = "expanded") -> torch.Tensor:
    """One-shot form of :class:`CodeTable` (prepares the table on every call)."""
    return CodeTable(table).nearest(x, dtype=dtype, recipe=recipe)


def gaussian_to_indices(gaussian: torch.Tensor, table: CodeTable) -> torch.Tensor:
    """``VQGaussianDiffusion3DWrapper.gaussian_to_indices`` (diffusion_gaussian3d.py:543-570): ``gaussian`` is
    ``(B, L, D)`` or ``(B, 1, L, D)`` (squeezed like :547-548); returns ``(B, L)`` int64 indices of the nearest table rows
    after L2-normalising both sides, by Euclidean distance (``torch.cdist``)."""
    if gaussian.dim() == 4:
        gaussian = gaussian.squeeze(1)
    if gaussian.dim() != 3:
        raise ValueError(f"expected (B, L, D) or (B, 1, L, D), got {tuple(gaussian.shape)}")
    return table.nearest(gaussian, recipe="cdist_normalized")
