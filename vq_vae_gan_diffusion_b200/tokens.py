"""Token-stream formats right after the tokeniser, same names and argument meaning as the reference.

* :func:`index_to_log_onehot` -- ``index_to_log_onehot(x, num_classes)`` of network/vq_diffusion/vq_diffusion.py:29-35 and
  network/vqDiffusion/submodule/diffusion_vq_official.py:53-60: ``(B, ...)`` int64 tokens -> ``(B, num_classes, ...)`` fp32
  ``log(one_hot.clamp(min=1e-30))``, written once by ``vq_index_to_log_onehot`` instead of the reference's int64 one-hot,
  float copy, clamp and log passes.
* :func:`log_onehot_to_index` -- ``log_onehot_to_index(log_x)`` of network/vq_diffusion/vq_diffusion.py:37-38 (``argmax(1)``).
* :func:`mask_and_replace` -- the input corruption of ``VQTransformer.forward`` (network/vqTransformer/vqTransformer.py:117-141):
  the Bernoulli keep-mask and the random replacement tokens are drawn by torch exactly as the reference draws them (same
  calls, same order, same generator), the round / cast / blend / sos concatenation run as one ``vq_mask_replace`` launch.

No CPU path: CPU tensors raise.
"""
from __future__ import annotations

import torch

from . import _native
from .codebook import _ptr, _stream_ptr

_CLAMP_MIN = 1e-30          # vq_diffusion.py:34 / diffusion_vq_official.py:59


def index_to_log_onehot(x: torch.Tensor, num_classes: int, *, validate: bool = True) -> torch.Tensor:
    """``torch.log(F.one_hot(x, num_classes).permute(0, -1, 1, ...).float().clamp(min=1e-30))``.

    ``validate`` keeps ``F.one_hot``'s error behaviour (a ``RuntimeError`` for a class value outside
    ``[0, num_classes)``) at the price of one host synchronisation -- the reference's own
    ``assert x.max().item() < num_classes`` (diffusion_vq_official.py:54) costs the same."""
    if not x.is_cuda:
        raise RuntimeError("index_to_log_onehot has no CPU path")
    if x.dtype != torch.int64:
        raise RuntimeError("one_hot is only applicable to index tensor of type LongTensor.")
    if x.dim() < 1:
        raise RuntimeError("index_to_log_onehot needs a batch dimension")
    num_classes = int(num_classes)
    if num_classes < 1:
        raise RuntimeError("num_classes must be positive")
    if validate and x.numel() > 0:
        lo, hi = (int(v) for v in torch.aminmax(x))          # one reduction, one host synchronisation
        if lo < 0:
            raise RuntimeError("Class values must be non-negative.")
        if hi >= num_classes:
            raise RuntimeError("Class values must be smaller than num_classes.")
    B = x.shape[0]
    rest = tuple(x.shape[1:])
    L = 1
    for s in rest:
        L *= s
    xc = x.contiguous()
    out = torch.empty((B, num_classes) + rest, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _native.lib().vq_index_to_log_onehot(_ptr(xc), B, L, num_classes, _CLAMP_MIN, _ptr(out), _stream_ptr(x.device))
        _native.check(rc, "vq_index_to_log_onehot")
    return out


def log_onehot_to_index(log_x: torch.Tensor) -> torch.Tensor:
    """``log_x.argmax(1)`` (network/vq_diffusion/vq_diffusion.py:37-38): ``(B, C, ...)`` fp32 -> ``(B, ...)`` int64, the first
    maximal class wins and a NaN counts as the maximum, as in ``torch.argmax``.  One pass over the input."""
    if not log_x.is_cuda:
        raise RuntimeError("log_onehot_to_index has no CPU path")
    if log_x.dtype != torch.float32 or log_x.dim() < 2:
        raise RuntimeError("log_onehot_to_index expects a float32 (B, C, ...) tensor")
    B, C = log_x.shape[0], log_x.shape[1]
    if C < 1:
        raise RuntimeError("argmax over an empty class axis")
    rest = tuple(log_x.shape[2:])
    L = 1
    for s in rest:
        L *= s
    xc = log_x.contiguous()
    out = torch.empty((B,) + rest, dtype=torch.int64, device=log_x.device)
    with torch.cuda.device(log_x.device):
        rc = _native.lib().vq_log_onehot_to_index(_ptr(xc), B, L, C, _ptr(out), _stream_ptr(log_x.device))
        _native.check(rc, "vq_log_onehot_to_index")
    return out


def mask_and_replace(indices: torch.Tensor, pkeep: float, vocab_size: int, sos_token: int) -> torch.Tensor:
    """``cat(sos, mask * indices + (1 - mask) * random_indices)`` with ``mask ~ Bernoulli(pkeep)`` and
    ``random_indices ~ randint(vocab_size)`` drawn like vqTransformer.py:121-129 (so a seeded run sees the reference's
    tokens).  ``indices``: ``(B, L)`` int64 on a B200; returns ``(B, L + 1)`` int64."""
    if not indices.is_cuda:
        raise RuntimeError("mask_and_replace has no CPU path")
    if indices.dtype != torch.int64 or indices.dim() != 2:
        raise RuntimeError("indices must be a (B, L) int64 tensor")
    mask = torch.bernoulli(pkeep * torch.ones(indices.shape, device=indices.device))        # vqTransformer.py:121-123
    random_indices = torch.randint_like(indices, high=vocab_size)                             # vqTransformer.py:127-129
    return blend_with_sos(indices, mask, random_indices, sos_token)


def blend_with_sos(indices: torch.Tensor, mask: torch.Tensor, random_indices: torch.Tensor, sos_token: int) -> torch.Tensor:
    """The deterministic part of :func:`mask_and_replace` (vqTransformer.py:117-118, 124, 138-141): ``mask`` is the fp32 Bernoulli
    draw; returns ``cat(full((B, 1), sos), mask.round().long() * indices + (1 - mask.round().long()) * random_indices)``."""
    if not (indices.is_cuda and mask.is_cuda and random_indices.is_cuda):
        raise RuntimeError("blend_with_sos has no CPU path")
    if indices.dtype != torch.int64 or random_indices.dtype != torch.int64 or mask.dtype != torch.float32:
        raise RuntimeError("indices / random_indices must be int64 and mask float32")
    if indices.dim() != 2 or mask.shape != indices.shape or random_indices.shape != indices.shape:
        raise RuntimeError("indices, mask and random_indices must share one (B, L) shape")
    B, L = indices.shape
    ic, mc, rc_ = indices.contiguous(), mask.contiguous(), random_indices.contiguous()
    out = torch.empty((B, L + 1), dtype=torch.int64, device=indices.device)
    with torch.cuda.device(indices.device):
        rc = _native.lib().vq_mask_replace(_ptr(ic), _ptr(mc), _ptr(rc_), int(sos_token), B, L, _ptr(out), _stream_ptr(indices.device))
        _native.check(rc, "vq_mask_replace")
    return out
