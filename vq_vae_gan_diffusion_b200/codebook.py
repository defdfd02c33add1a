"""Drop-in replacement of the reference vector quantiser.

Mirrors ``/root/reference/network/vqvae/submodule/codebook.py::CodeBook`` (codebook.py:13-111): same class
name, same constructor ``(num_codebook_vectors=1024, latent_dim=256, beta=0.25)`` (codebook.py:30-32), same
public attributes, same inner ``nn.Embedding`` called ``codebook`` (state-dict key ``codebook.weight``; callers
index it directly, e.g. worker/vqganVqvaeWorker.py:459), and the same ``(z_q, indices, loss)`` return
(codebook.py:111) with the same shapes, dtypes and strides.  All arithmetic runs in the hand-written sm_100a
kernels behind the C-ABI of ``include/vq_b200.h``; PyTorch only owns memory, streams and autograd plumbing.
There is no CPU or eager fallback: a CPU tensor, a missing library or a non-sm_100 device raises.

Stricter than the reference on purpose (documented in DESIGN.md): inputs must be CUDA fp32 of rank 4 with
``C == latent_dim <= 256``; the reference silently re-chunks rows when ``C != latent_dim`` (codebook.py:64-66).
The kernels are specialised for 256 channels (every reference config); narrower codebooks run zero-padded (exact).

Extras that do not change the 3-tuple: ``last_histogram`` (codebook usage, ``bincount(indices, K)``),
``last_stats`` (tie / re-rank / fallback row counts) and the keyword-only ``indices_only=True`` fast path used by
tokenisers (``encode_indices``).
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _native

__all__ = ["CodeBook", "vq_embed_nchw"]

_DERIVED_ATTRS = ("_E_h", "_e2", "_cb", "_derived_key")
_KERNEL_D = 256          # channel width the kernels are specialised for (latent_dim of every reference config)


def _stream_ptr(device) -> int:
    """Raw cudaStream_t of the current stream (the fast private accessor when this torch build has it)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    try:
        return int(torch._C._cuda_getCurrentRawStream(idx))
    except AttributeError:                                   # pragma: no cover
        return int(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t) -> int:
    return 0 if t is None else int(t.data_ptr())


class _on_device:
    """``torch.cuda.device(dev)`` only when ``dev`` is not already current (the context manager costs ~10 us of host
    time, which is visible on the small, launch-bound shapes)."""

    __slots__ = ("ctx",)

    def __init__(self, dev):
        self.ctx = None if dev.index is None or dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


class _Workspace:
    """Grow-only scratch buffer, one per (module, device)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        if self.buf is None or self.buf.device != device or self.buf.numel() < nbytes:
            self.buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        return self.buf


class _VQFunction(torch.autograd.Function):
    """forward -> vq_forward, backward -> vq_backward (include/vq_b200.h)."""

    @staticmethod
    def forward(ctx, z, weight, module, refresh):
        B, D, H, W = z.shape
        K = weight.shape[0]
        dev = z.device
        zc = z.contiguous()                                  # NCHW; the kernels read it in place
        with _on_device(dev):
            E_h, e2, cb = module._derived(weight, force=refresh)
            zq = torch.empty((B, H, W, D), dtype=torch.float32, device=dev)
            idx = torch.empty((B * H * W,), dtype=torch.int64, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            hist = torch.empty((K,), dtype=torch.int64, device=dev)
            stats = torch.empty((4,), dtype=torch.int64, device=dev)
            ws = module._workspace.get(_native.workspace_bytes_cached(B * H * W, K, D), dev)
            rc = _native.lib().vq_forward(_ptr(zc), B, H * W, D, _ptr(weight), _ptr(E_h), _ptr(e2), _ptr(cb), K,
                                          float(module.beta), _ptr(zq), _ptr(idx), _ptr(loss), _ptr(hist),
                                          _ptr(stats), _ptr(ws), ws.numel(), _stream_ptr(dev))
            if rc != 0:
                _native.check(rc, "vq_forward")
            if module.count_launches:
                module._launches = int(_native.lib().vq_last_launch_count())
        object.__setattr__(module, "last_histogram", hist)   # (plain tensors: skip nn.Module.__setattr__'s bookkeeping)
        object.__setattr__(module, "last_stats", stats)
        ctx.save_for_backward(zc, idx, weight)
        ctx.module = module
        ctx.shape = (B, D, H, W)
        ctx.mark_non_differentiable(idx)
        # NHWC memory exposed as NCHW: strides (H*W*D, 1, W*D, D), exactly what codebook.py:109 returns
        return zq.permute(0, 3, 1, 2), idx, loss

    @staticmethod
    def backward(ctx, g_zq, _g_idx, g_loss):
        zc, idx, weight = ctx.saved_tensors
        module = ctx.module
        B, D, H, W = ctx.shape
        K = weight.shape[0]
        dev = zc.device
        need_z, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_z or need_w):
            return None, None, None, None
        strides = None
        if g_zq is not None:
            if g_zq.dtype != torch.float32:
                g_zq = g_zq.float()
            sb, sd, sh, sw = g_zq.stride()
            if not (H == 1 or W == 1 or sh == W * sw):       # (h, w) not flattenable: take a dense copy
                g_zq = g_zq.contiguous()
                sb, sd, sh, sw = g_zq.stride()
            strides = (ctypes.c_int64 * 3)(sb, sd, sw if W > 1 else (sh if H > 1 else 1))
        g_loss_t = None
        if g_loss is not None:
            g_loss_t = g_loss.to(device=dev, dtype=torch.float32).contiguous()
        with _on_device(dev):
            grad_z = torch.empty((B, D, H, W), dtype=torch.float32, device=dev) if need_z else None
            grad_E = torch.empty((K, D), dtype=torch.float32, device=dev) if need_w else None
            n_global = B * H * W * int(module.grad_world_size)
            rc = _native.lib().vq_backward(_ptr(g_zq), strides, 0.0, _ptr(g_loss_t), _ptr(zc), _ptr(idx), _ptr(weight),
                                           B, H * W, D, K, float(module.beta), n_global, _ptr(grad_z), _ptr(grad_E),
                                           _stream_ptr(dev))
            if rc != 0:
                _native.check(rc, "vq_backward")
            if module.count_launches:
                module._launches_bwd = int(_native.lib().vq_last_launch_count())
        if grad_E is not None and module.grad_hook is not None:
            grad_E = module.grad_hook(grad_E)
        return grad_z, grad_E, None, None


class CodeBook(nn.Module):
    """Vector quantiser with the reference's interface (codebook.py:13-111), computed by sm_100a kernels.

    Args:
        num_codebook_vectors (int): number of codebook vectors K.
        latent_dim (int): dimension D of each vector (must be 256 for the CUDA path).
        beta (float): weight of the codebook term of the loss (codebook.py:96-103).
    """

    def __init__(self, num_codebook_vectors: int = 1024, latent_dim: int = 256, beta: float = 0.25):
        super().__init__()
        self.num_codebook_vectors = num_codebook_vectors
        self.latent_dim = latent_dim
        self.beta = beta

        # same RNG consumption as the reference: nn.Embedding's normal init, then uniform_ (codebook.py:40-45)
        self.codebook = nn.Embedding(num_codebook_vectors, latent_dim)
        self.codebook.weight.data.uniform_(-1 / num_codebook_vectors, 1 / num_codebook_vectors)

        # derived, non-persistent state (never in state_dict): fp16 operand copy, |e|^2, scalars
        self._E_h = None
        self._e2 = None
        self._cb = None
        self._derived_key = None
        self._workspace = _Workspace()
        self._launches = 0
        self._launches_bwd = 0
        self.count_launches = False      # bench.py: record how many kernels each call enqueued
        # data-parallel plumbing (see dist.py): loss mean runs over N_local * grad_world_size latents
        self.grad_world_size = 1
        self.grad_hook = None
        self.last_histogram = None
        self.last_stats = None

    # ------------------------------------------------------------------ derived codebook state
    def _derived(self, weight: torch.Tensor, force: bool = False):
        """fp16 operand image + |e|^2 + scalars of the current weight.

        While the codebook is being trained (grad mode on, weight requires grad) they are rebuilt on every call -- the
        optimizer changes the weight every step anyway, and two small kernels (15 us at K = 16384) are cheaper than a
        stale operand copy.  For a frozen codebook / under ``no_grad`` (the stage-2 tokenisers) they are cached and
        refreshed whenever the weight's storage, version counter, device or shape changed (``optimizer.step()``,
        ``load_state_dict()``, ``.to(device)`` all change one of them).  In-place edits through ``weight.data`` bypass
        the version counter: call :meth:`refresh_codebook` after those."""
        key = (weight.data_ptr(), weight._version, weight.device, tuple(weight.shape))
        if force or key != self._derived_key:
            K, D = weight.shape
            dev = weight.device
            k_pad = _native.padded_codes(K)
            if self._E_h is None or self._E_h.device != dev or self._E_h.shape[0] != k_pad:
                self._E_h = torch.empty((k_pad, D), dtype=torch.float16, device=dev)
                self._e2 = torch.empty((k_pad,), dtype=torch.float32, device=dev)
                self._cb = torch.empty((4,), dtype=torch.float32, device=dev)
            rc = _native.lib().vq_prepare_codebook(_ptr(weight), K, D, _ptr(self._E_h), _ptr(self._e2), _ptr(self._cb),
                                                   _stream_ptr(dev))
            _native.check(rc, "vq_prepare_codebook")
            self._derived_key = key
        return self._E_h, self._e2, self._cb

    def refresh_codebook(self) -> None:
        """Drop the cached derived state (needed only after in-place edits through ``weight.data``)."""
        self._derived_key = None

    def _check_input(self, z: torch.Tensor):
        if self.latent_dim > _KERNEL_D:
            raise ValueError(f"latent_dim {self.latent_dim} > {_KERNEL_D} is not supported by the sm_100a kernels")
        if not isinstance(z, torch.Tensor) or z.dim() != 4:
            raise ValueError(f"CodeBook expects a 4-D (B, C, H, W) tensor, got {tuple(getattr(z, 'shape', ()))}")
        if not z.is_cuda:
            raise RuntimeError("CodeBook (B200 build) has no CPU path: the input must be a CUDA tensor")
        if z.dtype != torch.float32:
            raise RuntimeError(f"CodeBook expects float32 latents, got {z.dtype}")
        if z.shape[1] != self.latent_dim:
            raise ValueError(f"channel dimension {z.shape[1]} != latent_dim {self.latent_dim}")

        w = self.codebook.weight
        if w.device != z.device:
            raise RuntimeError(f"codebook weight on {w.device}, input on {z.device}")
        if w.dtype != torch.float32:
            raise RuntimeError(f"codebook weight must be float32, got {w.dtype}")

    # ------------------------------------------------------------------ reference interface
    def forward(self, z: torch.Tensor, *, indices_only: bool = False):
        """Returns ``(z_q, min_distance_indices, loss)`` like codebook.py:47-111.

        z_q: (B, D, H, W) fp32, strides (H*W*D, 1, W*D, D); indices: (B*H*W,) int64; loss: 0-dim fp32.
        With ``indices_only=True`` only the indices are computed (``(None, indices, None)``).
        """
        self._check_input(z)
        if indices_only:
            return None, self.encode_indices(z), None
        weight = self.codebook.weight
        # (grad mode is off inside autograd.Function.forward, so "is the codebook being trained" is decided here)
        refresh = weight.requires_grad and torch.is_grad_enabled()
        if self.latent_dim == _KERNEL_D:
            return _VQFunction.apply(z, weight, self, refresh)
        # Narrower latents run zero-padded to the kernels' 256 channels.  Exact: a zero channel adds fma(0, 0, p) == p to
        # every norm and dot product, 0 to (e - z)^2 and to every gradient; only the means run over 256 / D times too
        # many elements, which the factor below undoes (a power of two for D = 32, 64, 128: bit-exact).  Costs 256 / D
        # times the memory traffic and FLOPs of a native kernel -- a compatibility path, every reference config is 256.
        pad = _KERNEL_D - self.latent_dim
        zp = torch.nn.functional.pad(z, (0, 0, 0, 0, 0, pad))
        wp = torch.nn.functional.pad(weight, (0, pad))
        zq_p, idx, loss_p = _VQFunction.apply(zp, wp, self, True)
        z_q = zq_p.permute(0, 2, 3, 1)[..., : self.latent_dim].contiguous().permute(0, 3, 1, 2)   # strides of codebook.py:109
        return z_q, idx, loss_p * (_KERNEL_D / self.latent_dim)

    @torch.no_grad()
    def encode_indices(self, z: torch.Tensor, dtype=torch.int64) -> torch.Tensor:
        """Tokeniser mode (vq_argmin): what VQTransformer/VQDiffusion.encode_to_z keep of the forward.

        ``dtype`` opts into a narrower token stream (int32; int16 for K <= 32768; uint16 for K <= 65536) -- the
        reference's own dtype is int64."""
        from .nearest import _index_bits
        self._check_input(z)
        bits = _index_bits(dtype, self.codebook.weight.shape[0])
        if self.latent_dim != _KERNEL_D:                     # zero-padded compatibility path (see forward)
            pad = _KERNEL_D - self.latent_dim
            z = torch.nn.functional.pad(z, (0, 0, 0, 0, 0, pad))
            weight_p = torch.nn.functional.pad(self.codebook.weight.detach(), (0, pad))
            return self._encode_indices_256(z, weight_p, dtype, bits, force=True)
        return self._encode_indices_256(z, self.codebook.weight, dtype, bits, force=False)

    def _encode_indices_256(self, z, weight, dtype, bits, force):
        B, D, H, W = z.shape
        K = weight.shape[0]
        dev = z.device
        zc = z.contiguous()
        with _on_device(dev):
            E_h, e2, cb = self._derived(weight, force=force)
            idx = torch.empty((B * H * W,), dtype=dtype, device=dev)
            stats = torch.empty((4,), dtype=torch.int64, device=dev)
            ws = self._workspace.get(_native.workspace_bytes_cached(B * H * W, K, D), dev)
            if bits == 64:
                rc = _native.lib().vq_argmin(_ptr(zc), B, H * W, D, _ptr(weight), _ptr(E_h), _ptr(e2), _ptr(cb), K,
                                             _ptr(idx), _ptr(stats), _ptr(ws), ws.numel(), _stream_ptr(dev))
            else:
                rc = _native.lib().vq_argmin_narrow(_ptr(zc), B, H * W, D, _ptr(weight), _ptr(E_h), _ptr(e2), _ptr(cb), K,
                                                    _ptr(idx), bits, _ptr(stats), _ptr(ws), ws.numel(), _stream_ptr(dev))
            if rc != 0:
                _native.check(rc, "vq_argmin")
            if self.count_launches:
                self._launches = int(_native.lib().vq_last_launch_count())
        object.__setattr__(self, "last_stats", stats)
        return idx

    def stats_dict(self):
        """Host copy of the last call's counters (synchronises)."""
        if self.last_stats is None:
            return None
        return dict(zip(_native.VQ_STAT_NAMES, (int(v) for v in self.last_stats.tolist())))

    # keep derived buffers out of pickles / deepcopies of the module
    def __getstate__(self):
        state = self.__dict__.copy()
        for k in _DERIVED_ATTRS:
            state[k] = None
        state["_workspace"] = _Workspace()
        state["last_histogram"] = None
        state["last_stats"] = None
        return state


def vq_embed_nchw(indices: torch.Tensor, weight: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
    """``weight[indices].reshape(B, H, W, D).permute(0, 3, 1, 2)`` materialised contiguous NCHW in one kernel
    (the decode-side lookup of worker/vqganVqvaeWorker.py:459 / vqTransformer.py:98)."""
    if not (indices.is_cuda and weight.is_cuda):
        raise RuntimeError("vq_embed_nchw has no CPU path")
    idx = indices.reshape(-1).to(torch.int64).contiguous()
    K, D = weight.shape
    if idx.numel() != B * H * W:
        raise ValueError("indices do not match B*H*W")
    w = weight.detach().contiguous()
    out = torch.empty((B, D, H, W), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        rc = _native.lib().vq_embed_nchw(_ptr(idx), _ptr(w), B, H * W, D, K, _ptr(out), _stream_ptr(w.device))
        _native.check(rc, "vq_embed_nchw")
    return out
