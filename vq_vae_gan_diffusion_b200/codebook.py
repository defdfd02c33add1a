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
    """Grow-only scratch buffers, one per (module, device, CUDA stream): the workspace holds the call's control words and
    candidate lists, so two streams running the same module concurrently must not share one."""

    def __init__(self):
        self.bufs = {}

    def get(self, nbytes: int, device, stream: int = 0) -> torch.Tensor:
        key = (device, stream)
        buf = self.bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            if len(self.bufs) > 8:                           # streams come and go: do not hoard scratch memory
                self.bufs.clear()
            buf = self.bufs[key] = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        return buf


def _kernel_weight(weight: torch.Tensor) -> torch.Tensor:
    """The (K, D) fp32 matrix the kernels read: they use 16-byte loads on dense rows, so a non-contiguous weight or one
    at a storage offset that is not 16-byte aligned (a slice of a flat parameter buffer) is copied first."""
    if weight.is_contiguous() and weight.data_ptr() % 16 == 0:
        return weight
    return weight.detach().clone(memory_format=torch.contiguous_format)


class _VQFunction(torch.autograd.Function):
    """forward -> vq_forward, backward -> vq_backward (include/vq_b200.h)."""

    @staticmethod
    def forward(ctx, z, weight, module, refresh):
        B, D, H, W = z.shape
        K = weight.shape[0]
        dev = z.device
        zc = z.contiguous()                                  # NCHW; the kernels read it in place
        weight = _kernel_weight(weight)
        with _on_device(dev):
            st = _stream_ptr(dev)
            E_h, e2, cb = module._derived(weight, force=refresh, stream=st)
            zq = torch.empty((B, H, W, D), dtype=torch.float32, device=dev)
            idx = torch.empty((B * H * W,), dtype=torch.int64, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            hist = torch.empty((K,), dtype=torch.int64, device=dev)
            stats = torch.empty((4,), dtype=torch.int64, device=dev)
            ws = module._workspace.get(_native.workspace_bytes_cached(B * H * W, K, D), dev, st)
            # scatter_in_forward (set by a data-parallel wrapper, dist.py): the per-code sums of (e - z) -- the codebook gradient
            # up to a scalar -- are accumulated by the forward, so that their all-reduce overlaps the rest of the step
            scat = None
            if refresh and module.scatter_in_forward and not module.deterministic:
                scat = module.scatter_alloc(K, D, dev) if module.scatter_alloc is not None else \
                    torch.empty((K, D), dtype=torch.float32, device=dev)
            rc = _native.lib().vq_forward_ex(_ptr(zc), B, H * W, D, _ptr(weight), _ptr(E_h), _ptr(e2), _ptr(cb), K,
                                             float(module.beta), _ptr(zq), _ptr(idx), _ptr(loss), _ptr(hist), _ptr(scat),
                                             _ptr(stats), _ptr(ws), ws.numel(), st)
            if rc != 0:
                _native.check(rc, "vq_forward_ex")
            if module.count_launches:
                module._launches = int(_native.lib().vq_last_launch_count())
        object.__setattr__(module, "last_histogram", hist)   # (plain tensors: skip nn.Module.__setattr__'s bookkeeping)
        object.__setattr__(module, "last_stats", stats)
        ctx.save_for_backward(zc, idx, weight)
        ctx.module = module
        ctx.shape = (B, D, H, W)
        ctx.scat = scat
        ctx.scat_ready = module.scatter_ready if scat is not None else None
        ctx.mark_non_differentiable(idx)
        ctx.set_materialize_grads(False)      # no zeros_like(idx) / zeros for unused outputs: backward handles None
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
            st = _stream_ptr(dev)
            grad_z = torch.empty((B, D, H, W), dtype=torch.float32, device=dev) if need_z else None
            grad_E = None
            scat = ctx.scat if need_w else None
            if need_w:
                # a data-parallel wrapper may hand out the head of its flat exchange buffer (dist.py): the scatter-add
                # then lands where the all-reduce reads, with no packing copy
                grad_E = module.grad_alloc(K, D, dev) if (module.grad_alloc is not None and scat is None) else \
                    torch.empty((K, D), dtype=torch.float32, device=dev)
            det = bool(module.deterministic) and need_w and scat is None
            ws = None
            if det:
                ws = module._workspace_bwd.get(_native.backward_workspace_bytes_cached(K, D), dev, st)
            L = _native.lib()
            n_launch = 0
            if scat is None:
                rc = L.vq_backward_ex(_ptr(g_zq), strides, 0.0, _ptr(g_loss_t), _ptr(zc), _ptr(idx), _ptr(weight),
                                      B, H * W, D, K, float(module.beta), B * H * W, float(module.grad_scale),
                                      1 if det else 0, 0, _ptr(grad_z), _ptr(grad_E), _ptr(ws),
                                      0 if ws is None else ws.numel(), st)
                if rc != 0:
                    _native.check(rc, "vq_backward_ex")
            else:
                # The forward accumulated the per-code sums and (data-parallel) their all-reduce is in flight: grad_z first --
                # it does not depend on them, so the collective keeps running under this kernel --, then the stream waits for
                # the summed sums and one scaling pass turns them into the codebook gradient.
                if need_z:
                    rc = L.vq_backward_ex(_ptr(g_zq), strides, 0.0, _ptr(g_loss_t), _ptr(zc), _ptr(idx), _ptr(weight),
                                          B, H * W, D, K, float(module.beta), B * H * W, 1.0, 0, 0, _ptr(grad_z), 0, 0, 0, st)
                    if rc != 0:
                        _native.check(rc, "vq_backward_ex")
                    n_launch = int(L.vq_last_launch_count()) if module.count_launches else 0
                if ctx.scat_ready is not None:
                    ctx.scat_ready(scat)
                rc = L.vq_backward_ex(0, None, 0.0, _ptr(g_loss_t), _ptr(zc), _ptr(idx), _ptr(weight),
                                      B, H * W, D, K, float(module.beta), B * H * W, float(module.grad_scale), 0, _ptr(scat),
                                      0, _ptr(grad_E), 0, 0, st)
                if rc != 0:
                    _native.check(rc, "vq_backward_ex")
            if module.count_launches:
                module._launches_bwd = n_launch + int(L.vq_last_launch_count())
        return grad_z, grad_E, None, None


class _GraphState:
    """CUDA-graph fast path for one (module, input shape, device): static buffers and the captured launch sequences.

    Small shapes (BASELINE.json configs[0] / configs[1]: 200 to 16 384 latents) are bound by launch latency and by the host
    time of the eager path (~180 us of Python / ctypes / allocator work per step against ~80 us of kernels); replaying a
    captured graph costs one launch per direction.  The graphs read and write fixed buffers, so a call copies its input in
    and -- unless ``graph_outputs == "static"`` -- clones its results out."""

    def __init__(self, module, weight, shape):
        B, D, H, W = shape
        K = weight.shape[0]
        dev = weight.device
        self.shape, self.K, self.dev, self.weight = shape, K, dev, weight
        self.generation = 0
        f32 = dict(dtype=torch.float32, device=dev)
        self.z = torch.empty((B, D, H, W), **f32)
        self.zq = torch.empty((B, H, W, D), **f32)
        self.idx = torch.empty((B * H * W,), dtype=torch.int64, device=dev)
        self.loss = torch.empty((), **f32)
        self.hist = torch.empty((K,), dtype=torch.int64, device=dev)
        self.stats = torch.empty((4,), dtype=torch.int64, device=dev)
        self.g = torch.zeros((B, H, W, D), **f32)                      # upstream gradient, channels-last like z_q
        self.g_loss = torch.ones((), **f32)
        self.grad_z = torch.empty((B, D, H, W), **f32)
        self.grad_E = torch.empty((K, D), **f32)
        k_pad = _native.padded_codes(K)
        self.E_h = torch.empty((k_pad, D), dtype=torch.float16, device=dev)
        self.e2 = torch.empty((k_pad,), **f32)
        self.cb = torch.empty((4,), **f32)
        self.ws = torch.empty(_native.workspace_bytes_cached(B * H * W, K, D), dtype=torch.uint8, device=dev)
        self.ws_bwd = torch.empty(_native.backward_workspace_bytes_cached(K, D), dtype=torch.uint8, device=dev)
        self.derived_key = None
        self.beta = float(module.beta)
        self.graphs = {}

    # ---- what the graphs contain (also used eagerly for the warm-up run before capture)
    def _prepare(self):
        K, D = self.weight.shape
        rc = _native.lib().vq_prepare_codebook(_ptr(self.weight), K, D, _ptr(self.E_h), _ptr(self.e2), _ptr(self.cb),
                                               _stream_ptr(self.dev))
        _native.check(rc, "vq_prepare_codebook")

    def _forward(self, refresh):
        B, D, H, W = self.shape
        if refresh:
            self._prepare()
        rc = _native.lib().vq_forward(_ptr(self.z), B, H * W, D, _ptr(self.weight), _ptr(self.E_h), _ptr(self.e2), _ptr(self.cb),
                                      self.K, self.beta, _ptr(self.zq), _ptr(self.idx), _ptr(self.loss), _ptr(self.hist),
                                      _ptr(self.stats), _ptr(self.ws), self.ws.numel(), _stream_ptr(self.dev))
        _native.check(rc, "vq_forward")

    def _backward(self, need_z, need_w, has_g, scale, det):
        B, D, H, W = self.shape
        strides = (ctypes.c_int64 * 3)(H * W * D, 1, D)
        rc = _native.lib().vq_backward_ex(_ptr(self.g) if has_g else 0, strides, 0.0, _ptr(self.g_loss), _ptr(self.z), _ptr(self.idx),
                                          _ptr(self.weight), B, H * W, D, self.K, self.beta, B * H * W, float(scale), 1 if det else 0, 0,
                                          _ptr(self.grad_z) if need_z else 0, _ptr(self.grad_E) if need_w else 0,
                                          _ptr(self.ws_bwd), self.ws_bwd.numel(), _stream_ptr(self.dev))
        _native.check(rc, "vq_backward_ex")

    def _graph(self, key, fn):
        g = self.graphs.get(key)
        if g is None:
            cur = torch.cuda.current_stream(self.dev)
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):                              # warm-up outside capture (module load, attributes)
                fn()
            cur.wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            self.graphs[key] = g
        return g

    def run_forward(self, z, refresh):
        self.z.copy_(z)
        if not refresh:                                                # frozen codebook: derived state follows the weight's version
            key = (self.weight.data_ptr(), self.weight._version)
            if key != self.derived_key:
                self._prepare()
                self.derived_key = key
        else:
            self.derived_key = None
        self._graph(("fwd", bool(refresh)), lambda: self._forward(refresh)).replay()
        self.generation += 1
        return self.generation

    def run_backward(self, g_zq, g_loss, need_z, need_w, scale, det):
        has_g = g_zq is not None and need_z
        if has_g:
            self.g.permute(0, 3, 1, 2).copy_(g_zq)
        if g_loss is not None:
            self.g_loss.copy_(g_loss)
        else:
            self.g_loss.zero_()
        key = ("bwd", need_z, need_w, has_g, float(scale), bool(det))
        self._graph(key, lambda: self._backward(need_z, need_w, has_g, scale, det)).replay()


class _VQGraphFunction(torch.autograd.Function):
    """The CodeBook step through captured CUDA graphs (see _GraphState)."""

    @staticmethod
    def forward(ctx, z, weight, module, refresh):
        st = module._graph_state(weight, tuple(z.shape))
        with _on_device(z.device):
            gen = st.run_forward(z, refresh)
            static = module.graph_outputs == "static"
            # "static": fresh VIEWS of the graph's buffers (never the same tensor objects twice: autograd would re-point the
            # earlier call's outputs at this call's node)
            zq = st.zq.view(st.zq.shape) if static else st.zq.clone()
            idx = st.idx.view(-1) if static else st.idx.clone()
            loss = st.loss.view(()) if static else st.loss.clone()
            hist = st.hist.view(-1) if static else st.hist.clone()
        object.__setattr__(module, "last_histogram", hist)
        object.__setattr__(module, "last_stats", st.stats)
        ctx.state, ctx.generation, ctx.module = st, gen, module
        ctx.mark_non_differentiable(idx)
        ctx.set_materialize_grads(False)
        return zq.permute(0, 3, 1, 2), idx, loss

    @staticmethod
    def backward(ctx, g_zq, _g_idx, g_loss):
        st, module = ctx.state, ctx.module
        need_z, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_z or need_w):
            return None, None, None, None
        if st.generation != ctx.generation:
            raise RuntimeError("CodeBook(use_cuda_graphs=True): the graph's static buffers were overwritten by a later forward "
                               "before this backward ran; use the eager path (use_cuda_graphs = False) for this call pattern")
        with _on_device(st.dev):
            st.run_backward(g_zq, g_loss, need_z, need_w, module.grad_scale, bool(module.deterministic) and need_w)
            static = module.graph_outputs == "static"
            grad_z = (st.grad_z.view(st.grad_z.shape) if static else st.grad_z.clone()) if need_z else None
            grad_E = st.grad_E.clone() if need_w else None               # autograd may keep it as weight.grad: never the static buffer
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
        self._derived_by_stream = {}
        self._workspace = _Workspace()
        self._workspace_bwd = _Workspace()
        self._launches = 0
        self._launches_bwd = 0
        self.count_launches = False      # bench.py: record how many kernels each call enqueued
        # deterministic=True: the codebook-gradient scatter-add runs in 64-bit fixed point (vq_backward_ex), bit-reproducible
        # from run to run; default is the faster red.global.add.v4.f32 (order of the float additions varies)
        self.deterministic = False
        # data-parallel plumbing (see dist.py).  The backward always returns the gradient of THIS rank's mean loss for z
        # (what DDP-averaged upstream layers expect); grad_scale (1 / world size) applies to the codebook gradient only,
        # so that the SUM over ranks is the gradient of the global-batch mean.  grad_alloc lets the wrapper place grad_E.
        self.grad_scale = 1.0
        self.grad_alloc = None
        # scatter_in_forward: the forward also accumulates sum (e - z) per code (vq_forward_ex), the backward turns it into the
        # codebook gradient with one scaling pass; scatter_alloc / scatter_ready let the wrapper place and exchange the sums
        self.scatter_in_forward = False
        self.scatter_alloc = None
        self.scatter_ready = None
        # use_cuda_graphs=True: forward and backward replay captured CUDA graphs on static buffers (one launch per direction
        # instead of 3-7 kernels plus ~180 us of host work): for the small, launch-bound shapes.  graph_outputs = "clone"
        # hands out copies (safe to keep); "static" hands out the graph's own buffers, valid until the next call.
        self.use_cuda_graphs = False
        self.graph_outputs = "clone"
        self._graph_states = {}
        self.last_histogram = None
        self.last_stats = None

    # ------------------------------------------------------------------ derived codebook state
    def _derived(self, weight: torch.Tensor, force: bool = False, stream: int = 0):
        """fp16 operand image + |e|^2 + scalars of the current weight.

        While the codebook is being trained (grad mode on, weight requires grad) they are rebuilt on every call -- the
        optimizer changes the weight every step anyway, and two small kernels (15 us at K = 16384) are cheaper than a
        stale operand copy.  For a frozen codebook / under ``no_grad`` (the stage-2 tokenisers) they are cached and
        refreshed whenever the weight's storage, version counter, device or shape changed (``optimizer.step()``,
        ``load_state_dict()``, ``.to(device)`` all change one of them).  In-place edits through ``weight.data`` bypass
        the version counter: call :meth:`refresh_codebook` after those."""
        key = (weight.data_ptr(), weight._version, weight.device, tuple(weight.shape))
        # one set of derived buffers per CUDA stream: they are written and read in stream order, so a second stream using
        # the module concurrently gets its own (and rebuilds it) instead of racing on a shared one
        ent = self._derived_by_stream.get(stream)
        if force or ent is None or ent[3] != key:
            K, D = weight.shape
            dev = weight.device
            k_pad = _native.padded_codes(K)
            if ent is None or ent[0].device != dev or ent[0].shape[0] != k_pad:
                if len(self._derived_by_stream) > 8:
                    self._derived_by_stream.clear()
                ent = [torch.empty((k_pad, D), dtype=torch.float16, device=dev),
                       torch.empty((k_pad,), dtype=torch.float32, device=dev),
                       torch.empty((4,), dtype=torch.float32, device=dev), None]
                self._derived_by_stream[stream] = ent
            rc = _native.lib().vq_prepare_codebook(_ptr(weight), K, D, _ptr(ent[0]), _ptr(ent[1]), _ptr(ent[2]), stream or _stream_ptr(dev))
            _native.check(rc, "vq_prepare_codebook")
            ent[3] = key
            self._E_h, self._e2, self._cb, self._derived_key = ent[0], ent[1], ent[2], key
        return ent[0], ent[1], ent[2]

    def _graph_state(self, weight, shape):
        key = (shape, weight.device, weight.data_ptr())
        st = self._graph_states.get(key)
        if st is None:
            if len(self._graph_states) >= 4:                 # a handful of static shapes, not a cache of everything ever seen
                self._graph_states.clear()
            if not (weight.is_contiguous() and weight.data_ptr() % 16 == 0):
                raise RuntimeError("use_cuda_graphs needs a contiguous, 16-byte aligned codebook weight")
            st = self._graph_states[key] = _GraphState(self, weight, shape)
        return st

    def refresh_codebook(self) -> None:
        """Drop the cached derived state (needed only after in-place edits through ``weight.data``)."""
        self._derived_key = None
        for ent in self._derived_by_stream.values():
            ent[3] = None
        for st in self._graph_states.values():
            st.derived_key = None

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
            if self.use_cuda_graphs and z.numel() > 0 and not torch.cuda.is_current_stream_capturing():
                return _VQGraphFunction.apply(z, weight, self, refresh)
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
        weight = _kernel_weight(weight)
        with _on_device(dev):
            st = _stream_ptr(dev)
            E_h, e2, cb = self._derived(weight, force=force, stream=st)
            idx = torch.empty((B * H * W,), dtype=dtype, device=dev)
            stats = torch.empty((4,), dtype=torch.int64, device=dev)
            ws = self._workspace.get(_native.workspace_bytes_cached(B * H * W, K, D), dev, st)
            if bits == 64:
                rc = _native.lib().vq_argmin(_ptr(zc), B, H * W, D, _ptr(weight), _ptr(E_h), _ptr(e2), _ptr(cb), K,
                                             _ptr(idx), _ptr(stats), _ptr(ws), ws.numel(), st)
            else:
                rc = _native.lib().vq_argmin_narrow(_ptr(zc), B, H * W, D, _ptr(weight), _ptr(E_h), _ptr(e2), _ptr(cb), K,
                                                    _ptr(idx), bits, _ptr(stats), _ptr(ws), ws.numel(), st)
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
        state["_derived_by_stream"] = {}
        state["_workspace"] = _Workspace()
        state["_workspace_bwd"] = _Workspace()
        state["grad_alloc"] = None
        state["scatter_alloc"] = None
        state["scatter_ready"] = None
        state["_graph_states"] = {}
        state["last_histogram"] = None
        state["last_stats"] = None
        return state


def vq_embed_nchw(indices: torch.Tensor, weight: torch.Tensor, B: int, H: int, W: int) -> torch.Tensor:
    """``weight[indices].reshape(B, H, W, D).permute(0, 3, 1, 2)`` materialised contiguous NCHW in one kernel
    (the decode-side lookup of worker/vqganVqvaeWorker.py:459 / vqTransformer.py:98).  ``weight`` is ``(K, D <= 256)`` fp32;
    a narrower table runs zero-padded to the kernel's 256 columns and the padding is sliced off again."""
    if not (indices.is_cuda and weight.is_cuda):
        raise RuntimeError("vq_embed_nchw has no CPU path")
    if weight.dtype != torch.float32 or weight.dim() != 2:
        raise RuntimeError(f"vq_embed_nchw expects a float32 (K, D) table, got {weight.dtype} {tuple(weight.shape)}")
    K, D = weight.shape
    if D > _KERNEL_D:
        raise ValueError(f"table width {D} > {_KERNEL_D} is not supported")
    idx = indices.reshape(-1).to(torch.int64).contiguous()
    if idx.numel() != B * H * W:
        raise ValueError("indices do not match B*H*W")
    w = weight.detach().contiguous()
    if D < _KERNEL_D:
        w = torch.nn.functional.pad(w, (0, _KERNEL_D - D))
    out = torch.empty((B, _KERNEL_D, H, W), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        rc = _native.lib().vq_embed_nchw(_ptr(idx), _ptr(w), B, H * W, _KERNEL_D, K, _ptr(out), _stream_ptr(w.device))
        _native.check(rc, "vq_embed_nchw")
    return out if D == _KERNEL_D else out[:, :D].contiguous()
