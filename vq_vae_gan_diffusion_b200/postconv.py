"""post_quant_conv folded into the quantiser (the decoder-side half).

The reference runs ``post_quant_x = self.post_quant_conv(codebook_mapping)`` right after the CodeBook
(/root/reference/network/vqvae/vqvae.py:131-133), a 1x1 convolution ``Conv2d(256, 256, 1)`` (vqvae.py:84) over the
straight-through value ``z_q = z + (e - z)``.  That value is the chosen code row up to one fp32 rounding, so the
convolution of every latent is one of only K different vectors::

    post_quant_conv(z_q)[n] = W_p z_q[n] + b_p  =  (E W_p^T + b_p)[idx[n]]  (1 +- 2^-23-ish)

:class:`FoldedPostQuant` therefore replaces ``CodeBook -> post_quant_conv`` by

1. the CodeBook forward WITHOUT its z_q output (``vq_forward`` with ``zq_nhwc = NULL``: indices, loss, histogram),
2. a (K, 256) x (256, 256) product ``T = E W_p^T + b_p`` per call (2 K 256^2 FLOP: 2 GFLOP at K = 16384, against the
   2 N 256^2 = 34 GFLOP of the convolution at N = 262144), a plain library GEMM in fp32,
3. the NCHW lookup ``T[idx]`` (``vq_embed_nchw``), written straight in the decoder's input layout.

Per latent that removes the z_q write (4 D bytes) and the convolution's read of it (4 D): 2 KiB of the forward's HBM
traffic, and 2 D^2 FLOP of library convolution.  The result differs from the reference's by the rounding of
``z + (e - z)`` against ``e`` (relative 2^-23 of |z| + |e| per element, amplified by at most |W_p| row sums): inside the
1e-5 bar of the north star; tests/test_gpu_parity.py compares against fp32 ``Conv2d`` on the reference's z_q.

Backward (autograd of vqvae.py:131-133): the upstream gradient on ``post_quant_x`` flows through the convolution to z_q and
from there -- straight-through -- to z only.  The convolution's backward is the library's (``aten.convolution_backward``,
i.e. cuDNN, exactly what autograd runs for the unfused layer) on the re-materialised lookup ``E[idx]``; its input gradient
goes into ``vq_backward`` as the upstream gradient.  The codebook receives its loss gradient only, as in the reference.

Accuracy note: the reference's ``z_q = fl(z + fl(e - z))`` carries a rounding error of ulp(|z|), which is NOT small against
|e| when the codebook is still at its U(-1/K, 1/K) initialisation (|e| << |z|): there the reference's own z_q is ~6e-5
(relative to |e|) away from e, and so are quantities linear in it (the convolution's weight gradient).  The fold uses e
itself -- the value the straight-through construction stands for; tests compare both against a float64 evaluation.

The encoder-side ``quant_conv`` (vqvae.py:83,128) is folded by :class:`~vq_vae_gan_diffusion_b200.preconv.FoldedQuantConv`; both sides
together: :class:`~vq_vae_gan_diffusion_b200.preconv.FoldedVQ`.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _native
from .codebook import CodeBook, _kernel_weight, _on_device, _ptr, _stream_ptr

__all__ = ["FoldedPostQuant"]


class _FoldedFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, weight, conv_w, conv_b, module, refresh):
        cbm = module.codebook
        B, D, H, W = z.shape
        K = weight.shape[0]
        dev = z.device
        zc = z.contiguous()
        wk = _kernel_weight(weight)
        with _on_device(dev):
            st = _stream_ptr(dev)
            E_h, e2, cbs = cbm._derived(wk, force=refresh, stream=st)
            idx = torch.empty((B * H * W,), dtype=torch.int64, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            hist = torch.empty((K,), dtype=torch.int64, device=dev)
            stats = torch.empty((4,), dtype=torch.int64, device=dev)
            ws = cbm._workspace.get(_native.workspace_bytes_cached(B * H * W, K, D), dev, st)
            rc = _native.lib().vq_forward(_ptr(zc), B, H * W, D, _ptr(wk), _ptr(E_h), _ptr(e2), _ptr(cbs), K, float(cbm.beta),
                                          0, _ptr(idx), _ptr(loss), _ptr(hist), _ptr(stats), _ptr(ws), ws.numel(), st)
            _native.check(rc, "vq_forward")
            # the K possible outputs of the convolution: T = E W_p^T + b_p   (fp32 library GEMM, TF32 off)
            w2 = conv_w.reshape(conv_w.shape[0], conv_w.shape[1])
            table = torch.addmm(conv_b, wk, w2.t()) if conv_b is not None else wk @ w2.t()
            out = torch.empty((B, table.shape[1], H, W), dtype=torch.float32, device=dev)
            rc = _native.lib().vq_embed_nchw(_ptr(idx), _ptr(table), B, H * W, table.shape[1], K, _ptr(out), st)
            _native.check(rc, "vq_embed_nchw")
        object.__setattr__(cbm, "last_histogram", hist)
        object.__setattr__(cbm, "last_stats", stats)
        ctx.save_for_backward(zc, idx, wk, conv_w)
        ctx.module = module
        ctx.shape = (B, D, H, W)
        ctx.has_bias = conv_b is not None
        ctx.mark_non_differentiable(idx)
        return out, idx, loss

    @staticmethod
    def backward(ctx, g_y, _g_idx, g_loss):
        zc, idx, wk, conv_w = ctx.saved_tensors
        cbm = ctx.module.codebook
        B, D, H, W = ctx.shape
        K = wk.shape[0]
        dev = zc.device
        need_z, need_E, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2], \
            bool(ctx.needs_input_grad[3] and ctx.has_bias)
        grad_w = grad_b = g_zq = None
        strides = None
        if g_y is not None and (need_z or need_w or need_b):
            # The convolution's own backward, through the library exactly as autograd would run it for the unfused layer
            # (cuDNN; honours torch.backends.cudnn.allow_tf32 like the reference's layer does).  Its input z_q is rebuilt as the
            # NCHW lookup E[idx] (the straight-through value up to one rounding) instead of having been stored by the forward.
            with _on_device(dev):
                zq_nchw = torch.empty((B, D, H, W), dtype=torch.float32, device=dev)
                rc = _native.lib().vq_embed_nchw(_ptr(idx), _ptr(wk), B, H * W, D, K, _ptr(zq_nchw), _stream_ptr(dev))
                _native.check(rc, "vq_embed_nchw")
            g_zq, grad_w, grad_b = torch.ops.aten.convolution_backward(
                g_y.float().contiguous(), zq_nchw, conv_w, [conv_w.shape[0]] if need_b else None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1,
                [bool(need_z), bool(need_w), bool(need_b)])
            if g_zq is not None:
                g_zq = g_zq.contiguous()
                strides = (ctypes.c_int64 * 3)(D * H * W, H * W, 1)                     # NCHW
        g_loss_t = None if g_loss is None else g_loss.to(device=dev, dtype=torch.float32).contiguous()
        grad_z = grad_E = None
        if need_z or need_E:
            with _on_device(dev):
                st = _stream_ptr(dev)
                grad_z = torch.empty((B, D, H, W), dtype=torch.float32, device=dev) if need_z else None
                grad_E = torch.empty((K, D), dtype=torch.float32, device=dev) if need_E else None
                det = bool(cbm.deterministic) and need_E
                ws = cbm._workspace_bwd.get(_native.backward_workspace_bytes_cached(K, D), dev, st) if det else None
                rc = _native.lib().vq_backward_ex(_ptr(g_zq), strides, 0.0, _ptr(g_loss_t), _ptr(zc), _ptr(idx), _ptr(wk), B, H * W, D, K,
                                                  float(cbm.beta), B * H * W, float(cbm.grad_scale), 1 if det else 0, 0, _ptr(grad_z),
                                                  _ptr(grad_E), _ptr(ws), 0 if ws is None else ws.numel(), st)
                _native.check(rc, "vq_backward_ex")
        return grad_z, grad_E, grad_w, grad_b, None, None


class FoldedPostQuant(nn.Module):
    """``CodeBook`` followed by ``post_quant_conv`` (vqvae.py:131-133), the convolution folded into a codebook-sized lookup.

        fused = FoldedPostQuant(vqvae.codebook, vqvae.post_quant_conv)      # shares both modules' parameters
        post_quant_x, indices, q_loss = fused(quant_x)                      # == post_quant_conv(codebook(quant_x)[0]), ...

    Opt-in: the 3-tuple's first element is the convolution's OUTPUT (contiguous NCHW), not z_q.
    """

    def __init__(self, codebook: CodeBook, post_quant_conv: nn.Conv2d):
        super().__init__()
        if not isinstance(post_quant_conv, nn.Conv2d) or post_quant_conv.kernel_size != (1, 1) or post_quant_conv.stride != (1, 1) \
                or post_quant_conv.padding not in ((0, 0), "valid") or post_quant_conv.groups != 1 or post_quant_conv.dilation != (1, 1):
            raise ValueError("FoldedPostQuant folds a plain 1x1 convolution (vqvae.py:84: nn.Conv2d(C, C, 1))")
        if post_quant_conv.in_channels != codebook.latent_dim or codebook.latent_dim != 256 or post_quant_conv.out_channels != 256:
            raise ValueError("the sm_100a kernels behind the fold are specialised for 256 channels on both sides")
        self.codebook = codebook
        self.post_quant_conv = post_quant_conv

    def forward(self, z: torch.Tensor):
        cb = self.codebook
        cb._check_input(z)
        weight = cb.codebook.weight
        refresh = weight.requires_grad and torch.is_grad_enabled()
        return _FoldedFunction.apply(z, weight, self.post_quant_conv.weight, self.post_quant_conv.bias, self, refresh)
