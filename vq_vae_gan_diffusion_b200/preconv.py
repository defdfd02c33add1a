"""quant_conv folded into the quantiser (the encoder-side half).

The reference runs ``quant_x = self.quant_conv(encoded_images)`` right before the CodeBook
(/root/reference/network/vqvae/vqvae.py:128-131), a 1x1 convolution ``Conv2d(256, 256, 1)`` (vqvae.py:83).
:class:`FoldedQuantConv` replaces ``quant_conv -> CodeBook`` by ONE call, ``vq_forward_qconv`` (include/vq_b200.h): the
operand-preparation kernel of the quantiser starts from the convolution's input ``h``, computes ``z = W h + b`` with fp32
accuracy on the tensor cores (split-precision fp16 operands, three products into one TMEM accumulator: csrc/vq_qconv.cuh) and
derives the distance GEMM's operands from the accumulator while it is on chip.  ``z`` is written once (the exact stage, z_q,
the loss and the backward need it in fp32) and never re-read by a preparation pass; there is no separate convolution launch.

Parity: ``|z - conv_fp32(h)| <= 1e-5 max|z|`` (the fp32 CPU convolution is the oracle, oracle/vq_oracle.py: quant_conv_fp32;
measured ~1e-6, the same as the distance between two fp32 GEMM libraries); everything downstream -- indices, z_q, histogram,
loss -- is computed from exactly that ``z`` as the CodeBook would, bit for bit (tests feed the returned ``z`` to the oracle).

Backward (autograd of vqvae.py:128-131): ``vq_backward`` gives the gradient on ``z`` and the codebook gradient; the
convolution's own backward is the library's (``aten.convolution_backward``, exactly what autograd runs for the unfused layer).

Shapes the fused kernel does not take (``H * W`` not a multiple of 128, channel counts other than 256, a non-1x1 convolution is
refused at construction) run the reference's composition -- the library convolution, then the CodeBook -- instead.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from . import _native
from .codebook import CodeBook, _kernel_weight, _on_device, _ptr, _stream_ptr

__all__ = ["FoldedQuantConv", "FoldedVQ"]

_W_IMG_BYTES = 2 * 4 * 256 * 64 * 2      # hi | lo operand images of the 256 x 256 weight (vq_prepare_quant_conv)


class _QconvFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, conv_w, conv_b, weight, module, refresh, refresh_conv):
        cbm = module.codebook
        B, D, H, W = h.shape
        K = weight.shape[0]
        dev = h.device
        hc = h.contiguous()
        wk = _kernel_weight(weight)
        with _on_device(dev):
            st = _stream_ptr(dev)
            E_h, e2, cbs = cbm._derived(wk, force=refresh, stream=st)
            w_img, w_sc = module._conv_images(conv_w, force=refresh_conv, stream=st)
            z = torch.empty((B, D, H, W), dtype=torch.float32, device=dev)
            zq = torch.empty((B, H, W, D), dtype=torch.float32, device=dev)
            idx = torch.empty((B * H * W,), dtype=torch.int64, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            hist = torch.empty((K,), dtype=torch.int64, device=dev)
            stats = torch.empty((4,), dtype=torch.int64, device=dev)
            ws = cbm._workspace.get(_native.workspace_bytes_cached(B * H * W, K, D), dev, st)
            bias = None if conv_b is None else conv_b.detach().contiguous()
            rc = _native.lib().vq_forward_qconv(_ptr(hc), B, H * W, D, _ptr(w_img), _ptr(w_sc), _ptr(bias), _ptr(z), _ptr(wk), _ptr(E_h),
                                                _ptr(e2), _ptr(cbs), K, float(cbm.beta), _ptr(zq), _ptr(idx), _ptr(loss), _ptr(hist),
                                                _ptr(stats), _ptr(ws), ws.numel(), st)
            _native.check(rc, "vq_forward_qconv")
            if cbm.count_launches:
                cbm._launches = int(_native.lib().vq_last_launch_count())
        object.__setattr__(cbm, "last_histogram", hist)
        object.__setattr__(cbm, "last_stats", stats)
        object.__setattr__(module, "last_z", z)              # the convolution's output (the reference's quant_x), detached
        ctx.save_for_backward(hc, z, idx, wk, conv_w)
        ctx.module = module
        ctx.shape = (B, D, H, W)
        ctx.has_bias = conv_b is not None
        ctx.mark_non_differentiable(idx)
        ctx.set_materialize_grads(False)
        return zq.permute(0, 3, 1, 2), idx, loss

    @staticmethod
    def backward(ctx, g_zq, _g_idx, g_loss):
        hc, z, idx, wk, conv_w = ctx.saved_tensors
        cbm = ctx.module.codebook
        B, D, H, W = ctx.shape
        K = wk.shape[0]
        dev = hc.device
        need_h, need_w, need_b, need_E = ctx.needs_input_grad[0], ctx.needs_input_grad[1], \
            bool(ctx.needs_input_grad[2] and ctx.has_bias), ctx.needs_input_grad[3]
        need_z = need_h or need_w or need_b
        if not (need_z or need_E):
            return None, None, None, None, None, None, None
        strides = None
        if g_zq is not None:
            if g_zq.dtype != torch.float32:
                g_zq = g_zq.float()
            sb, sd, sh, sw = g_zq.stride()
            if not (H == 1 or W == 1 or sh == W * sw):
                g_zq = g_zq.contiguous()
                sb, sd, sh, sw = g_zq.stride()
            strides = (ctypes.c_int64 * 3)(sb, sd, sw if W > 1 else (sh if H > 1 else 1))
        g_loss_t = None if g_loss is None else g_loss.to(device=dev, dtype=torch.float32).contiguous()
        with _on_device(dev):
            st = _stream_ptr(dev)
            grad_z = torch.empty((B, D, H, W), dtype=torch.float32, device=dev) if need_z else None
            grad_E = None
            if need_E:
                grad_E = cbm.grad_alloc(K, D, dev) if cbm.grad_alloc is not None else torch.empty((K, D), dtype=torch.float32, device=dev)
            det = bool(cbm.deterministic) and need_E
            ws = cbm._workspace_bwd.get(_native.backward_workspace_bytes_cached(K, D), dev, st) if det else None
            rc = _native.lib().vq_backward_ex(_ptr(g_zq), strides, 0.0, _ptr(g_loss_t), _ptr(z), _ptr(idx), _ptr(wk), B, H * W, D, K,
                                              float(cbm.beta), B * H * W, float(cbm.grad_scale), 1 if det else 0, 0, _ptr(grad_z),
                                              _ptr(grad_E), _ptr(ws), 0 if ws is None else ws.numel(), st)
            _native.check(rc, "vq_backward_ex")
        grad_h = grad_w = grad_b = None
        if need_z:
            # the convolution's own backward, through the library exactly as autograd would run it for the unfused layer
            grad_h, grad_w, grad_b = torch.ops.aten.convolution_backward(
                grad_z, hc, conv_w, [conv_w.shape[0]] if need_b else None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1,
                [bool(need_h), bool(need_w), bool(need_b)])
        return grad_h, grad_w, grad_b, grad_E, None, None, None


class FoldedQuantConv(nn.Module):
    """``quant_conv`` followed by the ``CodeBook`` (vqvae.py:128-131) as one fused call.

        fused = FoldedQuantConv(vqvae.quant_conv, vqvae.codebook)       # shares both modules' parameters
        z_q, indices, q_loss = fused(encoded_images)                    # == codebook(quant_conv(encoded_images))
        fused.last_z                                                    # the convolution's output (the reference's quant_x)
    """

    def __init__(self, quant_conv: nn.Conv2d, codebook: CodeBook):
        super().__init__()
        if not isinstance(quant_conv, nn.Conv2d) or quant_conv.kernel_size != (1, 1) or quant_conv.stride != (1, 1) \
                or quant_conv.padding not in ((0, 0), "valid") or quant_conv.groups != 1 or quant_conv.dilation != (1, 1):
            raise ValueError("FoldedQuantConv folds a plain 1x1 convolution (vqvae.py:83: nn.Conv2d(C, C, 1))")
        if quant_conv.out_channels != codebook.latent_dim:
            raise ValueError("quant_conv's output channels must equal the codebook's latent_dim")
        self.quant_conv = quant_conv
        self.codebook = codebook
        self._images = {}                 # per CUDA stream: [w_img, w_scalars, key]
        self.last_z = None

    def _conv_images(self, conv_w, force, stream):
        """hi / lo fp16 operand images of the convolution weight, rebuilt when the weight changed (or on every call while it
        is being trained, like the codebook's derived state)."""
        key = (conv_w.data_ptr(), conv_w._version, conv_w.device)
        ent = self._images.get(stream)
        if force or ent is None or ent[2] != key:
            if ent is None or ent[0].device != conv_w.device:
                if len(self._images) > 8:
                    self._images.clear()
                ent = [torch.empty(_W_IMG_BYTES, dtype=torch.uint8, device=conv_w.device),
                       torch.empty(4, dtype=torch.float32, device=conv_w.device), None]
                self._images[stream] = ent
            w2 = conv_w.detach().reshape(conv_w.shape[0], conv_w.shape[1]).contiguous()
            rc = _native.lib().vq_prepare_quant_conv(_ptr(w2), _ptr(ent[0]), _ptr(ent[1]), stream)
            _native.check(rc, "vq_prepare_quant_conv")
            ent[2] = key
        return ent[0], ent[1]

    def fusable(self, h: torch.Tensor) -> bool:
        qc = self.quant_conv
        return (h.dim() == 4 and h.is_cuda and h.dtype == torch.float32 and qc.in_channels == 256 and qc.out_channels == 256
                and self.codebook.latent_dim == 256 and h.shape[1] == 256 and h.shape[0] > 0 and (h.shape[2] * h.shape[3]) % 128 == 0
                and h.shape[2] * h.shape[3] > 0 and qc.weight.dtype == torch.float32 and qc.weight.device == h.device
                and not self.codebook.use_cuda_graphs and not torch.cuda.is_current_stream_capturing())

    def forward(self, h: torch.Tensor):
        cb = self.codebook
        if not self.fusable(h):
            z = self.quant_conv(h)                            # the reference's composition (library convolution)
            object.__setattr__(self, "last_z", z.detach())
            return cb(z)
        w = cb.codebook.weight
        if w.device != h.device or w.dtype != torch.float32:
            raise RuntimeError("codebook weight must be float32 on the input's device")
        grad_on = torch.is_grad_enabled()
        refresh = w.requires_grad and grad_on
        refresh_conv = self.quant_conv.weight.requires_grad and grad_on
        return _QconvFunction.apply(h, self.quant_conv.weight, self.quant_conv.bias, w, self, refresh, refresh_conv)


class _FoldedBothFunction(torch.autograd.Function):
    """quant_conv -> CodeBook -> post_quant_conv with BOTH convolutions folded (see :class:`FoldedVQ`)."""

    @staticmethod
    def forward(ctx, h, qw, qb, weight, pw, pb, module, refresh, refresh_conv):
        pre = module.pre
        cbm = pre.codebook
        B, D, H, W = h.shape
        K = weight.shape[0]
        dev = h.device
        hc = h.contiguous()
        wk = _kernel_weight(weight)
        with _on_device(dev):
            st = _stream_ptr(dev)
            E_h, e2, cbs = cbm._derived(wk, force=refresh, stream=st)
            w_img, w_sc = pre._conv_images(qw, force=refresh_conv, stream=st)
            z = torch.empty((B, D, H, W), dtype=torch.float32, device=dev)
            idx = torch.empty((B * H * W,), dtype=torch.int64, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            hist = torch.empty((K,), dtype=torch.int64, device=dev)
            stats = torch.empty((4,), dtype=torch.int64, device=dev)
            ws = cbm._workspace.get(_native.workspace_bytes_cached(B * H * W, K, D), dev, st)
            bias = None if qb is None else qb.detach().contiguous()
            # the quantiser without its z_q output (zq_nhwc = NULL): indices, loss, histogram -- and z, for the backward
            rc = _native.lib().vq_forward_qconv(_ptr(hc), B, H * W, D, _ptr(w_img), _ptr(w_sc), _ptr(bias), _ptr(z), _ptr(wk), _ptr(E_h),
                                                _ptr(e2), _ptr(cbs), K, float(cbm.beta), 0, _ptr(idx), _ptr(loss), _ptr(hist),
                                                _ptr(stats), _ptr(ws), ws.numel(), st)
            _native.check(rc, "vq_forward_qconv")
            # post_quant_conv of the straight-through value is one of K vectors: T = E W_p^T + b_p (postconv.py), looked up NCHW
            p2 = pw.reshape(pw.shape[0], pw.shape[1])
            table = torch.addmm(pb, wk, p2.t()) if pb is not None else wk @ p2.t()
            out = torch.empty((B, table.shape[1], H, W), dtype=torch.float32, device=dev)
            rc = _native.lib().vq_embed_nchw(_ptr(idx), _ptr(table), B, H * W, table.shape[1], K, _ptr(out), st)
            _native.check(rc, "vq_embed_nchw")
        object.__setattr__(cbm, "last_histogram", hist)
        object.__setattr__(cbm, "last_stats", stats)
        object.__setattr__(pre, "last_z", z)
        ctx.save_for_backward(hc, z, idx, wk, qw, pw)
        ctx.module = module
        ctx.shape = (B, D, H, W)
        ctx.has_qb, ctx.has_pb = qb is not None, pb is not None
        ctx.mark_non_differentiable(idx)
        ctx.set_materialize_grads(False)
        return out, idx, loss

    @staticmethod
    def backward(ctx, g_y, _g_idx, g_loss):
        hc, z, idx, wk, qw, pw = ctx.saved_tensors
        cbm = ctx.module.pre.codebook
        B, D, H, W = ctx.shape
        K = wk.shape[0]
        dev = hc.device
        nig = ctx.needs_input_grad
        need_h, need_qw, need_qb, need_E, need_pw, need_pb = nig[0], nig[1], bool(nig[2] and ctx.has_qb), nig[3], nig[4], \
            bool(nig[5] and ctx.has_pb)
        need_z = need_h or need_qw or need_qb
        g_zq = grad_pw = grad_pb = None
        strides = None
        if g_y is not None and (need_z or need_pw or need_pb):
            with _on_device(dev):
                zq_nchw = torch.empty((B, D, H, W), dtype=torch.float32, device=dev)
                rc = _native.lib().vq_embed_nchw(_ptr(idx), _ptr(wk), B, H * W, D, K, _ptr(zq_nchw), _stream_ptr(dev))
                _native.check(rc, "vq_embed_nchw")
            g_zq, grad_pw, grad_pb = torch.ops.aten.convolution_backward(
                g_y.float().contiguous(), zq_nchw, pw, [pw.shape[0]] if need_pb else None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1,
                [bool(need_z), bool(need_pw), bool(need_pb)])
            if g_zq is not None:
                g_zq = g_zq.contiguous()
                strides = (ctypes.c_int64 * 3)(D * H * W, H * W, 1)                     # NCHW
        g_loss_t = None if g_loss is None else g_loss.to(device=dev, dtype=torch.float32).contiguous()
        grad_z = grad_E = grad_h = grad_qw = grad_qb = None
        if need_z or need_E:
            with _on_device(dev):
                st = _stream_ptr(dev)
                grad_z = torch.empty((B, D, H, W), dtype=torch.float32, device=dev) if need_z else None
                if need_E:
                    grad_E = cbm.grad_alloc(K, D, dev) if cbm.grad_alloc is not None else \
                        torch.empty((K, D), dtype=torch.float32, device=dev)
                det = bool(cbm.deterministic) and need_E
                ws = cbm._workspace_bwd.get(_native.backward_workspace_bytes_cached(K, D), dev, st) if det else None
                rc = _native.lib().vq_backward_ex(_ptr(g_zq), strides, 0.0, _ptr(g_loss_t), _ptr(z), _ptr(idx), _ptr(wk), B, H * W, D, K,
                                                  float(cbm.beta), B * H * W, float(cbm.grad_scale), 1 if det else 0, 0, _ptr(grad_z),
                                                  _ptr(grad_E), _ptr(ws), 0 if ws is None else ws.numel(), st)
                _native.check(rc, "vq_backward_ex")
        if need_z:
            grad_h, grad_qw, grad_qb = torch.ops.aten.convolution_backward(
                grad_z, hc, qw, [qw.shape[0]] if need_qb else None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1,
                [bool(need_h), bool(need_qw), bool(need_qb)])
        return grad_h, grad_qw, grad_qb, grad_E, grad_pw, grad_pb, None, None, None


class FoldedVQ(nn.Module):
    """``quant_conv -> CodeBook -> post_quant_conv`` (vqvae.py:128-133) with both 1x1 convolutions folded into the quantiser:

        fused = FoldedVQ(vqvae.quant_conv, vqvae.codebook, vqvae.post_quant_conv)      # shares the three modules' parameters
        post_quant_x, indices, q_loss = fused(encoded_images)

    The encoder side runs inside the operand-preparation kernel (:class:`FoldedQuantConv`), the decoder side as a codebook-sized
    lookup (:class:`~vq_vae_gan_diffusion_b200.postconv.FoldedPostQuant`); z_q itself is never written.  Shapes the fused
    kernel does not take run the two halves separately.  Opt-in: the first element of the 3-tuple is the second convolution's
    OUTPUT (contiguous NCHW), not z_q."""

    def __init__(self, quant_conv: nn.Conv2d, codebook: CodeBook, post_quant_conv: nn.Conv2d):
        super().__init__()
        from .postconv import FoldedPostQuant
        self.pre = FoldedQuantConv(quant_conv, codebook)
        self.post = FoldedPostQuant(codebook, post_quant_conv)

    def forward(self, h: torch.Tensor):
        pre, post = self.pre, self.post
        if not pre.fusable(h):
            return post(pre.quant_conv(h))
        cb = pre.codebook
        w = cb.codebook.weight
        if w.device != h.device or w.dtype != torch.float32:
            raise RuntimeError("codebook weight must be float32 on the input's device")
        grad_on = torch.is_grad_enabled()
        return _FoldedBothFunction.apply(h, pre.quant_conv.weight, pre.quant_conv.bias, w, post.post_quant_conv.weight,
                                         post.post_quant_conv.bias, self, w.requires_grad and grad_on,
                                         pre.quant_conv.weight.requires_grad and grad_on)
