"""quant_conv -> CodeBook as the reference composes it (vqvae.py:128-131: a library 1x1 convolution, then our CodeBook) against
FoldedQuantConv (preconv.py), forward and forward + backward, on a BASELINE.json workload.  Prints one JSON line.

The unfused convolution is timed twice: with the library's default TF32 tensor-core path (what the reference gets on a GPU:
NOT fp32-accurate) and with TF32 off (the fp32 arithmetic the fold reproduces to 1e-6).

    python tools/fold_quantconv_bench.py [--workload cfg4] [--distribution trained] [--reps 20]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import vq_vae_gan_diffusion_b200 as vq  # noqa: E402
from bench import WORKLOADS, make_latents  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg4")
    ap.add_argument("--distribution", default="trained")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    wl = WORKLOADS[args.workload]
    B, H, W, K, D = wl["B"], wl["H"], wl["W"], wl["K"], 256
    N = B * H * W
    E, z, g_out = make_latents(torch, dev, B, H, W, K, args.distribution, 1234)
    torch.manual_seed(7)
    conv = torch.nn.Conv2d(D, D, 1).to(dev)
    # encoder activations whose convolution lands on the workload's latents would need W^-1; the timing only needs the shapes
    h = z.clone()
    cb = vq.CodeBook(K, D).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(E)
    fused = vq.FoldedQuantConv(conv, cb)
    hr = h.clone().requires_grad_(True)
    one = torch.ones((), device=dev)

    def zero():
        hr.grad = None
        for p in list(cb.parameters()) + list(conv.parameters()):
            p.grad = None

    def unfused_fwd():
        with torch.no_grad():
            return cb(conv(h))[0]

    def fused_fwd():
        with torch.no_grad():
            return fused(h)[0]

    def conv_only():
        with torch.no_grad():
            return conv(h)

    def unfused_step():
        zero()
        z_q, idx, loss = cb(conv(hr))
        torch.autograd.backward([z_q, loss], [g_out, one])

    def fused_step():
        zero()
        z_q, idx, loss = fused(hr)
        torch.autograd.backward([z_q, loss], [g_out, one])

    out = {"workload": args.workload, "distribution": args.distribution, "N": N, "K": K}
    old = torch.backends.cudnn.allow_tf32
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        tag = "tf32_conv" if tf32 else "fp32_conv"
        out[f"conv_only_ms_{tag}"] = timed(conv_only, args.reps)
        out[f"forward_ms_unfused_{tag}"] = timed(unfused_fwd, args.reps)
        out[f"forward_backward_ms_unfused_{tag}"] = timed(unfused_step, args.reps)
    torch.backends.cudnn.allow_tf32 = old
    # the whole sandwich quant_conv -> CodeBook -> post_quant_conv (vqvae.py:128-133): library convolutions vs FoldedVQ
    pconv = torch.nn.Conv2d(D, D, 1).to(dev)
    both = vq.FoldedVQ(conv, cb, pconv)
    g_y = g_out.contiguous()

    def zero3():
        zero()
        for p in pconv.parameters():
            p.grad = None

    def sandwich_unfused_fwd():
        with torch.no_grad():
            return pconv(cb(conv(h))[0])

    def sandwich_fused_fwd():
        with torch.no_grad():
            return both(h)[0]

    def sandwich_unfused_step():
        zero3()
        z_q, idx, loss = cb(conv(hr))
        torch.autograd.backward([pconv(z_q), loss], [g_y, one])

    def sandwich_fused_step():
        zero3()
        y, idx, loss = both(hr)
        torch.autograd.backward([y, loss], [g_y, one])

    out["sandwich_forward_ms"] = {"library_tf32_convs": timed(sandwich_unfused_fwd, args.reps), "folded": timed(sandwich_fused_fwd, args.reps)}
    out["sandwich_forward_backward_ms"] = {"library_tf32_convs": timed(sandwich_unfused_step, args.reps),
                                           "folded": timed(sandwich_fused_step, args.reps)}
    out["forward_ms_fused"] = timed(fused_fwd, args.reps)
    out["forward_backward_ms_fused"] = timed(fused_step, args.reps)
    with torch.no_grad():
        out["codebook_forward_ms"] = timed(lambda: cb(h)[0], args.reps)
    # accuracy of the three convolutions against float64 on a slice
    with torch.no_grad():
        hs = h[:2]
        z64 = torch.nn.functional.conv2d(hs.double(), conv.weight.double(), conv.bias.double())
        fused(hs.contiguous())
        zf = fused.last_z

        def rel(a):
            return float((a.double() - z64).abs().max() / z64.abs().max())

        out["conv_rel_err_fused"] = rel(zf)
        torch.backends.cudnn.allow_tf32 = True
        out["conv_rel_err_tf32"] = rel(conv(hs))
        torch.backends.cudnn.allow_tf32 = False
        out["conv_rel_err_fp32"] = rel(conv(hs))
        torch.backends.cudnn.allow_tf32 = old
    out["algorithmic_bytes_saved_forward"] = N * D * 4
    out["note"] = ("saved forward traffic: the operand preparation's read of z (4 D bytes per latent); the convolution itself moves "
                   "the same bytes either way (read h, write z)")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
