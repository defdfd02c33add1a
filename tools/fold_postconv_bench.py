"""CodeBook -> post_quant_conv as the reference composes it (vqvae.py:131-133: our CodeBook + a cuDNN 1x1 convolution) against
FoldedPostQuant (postconv.py), forward and forward + backward, on a BASELINE.json workload.  Prints one JSON line.

    python tools/fold_postconv_bench.py [--workload cfg4] [--distribution trained] [--reps 20]
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
    cb = vq.CodeBook(K, D).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(E)
    conv = torch.nn.Conv2d(D, D, 1).to(dev)
    fused = vq.FoldedPostQuant(cb, conv)
    g = g_out.contiguous()                                    # gradient on post_quant_x arrives NCHW from the decoder
    zr = z.clone().requires_grad_(True)

    def unfused_fwd():
        with torch.no_grad():
            z_q, idx, loss = cb(z)
            return conv(z_q)

    def fused_fwd():
        with torch.no_grad():
            return fused(z)[0]

    def unfused_step():
        zr.grad = None
        for p in list(cb.parameters()) + list(conv.parameters()):
            p.grad = None
        z_q, idx, loss = cb(zr)
        y = conv(z_q)
        torch.autograd.backward([y, loss], [g, torch.ones((), device=dev)])

    def fused_step():
        zr.grad = None
        for p in list(cb.parameters()) + list(conv.parameters()):
            p.grad = None
        y, idx, loss = fused(zr)
        torch.autograd.backward([y, loss], [g, torch.ones((), device=dev)])

    out = {"workload": args.workload, "distribution": args.distribution, "N": N, "K": K,
           "cudnn_allow_tf32": bool(torch.backends.cudnn.allow_tf32),
           "forward_ms": {"codebook_then_conv": timed(unfused_fwd, args.reps), "folded": timed(fused_fwd, args.reps)},
           "forward_backward_ms": {"codebook_then_conv": timed(unfused_step, args.reps), "folded": timed(fused_step, args.reps)},
           "algorithmic_bytes_saved_forward": 2 * N * D * 4,
           "note": "saved forward traffic: the z_q write and the convolution's read of it (4 D bytes each per latent); the fold adds "
                   "a (K, 256) x (256, 256) GEMM and reads the (K, 256) table from L2"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
