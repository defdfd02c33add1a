"""Same-box GPU baseline: the reference's op sequence (codebook.py:62-111) in PyTorch eager on the B200 -- fp32 cuBLAS
sgemm + ATen elementwise / argmin / embedding kernels, TF32 off -- timed with CUDA events next to this library's
CodeBook on identical inputs ("the number a user would otherwise get").  Not a bench.py line.

    python tools/gpu_eager_baseline.py [--configs cfg2,cfg3,cfg4,cfg5] [--reps 5]
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

D = 256
BETA = 0.25


def eager_forward(z, weight, beta):
    """The reference forward, op for op (codebook.py:62-111)."""
    zp = z.permute(0, 2, 3, 1).contiguous()
    zf = zp.view(-1, D)
    d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(weight ** 2, dim=1) - 2 * torch.matmul(zf, weight.t())
    idx = torch.argmin(d, dim=1)
    z_q = torch.nn.functional.embedding(idx, weight).view(zp.shape)
    loss = torch.mean((z_q.detach() - zp) ** 2 + beta * torch.mean((z_q - zp.detach()) ** 2))
    z_q = zp + (z_q - zp).detach()
    return z_q.permute(0, 3, 1, 2), idx, loss


def timed(fn, reps):
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
    ap.add_argument("--configs", default="cfg2,cfg3,cfg4,cfg5")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    assert torch.backends.cuda.matmul.allow_tf32 is False
    dev = torch.device("cuda:0")
    for name in args.configs.split(","):
        wl = WORKLOADS[name]
        B, H, W, K = wl["B"], wl["H"], wl["W"], wl["K"]
        tok = bool(wl.get("tokenizer"))
        N = B * H * W
        E, z, g_out = make_latents(torch, dev, B, H, W, K, "trained", 1234)
        w_ref = E.clone().requires_grad_(not tok)
        cb = vq.CodeBook(K, D, BETA).to(dev)
        with torch.no_grad():
            cb.codebook.weight.copy_(E)
        one = torch.ones((), device=dev)

        def step_eager():
            if tok:
                with torch.no_grad():
                    return eager_forward(z, w_ref, BETA)[1]
            zz = z.detach().requires_grad_(True)
            w_ref.grad = None
            z_q, idx, loss = eager_forward(zz, w_ref, BETA)
            torch.autograd.backward([z_q, loss], [g_out, one])
            return idx

        def step_ours():
            cb.refresh_codebook()
            if tok:
                return cb.encode_indices(z)
            zz = z.detach().requires_grad_(True)
            cb.codebook.weight.grad = None
            z_q, idx, loss = cb(zz)
            torch.autograd.backward([z_q, loss], [g_out, one])
            return idx

        idx_e = step_eager()
        idx_o = step_ours()
        torch.cuda.synchronize()
        agree = float((idx_e == idx_o).double().mean())
        for _ in range(2):
            step_eager(); step_ours()
        t_e = timed(step_eager, args.reps)
        t_o = timed(step_ours, args.reps)
        print(json.dumps({"config": name, "mode": "tokenize" if tok else "fwd+bwd", "N": N, "K": K,
                          "eager_torch_ms": round(t_e, 4), "ours_ms": round(t_o, 4), "speedup": round(t_e / t_o, 2),
                          "eager_Mlatents_s": round(N / t_e / 1e3, 2), "ours_Mlatents_s": round(N / t_o / 1e3, 2),
                          "index_agreement_with_cublas_sgemm_argmin": agree,
                          "eager_peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}))
        del cb, w_ref
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    main()
