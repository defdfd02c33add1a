"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / synccheck / initcheck), one tool per run:

    compute-sanitizer --tool memcheck python tools/sanitize_cases.py

Covers every kernel and both tile paths: vector (HW % 32 == 0) and generic shapes, K not a multiple of the code tile,
the tokeniser mode, the exact fallback (split scan and whole-codebook scan), every upstream-gradient layout of the
backward, and the NCHW embedding lookup.  Results are checked against the CPU oracle so a sanitizer-clean run is also
a correct one.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import vq_vae_gan_diffusion_b200 as vq  # noqa: E402
from oracle.vq_oracle import COracle  # noqa: E402


def run(name, B, H, W, K, E=None, z=None, seed=0):
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(seed)
    if E is None:
        E = rng.standard_normal((K, 256)).astype(np.float32)
    if z is None:
        z = rng.standard_normal((B, 256, H, W)).astype(np.float32)
    g = rng.standard_normal((B, H, W, 256)).astype(np.float32)
    cb = vq.CodeBook(K, 256).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(torch.from_numpy(E))
    orc = COracle()
    ref = orc.forward(z, E)
    for layout in ("nhwc", "nchw"):
        zt = torch.from_numpy(z).to(dev).requires_grad_(True)
        cb.zero_grad(set_to_none=True)
        z_q, idx, loss = cb(zt)
        gt = torch.from_numpy(g).to(dev).permute(0, 3, 1, 2)
        if layout == "nchw":
            gt = gt.contiguous()
        torch.autograd.backward([z_q, loss], [gt, torch.ones((), device=dev)])
        torch.cuda.synchronize()
        assert np.array_equal(idx.cpu().numpy(), ref["idx"]), name
        gz, gE = orc.backward(np.transpose(g, (0, 3, 1, 2)), 1.0, z, ref["idx"], E)
        assert np.abs(zt.grad.cpu().numpy() - gz).max() <= 1e-5 * np.abs(gz).max(), name
        assert np.abs(cb.codebook.weight.grad.cpu().numpy() - gE).max() <= 1e-5 * max(np.abs(gE).max(), 1e-30), name
    with torch.no_grad():
        idx_tok = cb.encode_indices(torch.from_numpy(z).to(dev))
        out = vq.vq_embed_nchw(idx_tok, cb.codebook.weight, B, H, W)
    torch.cuda.synchronize()
    assert np.array_equal(idx_tok.cpu().numpy(), ref["idx"]), name
    assert out.shape == (B, 256, H, W)
    print(f"{name}: ok  stats={cb.stats_dict()}")


def main():
    rng = np.random.default_rng(5)
    run("vector tiles, K=300", 2, 8, 8, 300)
    run("generic tiles, K=257", 3, 5, 7, 257, seed=1)
    run("single latent, K=1", 1, 1, 1, 1, seed=2)
    # init-like distribution: many near ties -> exact stage with several candidates per row
    K = 1024
    run("init distribution", 1, 16, 16, K, E=rng.uniform(-1 / K, 1 / K, (K, 256)).astype(np.float32), seed=3)
    # identical codes: every row overflows its candidate list -> whole-codebook fallback scan
    base = rng.standard_normal((3, 256)).astype(np.float32)
    run("degenerate codebook", 1, 8, 8, 300, E=np.repeat(base, 100, axis=0), seed=4)
    # a few rows next to a cluster of near-identical codes -> split fallback scan
    K = 2048
    E = rng.standard_normal((K, 256)).astype(np.float32)
    v = rng.standard_normal(256).astype(np.float32)
    E[100:200] = v + 1e-4 * rng.standard_normal((100, 256)).astype(np.float32)
    zf = E[rng.integers(1000, K, size=256)] + 0.3 * rng.standard_normal((256, 256)).astype(np.float32)
    zf[rng.choice(256, 9, replace=False)] = v + 0.05 * rng.standard_normal((9, 256)).astype(np.float32)
    z = np.ascontiguousarray(zf.reshape(1, 16, 16, 256).transpose(0, 3, 1, 2))
    run("partial fallback", 1, 16, 16, K, E=E, z=z, seed=6)
    print("SANITIZE CASES OK")


if __name__ == "__main__":
    main()
