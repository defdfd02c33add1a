"""Randomised parity sweep on one B200: random shapes (vector and generic tile paths, odd numbers of GEMM row tiles, K
across code-tile boundaries), random value distributions (well separated, near ties, duplicated codes, scaled-up /
scaled-down magnitudes), forward + backward + tokeniser + row-major search, every result checked against the CPU oracle
(indices / histogram / z_q bit-exact, loss 1e-6, gradients 1e-5).

    python tools/fuzz_parity.py [--cases 40] [--seed 0] [--kinds trained,init,dup,big,tiny,cluster]

"cluster" plants a run of identical / nearly identical codes and a few latents next to it, so that those rows overflow
their candidate lists and take the exact full-scan fallback (split over code blocks, ragged row groups, K off the pass
boundaries).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import vq_vae_gan_diffusion_b200 as vq  # noqa: E402
from oracle.vq_oracle import COracle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--kinds", default="trained,init,dup,big,tiny,cluster")
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    dev = torch.device("cuda:0")
    orc = COracle()
    worst = dict(loss=0.0, gz=0.0, gE=0.0)
    for c in range(args.cases):
        vec = rng.random() < 0.6
        if vec:
            H, W = [(4, 8), (8, 8), (16, 16), (32, 32), (2, 48)][rng.integers(5)]
        else:
            H, W = int(rng.integers(1, 12)), int(rng.integers(1, 12))
        B = int(rng.integers(1, max(2, 6000 // (H * W))))
        K = int(rng.choice([1, 3, 31, 256, 257, 500, 1024, 1025, 3000, 4097]))
        kind = rng.choice(args.kinds.split(","))
        N = B * H * W
        if kind == "init":
            E = rng.uniform(-1.0 / K, 1.0 / K, size=(K, 256)).astype(np.float32)
            zf = rng.standard_normal((N, 256)).astype(np.float32)
        else:
            E = rng.standard_normal((K, 256)).astype(np.float32)
            if kind == "dup" and K >= 3:
                E[K // 2:] = E[: K - K // 2]
            zf = E[rng.integers(0, K, N)] + np.float32(rng.choice([0.1, 0.5, 1.5])) * rng.standard_normal((N, 256)).astype(np.float32)
            if kind == "cluster" and K >= 200:
                csize = int(rng.integers(40, min(K // 2, 400)))
                a = int(rng.integers(0, K - csize))
                v = rng.standard_normal(256).astype(np.float32)
                E[a:a + csize] = v + np.float32(rng.choice([0.0, 1e-4])) * rng.standard_normal((csize, 256)).astype(np.float32)
                hot = rng.choice(N, size=min(N, int(rng.integers(1, 70))), replace=False)
                zf[hot] = v + np.float32(0.05) * rng.standard_normal((len(hot), 256)).astype(np.float32)
            if kind == "big":
                E *= np.float32(300.0); zf *= np.float32(300.0)
            if kind == "tiny":
                E *= np.float32(1e-4); zf *= np.float32(1e-4)
        z = np.ascontiguousarray(zf.reshape(B, H, W, 256).transpose(0, 3, 1, 2))
        g = rng.standard_normal((B, H, W, 256)).astype(np.float32)
        cb = vq.CodeBook(K, 256).to(dev)
        with torch.no_grad():
            cb.codebook.weight.copy_(torch.from_numpy(E))
        zt = torch.from_numpy(z).to(dev).requires_grad_(True)
        z_q, idx, loss = cb(zt)
        gt = torch.from_numpy(g).to(dev).permute(0, 3, 1, 2)
        if rng.random() < 0.5:
            gt = gt.contiguous()
        torch.autograd.backward([z_q, loss], [gt, torch.ones((), device=dev)])
        ref = orc.forward(z, E)
        tag = f"case {c}: B={B} H={H} W={W} K={K} {kind}"
        assert np.array_equal(idx.cpu().numpy(), ref["idx"]), tag
        assert np.array_equal(cb.last_histogram.cpu().numpy(), ref["hist"]), tag
        assert np.array_equal(z_q.detach().permute(0, 2, 3, 1).reshape(-1, 256).cpu().numpy(), ref["zq_nhwc"]), tag
        st = cb.stats_dict()
        assert st["tie_rows"] == ref["tie_rows"], (tag, st, ref["tie_rows"])
        el = abs(float(loss.detach()) - float(ref["loss"])) / max(abs(float(ref["loss"])), 1e-30)
        gz, gE = orc.backward(np.transpose(g, (0, 3, 1, 2)), 1.0, z, ref["idx"], E)
        egz = float(np.abs(zt.grad.cpu().numpy() - gz).max() / max(np.abs(gz).max(), 1e-30))
        egE = float(np.abs(cb.codebook.weight.grad.cpu().numpy() - gE).max() / max(np.abs(gE).max(), 1e-30))
        assert el <= 1e-6 and egz <= 1e-5 and egE <= 1e-5, (tag, el, egz, egE)
        worst = dict(loss=max(worst["loss"], el), gz=max(worst["gz"], egz), gE=max(worst["gE"], egE))
        with torch.no_grad():
            assert np.array_equal(cb.encode_indices(zt.detach(), dtype=torch.int32).cpu().numpy().astype(np.int64), ref["idx"]), tag
            rows = torch.from_numpy(zf).to(dev)
            tab = vq.CodeTable(cb.codebook.weight.detach())
            assert np.array_equal(tab.nearest(rows).cpu().numpy(), ref["idx"]), tag
            assert np.array_equal(tab.nearest(rows, recipe="diffsq").cpu().numpy(), orc.nearest_diffsq(zf, E)["idx"]), tag
        print(f"{tag}: ok  ties={ref['tie_rows']} fallback={st['fallback_rows']}")
    print("FUZZ OK", worst)


if __name__ == "__main__":
    main()
