"""Timing of the exact full-scan fallback in its two regimes on one B200 (CUDA events around CodeBook.forward, median).

    [VQ_B200_LIB=other.so] python tools/fallback_bench.py

  mass:    every row overflows its candidate list (a codebook of 4 distinct codes, each repeated K/4 times): one CTA per
           row group scans all K codes -- throughput regime
  sparse:  the reference's init distribution at K=16384 (about 40 of 262144 rows overflow): the scan is split over the
           chip -- latency regime (the forward also contains the GEMM and the select pass; compare builds, not cases)
Not a bench.py line: it feeds DESIGN.md.
"""
from __future__ import annotations

import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import vq_vae_gan_diffusion_b200 as vq  # noqa: E402


def timed(fn, reps=7):
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
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(3)
    out = []
    for name, K, B in (("mass", 4096, 64), ("mass", 16384, 16), ("sparse", 16384, 256)):
        D, H, W = 256, 32, 32
        if name == "mass":
            E = torch.randn(4, D, device=dev, generator=g).repeat_interleave(K // 4, dim=0)
        else:
            E = (torch.rand(K, D, device=dev, generator=g) * 2 - 1) / K
        z = torch.randn(B, D, H, W, device=dev, generator=g)
        cb = vq.CodeBook(K, D).to(dev)
        with torch.no_grad():
            cb.codebook.weight.copy_(E)
            cb(z)
            torch.cuda.synchronize()
            ms = timed(lambda: cb(z))
        st = cb.stats_dict()
        out.append(dict(case=name, K=K, N=B * H * W, forward_ms=round(ms, 4), fallback_rows=st["fallback_rows"],
                        lib=os.path.basename(os.environ.get("VQ_B200_LIB", "libvq_b200.so"))))
        print(json.dumps(out[-1]), flush=True)


if __name__ == "__main__":
    main()
