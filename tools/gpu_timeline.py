"""Per-tile timeline of the distance-GEMM kernel on CTA 0 (debug instrumentation, vq_debug_timeline).

Prints, for the steady state, the average cycles between the pipeline events of one code tile:
    MMA warp: wait for the TMEM buffer, issue; epilogue: wake after tcgen05.commit, release of the buffer, end of tile.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import vq_vae_gan_diffusion_b200 as vq  # noqa: E402
from vq_vae_gan_diffusion_b200 import _native  # noqa: E402
from bench import make_latents  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    dist = sys.argv[1] if len(sys.argv) > 1 else "init"
    B, H, W, K, D = 256, 32, 32, 16384, 256
    if len(sys.argv) > 3:                                  # python tools/gpu_timeline.py init <B> <K>: small shapes, raw rows
        B, K = int(sys.argv[2]), int(sys.argv[3])
    E, z, _ = make_latents(torch, dev, B, H, W, K, dist, 1234)
    cb = vq.CodeBook(K, D).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(E)
        cb.encode_indices(z)
        tiles = 512
        stamps = torch.zeros((tiles, 12), dtype=torch.int64, device=dev)
        _native.check(_native.lib().vq_debug_timeline(stamps.data_ptr(), tiles), "timeline")
        cb.encode_indices(z)
        torch.cuda.synchronize()
        _native.check(_native.lib().vq_debug_timeline(None, 0), "timeline off")
    t = stamps.cpu().numpy().astype(np.int64)
    if len(sys.argv) > 3:
        # raw view of CTA 0's first code tiles: cycles since its first MMA wait; a row tile is K / 256 consecutive tiles
        kt = (K + 255) // 256
        t0 = t[0, 0]
        print(f"B={B} K={K}: {kt} code tiles per row tile; columns: tile, mma_wait_start, mma_wait_end, mma_issued, epi0_woke, epi0_released, epi0_done, b_wait")
        for i in range(min(tiles, 5 * kt)):
            if t[i, 1] == 0:
                break
            print(i, "|" if i % kt == 0 else " ", *(int(v - t0) for v in t[i, :6]), int(t[i, 8]))
        return
    t = t[64:448]                                    # steady state (skip pipeline fill, stay inside row tiles)
    names = ["mma_wait_start", "mma_wait_end", "mma_issued", "epi0_woke", "epi0_released", "epi0_done", "epi1_released", "epi1_done"]
    per_tile = np.diff(t[:, 1]).mean()
    print(f"distribution={dist}  cycles per code tile (MMA start to MMA start): {per_tile:.0f}")
    print(f"  MMA warp waits for TMEM buffer      : {np.mean(t[:,1]-t[:,0]):7.0f}")
    print(f"  MMA issue of 16 UMMAs + commits     : {np.mean(t[:,2]-t[:,1]):7.0f}")
    print(f"    of which waiting for codebook stages (TMA): {np.mean(t[:,8]):7.0f}")
    print(f"  issue end -> epilogue g0 wakes      : {np.mean(t[:,3]-t[:,2]):7.0f}   (MMA execution tail + commit + mbarrier wake)")
    print(f"  epilogue g0 wake -> buffer released : {np.mean(t[:,4]-t[:,3]):7.0f}")
    print(f"  epilogue g1 release after g0 wake   : {np.mean(t[:,6]-t[:,3]):7.0f}")
    print(f"  release (later of g0,g1) -> MMA wait end of tile+2 : {np.mean(t[2:,1]-np.maximum(t[:-2,4], t[:-2,6])):7.0f}")
    print(f"  epilogue g0 wake -> tile done       : {np.mean(t[:,5]-t[:,3]):7.0f}")
    print(f"  epilogue g1 done after g0 wake      : {np.mean(t[:,7]-t[:,3]):7.0f}")
    print(f"  epilogue g0 done -> next wake (idle): {np.mean(t[1:,3]-t[:-1,5]):7.0f}")


if __name__ == "__main__":
    main()
