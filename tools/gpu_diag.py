"""GPU diagnostic for the tcgen05 distance GEMM (run on the B200 box through gpurun).

Dumps the approximate scores with vq_debug_scores and compares them with a float64 evaluation of exactly the
fp16-rounded, power-of-two-scaled operands the kernel consumes.  Prints an error map that localises descriptor /
swizzle / pipeline mistakes, then checks the rigorous error bound the candidate margin relies on.
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
from vq_vae_gan_diffusion_b200 import _native  # noqa: E402


def debug_scores(z: torch.Tensor, E: torch.Tensor):
    B, D, H, W = z.shape
    K = E.shape[0]
    dev = z.device
    L = _native.lib()
    k_pad = _native.padded_codes(K)
    E_h = torch.empty((k_pad, D), dtype=torch.float16, device=dev)
    e2 = torch.empty((k_pad,), dtype=torch.float32, device=dev)
    cb = torch.empty((4,), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _native.check(L.vq_prepare_codebook(E.data_ptr(), K, D, E_h.data_ptr(), e2.data_ptr(), cb.data_ptr(), st), "prep")
    N = B * H * W
    ws = torch.empty(_native.workspace_bytes(N, K, D), dtype=torch.uint8, device=dev)
    scores = torch.full((N, k_pad), float("nan"), dtype=torch.float32, device=dev)
    _native.check(L.vq_debug_scores(z.data_ptr(), B, H * W, D, E_h.data_ptr(), e2.data_ptr(), cb.data_ptr(), K,
                                    scores.data_ptr(), ws.data_ptr(), ws.numel(), st), "debug_scores")
    torch.cuda.synchronize()
    return scores, untile_operand(E_h, 256), e2, cb


def untile_operand(img: torch.Tensor, rows: int) -> torch.Tensor:
    """Operand image [tile][D chunk][rows][8 pieces ^ (row & 7)][8 halves] -> row-major (n_rows, 256)."""
    n = img.numel() // 256
    x = img.reshape(n // rows, 4, rows, 8, 8)
    r = torch.arange(rows, device=img.device)
    j = torch.arange(8, device=img.device)
    phys = (j[None, :] ^ (r[:, None] & 7))                                  # physical piece holding logical piece j
    idx = phys[None, None, :, :, None].expand(x.shape[0], 4, rows, 8, 8)
    return torch.gather(x, 3, idx).permute(0, 2, 1, 3, 4).reshape(n, 256)


def expected_scores(z: torch.Tensor, E: torch.Tensor, E_h, e2, cb):
    """float64 model of the kernel: fp16(z * 2^a_n) . fp16(E * 2^b), rescaled, subtracted from e2."""
    B, D, H, W = z.shape
    zf = z.permute(0, 2, 3, 1).reshape(-1, D)
    mx = zf.abs().amax(dim=1)
    ex = torch.frexp(mx)[1].clamp(-100, 100)
    ex = torch.where(mx > 0, ex, torch.zeros_like(ex))
    scale = torch.pow(torch.tensor(2.0, device=z.device, dtype=torch.float64), (15 - ex).double())
    z_h = (zf.double() * scale[:, None]).float().half()
    dot = z_h.double() @ E_h.double().t()
    inv = (1.0 / scale)[:, None] * float(cb[2].item())
    return e2.double()[None, :] - 2.0 * dot * inv


def main():
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    print("device:", torch.cuda.get_device_name(0), "capability", torch.cuda.get_device_capability(0))
    ok_all = True
    for (B, H, W, K, dist) in [(1, 8, 16, 256, "trained"), (3, 5, 7, 300, "init"), (2, 16, 32, 1024, "trained"),
                               (8, 32, 32, 4096, "init")]:
        D = 256
        if dist == "init":
            E = (torch.rand(K, D, device=dev) * 2 - 1) / K
            z = torch.randn(B, D, H, W, device=dev)
        else:
            E = torch.randn(K, D, device=dev)
            z = (E[torch.randint(0, K, (B * H * W,), device=dev)] + 0.3 * torch.randn(B * H * W, D, device=dev))
            z = z.reshape(B, H, W, D).permute(0, 3, 1, 2).contiguous()
        scores, E_h, e2, cb = debug_scores(z, E)
        exp = expected_scores(z, E, E_h, e2, cb)
        got = scores[:, :K].double()
        err = (got - exp[:, :K]).abs()
        zf = z.permute(0, 2, 3, 1).reshape(-1, D)
        a = zf.double().norm(dim=1) * E.double().norm(dim=1).max()
        rel = (err / a[:, None]).max().item()
        nan = torch.isnan(got).sum().item()
        print(f"[{B}x{H}x{W} K={K} {dist}] max|err|={err.max().item():.3e}  max err/a={rel:.3e} "
              f"(accumulation budget 2^-13={2**-13:.3e})  nan={nan}  pad cols inf={torch.isinf(scores[:, K:]).all().item()}")
        # operand copy check
        exE = torch.frexp(E.abs().max())[1].item()
        E_h_exp = (E.double() * 2.0 ** (15 - exE)).float().half()
        print("   E_h exact:", torch.equal(E_h[:K], E_h_exp), " e2 rel err vs fp64:",
              ((e2[:K].double() - (E.double() ** 2).sum(1)).abs() / (E.double() ** 2).sum(1)).max().item())
        if not (rel < 2 ** -13) or nan:
            ok_all = False
            # error map: by row%128 block of 8 and by 32-column chunk
            N = got.shape[0]
            bad = (err / a[:, None]) > 2 ** -13
            print("   bad fraction:", bad.double().mean().item())
            rows_bad = bad.any(dim=1).nonzero().flatten()[:20].tolist()
            cols_bad = bad.any(dim=0).nonzero().flatten()[:40].tolist()
            print("   first bad rows:", rows_bad)
            print("   first bad cols:", cols_bad)
            r0 = rows_bad[0] if rows_bad else 0
            print("   row", r0, "got[:8]", got[r0, :8].tolist(), "exp[:8]", exp[r0, :8].tolist())
        # the bound the margin relies on: |score + z2 - d_oracle| <= eps  (d_oracle ~ fp64 distance up to fp32 rounding)
        z2 = (zf.double() ** 2).sum(1)
        d64 = z2[:, None] + (E.double() ** 2).sum(1)[None, :] - 2 * zf.double() @ E.double().t()
        e2max = (E.double() ** 2).sum(1).max()
        r = z2 + e2max + 2 * a
        eps = a * (2 ** -9 + 2 ** -12) + (r + e2max + 2 * a) * 2 ** -23       # vq_common.cuh: candidate_margin
        viol = ((got + z2[:, None] - d64).abs() > eps[:, None]).sum().item()
        worst = ((got + z2[:, None] - d64).abs() / eps[:, None]).max().item()
        print(f"   margin bound: violations={viol}, worst |err|/eps={worst:.3f}")
        if viol:
            ok_all = False
    print("DIAG", "OK" if ok_all else "FAILED")
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
