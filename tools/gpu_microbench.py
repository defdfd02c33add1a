"""Per-stage timing of the C-ABI entry points on one B200 (CUDA events, L2-flushed, median of reps).

    python tools/gpu_microbench.py [--configs cfg2,cfg3,cfg4,cfg5] [--dist init,trained] [--reps 10]

Prints one JSON object per (config, distribution) with the milliseconds of
    prepare_codebook | argmin (tokeniser mode) | forward | backward (full) | backward without grad_E |
    backward with NCHW-contiguous g_out
and the derived throughputs / roofline fractions of SURVEY.md 8(d).  Not a bench.py line: it feeds DESIGN.md.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import vq_vae_gan_diffusion_b200 as vq  # noqa: E402
from vq_vae_gan_diffusion_b200 import _native  # noqa: E402
from bench import WORKLOADS, load_peaks, make_latents  # noqa: E402

D = 256


def timed(fn, reps, flush):
    ts = []
    for _ in range(reps):
        flush.zero_()                       # 512 MB write: evicts the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts), min(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="cfg1,cfg2,cfg3,cfg4,cfg5")
    ap.add_argument("--dist", default="init,trained")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--rows", action="store_true", help="also time the row-major nearest-code search")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    L = _native.lib()
    peaks = load_peaks()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for name in args.configs.split(","):
        wl = WORKLOADS[name]
        B, H, W, K = wl["B"], wl["H"], wl["W"], wl["K"]
        N = B * H * W
        for dist in args.dist.split(","):
            E, z, g_out = make_latents(torch, dev, B, H, W, K, dist, 1234)
            g_nchw = g_out.contiguous()
            k_pad = _native.padded_codes(K)
            E_h = torch.empty((k_pad, D), dtype=torch.float16, device=dev)
            e2 = torch.empty((k_pad,), dtype=torch.float32, device=dev)
            cb = torch.empty((4,), dtype=torch.float32, device=dev)
            ws = torch.empty(_native.workspace_bytes(N, K, D), dtype=torch.uint8, device=dev)
            zq = torch.empty((N, D), dtype=torch.float32, device=dev)
            idx = torch.empty((N,), dtype=torch.int64, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            hist = torch.empty((K,), dtype=torch.int64, device=dev)
            stats = torch.zeros((4,), dtype=torch.int64, device=dev)
            grad_z = torch.empty_like(z)
            grad_E = torch.empty_like(E)
            s_cl = (ctypes.c_int64 * 3)(g_out.stride(0), g_out.stride(1), g_out.stride(3))
            s_nc = (ctypes.c_int64 * 3)(g_nchw.stride(0), g_nchw.stride(1), g_nchw.stride(3))

            def prep():
                _native.check(L.vq_prepare_codebook(E.data_ptr(), K, D, E_h.data_ptr(), e2.data_ptr(), cb.data_ptr(), st), "prep")

            def argmin():
                _native.check(L.vq_argmin(z.data_ptr(), B, H * W, D, E.data_ptr(), E_h.data_ptr(), e2.data_ptr(), cb.data_ptr(),
                                          K, idx.data_ptr(), stats.data_ptr(), ws.data_ptr(), ws.numel(), st), "argmin")

            def forward():
                _native.check(L.vq_forward(z.data_ptr(), B, H * W, D, E.data_ptr(), E_h.data_ptr(), e2.data_ptr(), cb.data_ptr(),
                                           K, 0.25, zq.data_ptr(), idx.data_ptr(), loss.data_ptr(), hist.data_ptr(),
                                           stats.data_ptr(), ws.data_ptr(), ws.numel(), st), "forward")

            def backward(g, s, gE=True, gz=True):
                _native.check(L.vq_backward(g.data_ptr(), s, 1.0, None, z.data_ptr(), idx.data_ptr(), E.data_ptr(), B, H * W, D, K,
                                            0.25, N, grad_z.data_ptr() if gz else None, grad_E.data_ptr() if gE else None, st), "bwd")

            prep(); forward(); backward(g_out, s_cl)
            torch.cuda.synchronize()
            out = {"config": name, "dist": dist, "N": N, "K": K}
            out["prep_ms"] = timed(prep, args.reps, flush)[0]
            _native.profile_enable(True)
            out["argmin_ms"] = timed(argmin, args.reps, flush)[0]
            gm = _native.profile_collect()
            out["gemm_kernel_ms"] = statistics.median(gm)
            out["forward_ms"] = timed(forward, args.reps, flush)[0]
            _native.profile_collect()
            _native.profile_enable(False)
            out["stats"] = dict(zip(_native.VQ_STAT_NAMES, stats.tolist()))
            out["backward_ms"] = timed(lambda: backward(g_out, s_cl), args.reps, flush)[0]
            out["backward_noE_ms"] = timed(lambda: backward(g_out, s_cl, gE=False), args.reps, flush)[0]
            out["backward_nchw_ms"] = timed(lambda: backward(g_nchw, s_nc), args.reps, flush)[0]
            flops = 2.0 * N * K * D
            out["gemm_tflops"] = flops / out["gemm_kernel_ms"] / 1e9
            out["gemm_frac_of_bf16_peak"] = out["gemm_tflops"] / peaks["bf16_tflops"]
            out["tokenize_Mlatents_s"] = N / out["argmin_ms"] / 1e3
            out["fwd_bwd_Mlatents_s"] = N / (out["prep_ms"] + out["forward_ms"] + out["backward_ms"]) / 1e3
            out["fwd_bwd_hbm_frac"] = N * 5136 / ((out["prep_ms"] + out["forward_ms"] + out["backward_ms"]) / 1e3) / 1e9 / peaks["hbm_gbs"]
            out["tokenize_hbm_frac"] = N * 1032 / (out["argmin_ms"] / 1e3) / 1e9 / peaks["hbm_gbs"]
            print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in out.items()}))

    # row-major nearest-code search (vq_argmin_rows) at the tables' own widths: the reference's
    # gaussian_to_indices shapes (B*L = 16*256 rows; gaussian_dim 96 of configs/*.yml, 512 of the 3D wrapper's runs; K = 1024)
    # and large ones on the headline codebook size
    if args.rows:
        for label, N, K, Dn, recipes in (("gaussian_to_indices 16x256 rows, D=96, K=1024", 4096, 1024, 96, ("expanded", "diffsq", "cdist_normalized")),
                                         ("3D wrapper 16x256 rows, D=512, K=1024", 4096, 1024, 512, ("cdist_normalized", "expanded")),
                                         ("rows 262144 x 96, K=16384", 262144, 16384, 96, ("expanded", "cdist_normalized")),
                                         ("rows 262144 x 256, K=16384", 262144, 16384, 256, ("expanded", "diffsq", "cdist_normalized")),
                                         ("rows 131072 x 512, K=16384", 131072, 16384, 512, ("cdist_normalized",))):
            g = torch.Generator(device=dev).manual_seed(7)
            table = torch.rand(K, Dn, device=dev, generator=g)
            x = table[torch.randint(0, K, (N,), device=dev, generator=g)] + 0.1 * torch.randn(N, Dn, device=dev, generator=g)
            tab = vq.CodeTable(table)
            out = {"config": "rows: " + label, "N": N, "K": K, "D": Dn}
            for recipe in recipes:
                tab.nearest(x, recipe=recipe)
                torch.cuda.synchronize()
                _native.profile_enable(True)
                t_all = timed(lambda: tab.nearest(x, recipe=recipe), args.reps, flush)[0]
                gm = _native.profile_collect()
                _native.profile_enable(False)
                d_pad = 64 * (1 if Dn <= 64 else 2 if Dn <= 128 else 4 if Dn <= 256 else 8)
                out[recipe] = {"nearest_ms": round(t_all, 4), "Mrows_s": round(N / t_all / 1e3, 2),
                               "gemm_kernel_ms": round(statistics.median(gm), 4) if gm else None,
                               "gemm_frac_of_bf16_peak_algorithmic": round(2.0 * N * K * Dn / statistics.median(gm) / 1e9 / peaks["bf16_tflops"], 4) if gm else None,
                               "gemm_frac_of_bf16_peak_executed": round(2.0 * N * K * d_pad / statistics.median(gm) / 1e9 / peaks["bf16_tflops"], 4) if gm else None,
                               "stats": dict(zip(_native.VQ_STAT_NAMES, tab.last_stats.tolist()))}
            print(json.dumps(out))


if __name__ == "__main__":
    main()
