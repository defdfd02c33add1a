"""Full-size parity on the BENCHMARKED configurations (``pytest -m gpu``): every row of BASELINE.json configs[1..4] goes
through the CPU oracle (its vectorised argmin stage, bit-identical to the scalar definition -- tests/test_oracle_golden.py)
and, where the unmodified reference is staged under baseline/_ref/, through the reference's own CodeBook class on the same
GPU (cuBLAS sgemm, TF32 off, row chunks), with every disagreeing row classified as exact tie / rounding band / real.

Bars: indices, histogram and z_q bit-exact against the oracle; loss <= 1e-6, gradients <= 1e-5 relative; against the
reference class: real == 0.
"""
import os
import sys

import numpy as np
import pytest
import torch

from parity import assert_close, classify_index_mismatches, rel_err

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = 256


@pytest.fixture(scope="module")
def vq():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vq_vae_gan_diffusion_b200 as m
    m.build()
    m._native.check(m._native.lib().vq_device_check(), "vq_device_check")
    return m


def synth(dev, B, H, W, K, dist, seed=1234):
    """bench.py's synthetic workload: init = U(-1/K, 1/K) codebook and N(0,1) latents (near-tie heavy),
    trained = N(0,1) codebook and latents 0.3 sigma around planted codes."""
    g = torch.Generator(device=dev).manual_seed(seed)
    N = B * H * W
    if dist == "init":
        E = (torch.rand(K, D, device=dev, generator=g) * 2 - 1) / K
        z = torch.randn(B, D, H, W, device=dev, generator=g)
    else:
        E = torch.randn(K, D, device=dev, generator=g)
        pick = torch.randint(0, K, (N,), device=dev, generator=g)
        z = (E[pick] + 0.3 * torch.randn(N, D, device=dev, generator=g)).reshape(B, H, W, D).permute(0, 3, 1, 2).contiguous()
    gout = torch.randn(B, H, W, D, device=dev, generator=g).permute(0, 3, 1, 2)    # NHWC memory, like z_q
    return E, z, gout


def reference_codebook_class():
    """The UNMODIFIED reference CodeBook (network/vqvae/submodule/codebook.py), loaded by file path from the staged tree so
    that an install()-ed replacement in sys.modules cannot shadow it.  None when no reference tree is available."""
    import importlib.util
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from stage_reference import staged_root
    root = staged_root()
    if root is None:
        return None
    path = os.path.join(root, "network", "vqvae", "submodule", "codebook.py")
    spec = importlib.util.spec_from_file_location("_reference_codebook_unmodified", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.CodeBook


def reference_indices_on_gpu(RefCodeBook, E, z, chunk_items):
    """The reference class itself on the GPU, row chunks of `chunk_items` batch items (its (N, K) distance matrix would
    not fit otherwise; rows are independent)."""
    assert torch.backends.cuda.matmul.allow_tf32 is False
    ref = RefCodeBook(E.shape[0], D).to(z.device)
    with torch.no_grad():
        ref.codebook.weight.copy_(E)
        out = [ref(z[b:b + chunk_items])[1] for b in range(0, z.shape[0], chunk_items)]
    return torch.cat(out)


FULL = [("cfg4", 256, 32, 32, 16384, "init"), ("cfg4", 256, 32, 32, 16384, "trained"),
        ("cfg3", 256, 32, 32, 8192, "init"), ("cfg3", 256, 32, 32, 8192, "trained"),
        ("cfg2", 64, 16, 16, 1024, "init"), ("cfg2", 64, 16, 16, 1024, "trained")]


@pytest.mark.parametrize("cfg,B,H,W,K,dist", FULL, ids=[f"{c[0]}-{c[5]}" for c in FULL])
def test_every_row_of_the_benchmarked_configs(cfg, B, H, W, K, dist, vq, oracle):
    """BASELINE.json configs[1] (B=64, 16x16, K=1024), configs[2] (B=256, 32x32, K=8192) and configs[3] (K=16384) at their
    real sizes, forward + backward through the module, ALL rows against the oracle."""
    dev = torch.device("cuda:0")
    N = B * H * W
    E, z, gout = synth(dev, B, H, W, K, dist)
    cb = vq.CodeBook(K, D).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(E)
    zt = z.clone().requires_grad_(True)
    z_q, idx, loss = cb(zt)
    (loss + (z_q * gout).sum()).backward()
    stats = cb.stats_dict()
    z_np, E_np = z.cpu().numpy(), E.cpu().numpy()
    ref = oracle.forward(z_np, E_np, beta=0.25, fast=True)

    got_idx = idx.cpu().numpy()
    bad = np.nonzero(got_idx != ref["idx"])[0]
    assert bad.size == 0, f"{bad.size} of {N} indices differ from the oracle, first rows {bad[:8]}"
    assert np.array_equal(cb.last_histogram.cpu().numpy(), ref["hist"])
    assert stats["tie_rows"] == ref["tie_rows"], (stats, ref["tie_rows"])
    assert np.array_equal(z_q.detach().permute(0, 2, 3, 1).reshape(N, D).cpu().numpy(), ref["zq_nhwc"]), "z_q must be bit-exact"
    assert tuple(z_q.stride()) == (H * W * D, 1, W * D, D)
    assert abs(float(loss.detach()) - float(ref["loss"])) <= 1e-6 * abs(float(ref["loss"]))
    del ref["zq_nhwc"]

    g_np = gout.cpu().numpy()                                   # NCHW-shaped view of NHWC memory, strides kept
    gz, gE = oracle.backward(g_np, 1.0, z_np, ref["idx"], E_np, beta=0.25)
    assert_close(zt.grad.cpu().numpy(), gz, "grad_z")
    assert_close(cb.codebook.weight.grad.cpu().numpy(), gE, "grad_E")

    # the reference's own class on this GPU (cuBLAS sgemm order): disagreements must be ties or inside the rounding band
    RefCodeBook = reference_codebook_class()
    if RefCodeBook is not None:
        ref_idx = reference_indices_on_gpu(RefCodeBook, E, z, max(1, 16384 // (H * W))).cpu().numpy()
        cls = classify_index_mismatches(z_np, E_np, got_idx, ref_idx, pair_dist=oracle.pair_dist)
        print(f"[{cfg}-{dist}] vs reference class on GPU: {cls}")
        assert cls["real"] == 0, cls
        assert cls["mismatch"] <= N // 20, cls                  # the init distribution has ~1-2 % tie / band rows


TOK = [(1024, "init"), (1024, "trained"), (2048, "init"), (2048, "trained")]


@pytest.mark.parametrize("K,dist", TOK)
def test_tokeniser_config_every_row(K, dist, vq, oracle):
    """BASELINE.json configs[4]: B=64 of 32x32 latents (512x512 images), K=1024 (small.yml) and 2048 (large.yml),
    encode_indices (vq_argmin), all 65 536 rows against the oracle; int32 / uint16 token streams carry the same values."""
    dev = torch.device("cuda:0")
    B, H, W = 64, 32, 32
    E, z, _ = synth(dev, B, H, W, K, dist, seed=4321)
    cb = vq.CodeBook(K, D).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(E)
    idx = cb.encode_indices(z)
    ref = oracle.forward(z.cpu().numpy(), E.cpu().numpy(), want_zq=False, fast=True)
    assert idx.dtype == torch.int64 and idx.shape == (B * H * W,)
    assert np.array_equal(idx.cpu().numpy(), ref["idx"])
    assert cb.stats_dict()["tie_rows"] == ref["tie_rows"]
    assert torch.equal(cb.encode_indices(z, dtype=torch.int32).long(), idx)
    assert torch.equal(cb.encode_indices(z, dtype=torch.uint16).long(), idx)
    RefCodeBook = reference_codebook_class()
    if RefCodeBook is not None:
        ref_idx = reference_indices_on_gpu(RefCodeBook, E, z, 16).cpu().numpy()
        cls = classify_index_mismatches(z.cpu().numpy(), E.cpu().numpy(), idx.cpu().numpy(), ref_idx, pair_dist=oracle.pair_dist)
        print(f"[cfg5 K={K} {dist}] vs reference class on GPU: {cls}")
        assert cls["real"] == 0, cls


def test_deterministic_backward_is_bit_reproducible(vq, oracle):
    """module.deterministic = True: the codebook-gradient scatter-add runs in 64-bit fixed point (vq_backward_ex) --
    bit-identical from run to run and within 1e-5 of the oracle; the default float atomics are only the latter."""
    dev = torch.device("cuda:0")
    B, H, W, K = 32, 32, 32, 512                                # 64 latents per code on average: heavy collisions
    E, z, gout = synth(dev, B, H, W, K, "trained", seed=77)
    cb = vq.CodeBook(K, D).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(E)

    def grads(det):
        cb.deterministic = det
        cb.codebook.weight.grad = None
        zt = z.clone().requires_grad_(True)
        z_q, idx, loss = cb(zt)
        (loss + (z_q * gout).sum()).backward()
        return idx, zt.grad.clone(), cb.codebook.weight.grad.clone()

    idx, gz0, gE0 = grads(True)
    for _ in range(4):
        _, gz, gE = grads(True)
        assert torch.equal(gE, gE0), "deterministic grad_E differs between runs"
        assert torch.equal(gz, gz0)
    _, _, gE_f = grads(False)
    z_np, E_np = z.cpu().numpy(), E.cpu().numpy()
    gz_o, gE_o = oracle.backward(gout.cpu().numpy(), 1.0, z_np, idx.cpu().numpy(), E_np, beta=0.25)
    assert_close(gE0.cpu().numpy(), gE_o, "deterministic grad_E")
    # (float atomics land in a varying order: with ~64 colliding latents per code only the max-norm bound is meaningful for
    #  elements that are sums of large cancelling terms)
    assert rel_err(gE_f.cpu().numpy(), gE_o) <= 1e-5
    assert_close(gz0.cpu().numpy(), gz_o, "grad_z")
    # frozen encoder input (no grad_z), and the tiny-magnitude regime (fixed-point scale follows the data)
    cb.deterministic = True
    cb.codebook.weight.grad = None
    with torch.no_grad():
        cb.codebook.weight.mul_(1e-6)
    z_q, idx2, loss = cb(z * 1e-6)
    loss.backward()
    _, gE_s = oracle.backward(None, 1.0, z_np * np.float32(1e-6), idx2.cpu().numpy(), cb.codebook.weight.detach().cpu().numpy(), beta=0.25)
    assert rel_err(cb.codebook.weight.grad.cpu().numpy(), gE_s) <= 1e-5, "deterministic grad_E, 1e-6 scale"
