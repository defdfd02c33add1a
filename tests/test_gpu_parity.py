"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).  Every call goes through the C-ABI of
include/vq_b200.h (ctypes), either directly or via the CodeBook module; the checker is the CPU oracle and the
golden vectors of the reference.  Nothing here reads /root/reference.

Bars (north_star / SURVEY 8(c)): indices and histogram bit-exact against the oracle; z_q bit-exact (it is one
IEEE add of one IEEE subtract); loss and gradients within 1e-5 relative; against the reference's golden
indices every disagreement must be an exact tie or inside the rounding band, and is counted.
"""
import os

import numpy as np
import pytest
import torch

from cases import CASES, make_inputs, sample_positions
from parity import assert_close, classify_index_mismatches, rel_err

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GPU_CASES = list(CASES)          # incl. the hand-computed K=4, D=2 case (narrow latents run zero-padded)


@pytest.fixture(scope="module")
def vq():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vq_vae_gan_diffusion_b200 as m
    m.build()                                                # no-op when lib/libvq_b200.so is up to date (it travels with the tree)
    assert os.path.exists(m._native.LIB_PATH), "libvq_b200.so must be built in-tree (no fallback)"
    m._native.check(m._native.lib().vq_device_check(), "vq_device_check")
    assert torch.backends.cuda.matmul.allow_tf32 is False
    return m


def run_module(vq, z_np, E_np, g_nhwc=None, beta=0.25, g_loss=1.0):
    """forward (+ backward) through the CodeBook module on cuda:0 -> numpy dict."""
    dev = torch.device("cuda:0")
    K, D = E_np.shape
    cb = vq.CodeBook(K, D, beta).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(torch.from_numpy(E_np))
    z = torch.from_numpy(z_np).to(dev).requires_grad_(True)
    z_q, idx, loss = cb(z)
    out = dict(z_q=z_q, idx=idx.cpu().numpy(), loss=float(loss.item()),
               hist=cb.last_histogram.cpu().numpy(), stats=cb.stats_dict(),
               zq_rows=z_q.detach().permute(0, 2, 3, 1).reshape(-1, D).cpu().numpy())
    if g_nhwc is not None:
        g = torch.from_numpy(g_nhwc).to(dev).permute(0, 3, 1, 2)          # NHWC memory viewed NCHW, like z_q
        (g_loss * loss + (z_q * g).sum()).backward()
        out["grad_z"] = z.grad.cpu().numpy()
        out["grad_E"] = cb.codebook.weight.grad.cpu().numpy()
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("name", GPU_CASES)
def test_forward_backward_vs_oracle(name, vq, oracle):
    spec = CASES[name]
    z, E, g = make_inputs(spec)
    D, K = spec["D"], spec["K"]
    got = run_module(vq, z, E, g)
    ref = oracle.forward(z, E, beta=0.25)

    assert got["idx"].dtype == np.int64 and got["idx"].shape == ref["idx"].shape
    nbad = int((got["idx"] != ref["idx"]).sum())
    assert nbad == 0, f"{nbad} indices differ from the oracle, first rows {np.nonzero(got['idx'] != ref['idx'])[0][:8]}"
    assert np.array_equal(got["hist"], ref["hist"])
    assert np.array_equal(got["hist"], np.bincount(got["idx"], minlength=K))
    assert np.array_equal(got["zq_rows"], ref["zq_nhwc"]), "z_q must be bit-exact: fl(z + fl(e - z))"
    assert abs(got["loss"] - float(ref["loss"])) <= 1e-6 * abs(float(ref["loss"])) + 1e-12
    assert got["stats"]["tie_rows"] == ref["tie_rows"], (got["stats"], ref["tie_rows"])

    gz, gE = oracle.backward(np.transpose(g, (0, 3, 1, 2)), 1.0, z, ref["idx"], E, beta=0.25)
    assert_close(got["grad_z"], gz, "grad_z")
    assert_close(got["grad_E"], gE, "grad_E")


@pytest.mark.parametrize("name", GPU_CASES)
def test_against_reference_golden(name, vq, oracle):
    spec = CASES[name]
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    z, E, g = make_inputs(spec)
    D = spec["D"]
    got = run_module(vq, z, E, g)
    ref_idx = gold["idx"].astype(np.int64)
    cls = classify_index_mismatches(z, E, got["idx"], ref_idx, pair_dist=oracle.pair_dist)
    assert cls["real"] == 0, cls
    assert cls["mismatch"] <= int(gold["ref_tie_rows"]) + int(gold["ref_ne_fp64"]) + 2, cls
    bad = np.nonzero(got["idx"] != ref_idx)[0]
    assert abs(got["loss"] - float(gold["loss"])) <= 1e-5 * abs(float(gold["loss"])) + 1e-12
    # z_q shape / strides exactly as the reference returned them
    assert tuple(got["z_q"].shape) == tuple(gold["zq_shape"])
    assert tuple(got["z_q"].stride()) == tuple(gold["zq_strides"])
    assert got["z_q"].dtype == torch.float32 and got["z_q"].grad_fn is not None
    keep = np.ones(got["zq_rows"].shape[0], bool)
    keep[bad] = False
    if "zq" in gold.files:
        assert_close(got["zq_rows"][keep], gold["zq"][keep], "z_q vs reference")
        if len(bad) == 0:
            assert_close(got["grad_z"], gold["grad_z"], "grad_z vs reference")
            assert_close(got["grad_E"], gold["grad_E"], "grad_E vs reference")
    else:
        for key, arr in (("zq", got["zq_rows"]), ("grad_z", got["grad_z"]), ("grad_E", got["grad_E"])):
            if key != "zq" and len(bad):
                continue
            pos = sample_positions(arr.size, spec["seed"])
            vals = arr.reshape(-1)[pos]
            refv = gold[key + "_samples"]
            if key == "zq":
                m = ~np.isin(pos // D, bad)
                vals, refv = vals[m], refv[m]
            err = float(np.abs(vals.astype(np.float64) - refv).max() / max(float(gold[key + "_absmax"]), 1e-30))
            assert err <= 1e-5, f"{key}: {err:.3e}"


def test_tokeniser_mode_matches_forward(vq, oracle):
    spec = CASES["cfg2s_init"]
    z, E, _ = make_inputs(spec)
    dev = torch.device("cuda:0")
    cb = vq.CodeBook(spec["K"], spec["D"]).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(torch.from_numpy(E))
        zt = torch.from_numpy(z).to(dev)
        idx_tok = cb.encode_indices(zt)
        none_q, idx_kw, none_l = cb(zt, indices_only=True)
        z_q, idx_fwd, loss = cb(zt)                      # no_grad full forward (what encode_to_z gets)
    assert none_q is None and none_l is None
    ref = oracle.forward(z, E)
    assert np.array_equal(idx_tok.cpu().numpy(), ref["idx"])
    assert torch.equal(idx_tok, idx_kw) and torch.equal(idx_tok, idx_fwd)
    assert z_q.grad_fn is None and not loss.requires_grad


def test_margin_bound_and_operands(vq):
    """The rigorous error bound behind the candidate margin holds on hardware (incl. tensor-core accumulation)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "gpu_diag.py")], capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0 and "DIAG OK" in res.stdout, res.stdout[-3000:] + res.stderr[-2000:]


def test_module_contract(vq):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    cb = vq.CodeBook(num_codebook_vectors=512, latent_dim=256, beta=0.25).to(dev)
    assert list(cb.state_dict().keys()) == ["codebook.weight"]
    assert isinstance(cb.codebook, torch.nn.Embedding) and cb.codebook.weight.shape == (512, 256)
    assert float(cb.codebook.weight.detach().abs().max()) <= 1.0 / 512
    z = torch.randn(2, 256, 4, 4, device=dev, requires_grad=True)
    z_q, idx, loss = cb(z)
    assert z_q.shape == (2, 256, 4, 4) and z_q.stride() == (4 * 4 * 256, 1, 4 * 256, 256)
    assert idx.shape == (32,) and idx.dtype == torch.int64 and loss.dim() == 0 and loss.dtype == torch.float32
    # straight-through: d(sum z_q)/dz == 1 (+ loss term), codebook gets only the beta term
    (z_q.sum() + loss).backward()
    e = cb.codebook.weight.detach()[idx].reshape(2, 4, 4, 256).permute(0, 3, 1, 2)
    exp_gz = 1.0 + 2.0 * (z.detach() - e) / z.numel()
    assert rel_err(z.grad.cpu().numpy(), exp_gz.cpu().numpy()) < 1e-6
    exp_gE = torch.zeros_like(cb.codebook.weight)
    exp_gE.index_add_(0, idx, (0.25 * 2.0 * (e - z.detach()) / z.numel()).permute(0, 2, 3, 1).reshape(-1, 256))
    assert rel_err(cb.codebook.weight.grad.cpu().numpy(), exp_gE.cpu().numpy()) < 1e-5

    # derived state follows in-place weight updates (optimizer.step) and load_state_dict
    opt = torch.optim.SGD(cb.parameters(), lr=10.0)
    opt.step()
    with torch.no_grad():
        _, idx2, _ = cb(z.detach())
        d = (z.detach().permute(0, 2, 3, 1).reshape(-1, 256)[:, None, :] - cb.codebook.weight[None]).pow(2).sum(-1)
        assert (d.gather(1, idx2[:, None]).squeeze(1) <= d.min(dim=1).values * (1 + 1e-5) + 1e-6).all()
    sd = {"codebook.weight": torch.randn(512, 256)}
    cb.load_state_dict(sd)
    with torch.no_grad():
        _, idx3, _ = cb(z.detach())
        d = (z.detach().permute(0, 2, 3, 1).reshape(-1, 256)[:, None, :] - cb.codebook.weight[None]).pow(2).sum(-1)
        assert torch.equal(idx3, d.argmin(1)) or (d.gather(1, idx3[:, None]).squeeze(1) <= d.min(1).values * (1 + 1e-5)).all()

    # edits through weight.data bypass the version counter: while training the derived state is rebuilt every call,
    # for a frozen / no_grad codebook refresh_codebook() does it
    with torch.no_grad():
        new_w = torch.randn(512, 256, device=dev)
    cb.codebook.weight.data.copy_(new_w)
    _, idx4, _ = cb(z.detach().requires_grad_(True))
    d = (z.detach().permute(0, 2, 3, 1).reshape(-1, 256)[:, None, :] - new_w[None]).pow(2).sum(-1)
    assert (d.gather(1, idx4[:, None]).squeeze(1) <= d.min(1).values * (1 + 1e-5)).all()
    with torch.no_grad():
        cb.codebook.weight.data.copy_(-new_w)
        cb.refresh_codebook()
        idx5 = cb.encode_indices(z.detach())
        d = (z.detach().permute(0, 2, 3, 1).reshape(-1, 256)[:, None, :] + new_w[None]).pow(2).sum(-1)
        assert (d.gather(1, idx5[:, None]).squeeze(1) <= d.min(1).values * (1 + 1e-5)).all()

    # frozen codebook (stage-2 models freeze the VQVAE, vqvae.py:103-104): grad only to z
    for p in cb.parameters():
        p.requires_grad_(False)
    cb.zero_grad(set_to_none=True)
    z2 = torch.randn(1, 256, 3, 5, device=dev, requires_grad=True)
    zq2, _, l2 = cb(z2)
    (zq2.sum() + l2).backward()
    assert z2.grad is not None and cb.codebook.weight.grad is None

    # stricter-than-reference input validation
    with pytest.raises(ValueError):
        cb(torch.randn(2, 128, 4, 4, device=dev))
    with pytest.raises(ValueError):
        cb(torch.randn(2, 256, device=dev))
    with pytest.raises(RuntimeError):
        cb(torch.randn(2, 256, 4, 4))
    with pytest.raises(RuntimeError):
        cb(torch.randn(2, 256, 4, 4, device=dev, dtype=torch.float64))


def test_empty_and_noncontiguous_inputs(vq, oracle):
    dev = torch.device("cuda:0")
    cb = vq.CodeBook(64, 256).to(dev)
    z_q, idx, loss = cb(torch.zeros(0, 256, 4, 4, device=dev))
    assert z_q.shape == (0, 256, 4, 4) and idx.numel() == 0 and torch.isnan(loss)
    # channels-last input (non-contiguous NCHW view)
    spec = CASES["small_trained"]
    z, E, _ = make_inputs(spec)
    with torch.no_grad():
        cb.codebook.weight.copy_(torch.from_numpy(E))
        zt = torch.from_numpy(z).to(dev).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        assert not zt.is_contiguous()
        _, idx, _ = cb(zt)
    assert np.array_equal(idx.cpu().numpy(), oracle.forward(z, E)["idx"])


def test_backward_gout_layouts(vq, oracle):
    """Upstream gradient as NCHW-contiguous, channels-last and broadcast (stride 0) tensors."""
    spec = CASES["ragged_trained"]
    z, E, g = make_inputs(spec)
    dev = torch.device("cuda:0")
    cb = vq.CodeBook(spec["K"], spec["D"]).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(torch.from_numpy(E))
    ref = oracle.forward(z, E)
    g_nchw = np.ascontiguousarray(np.transpose(g, (0, 3, 1, 2)))
    gz_ref, gE_ref = oracle.backward(g_nchw, 0.5, z, ref["idx"], E)
    for layout in ("nchw", "nhwc"):
        zt = torch.from_numpy(z).to(dev).requires_grad_(True)
        cb.zero_grad(set_to_none=True)
        z_q, idx, loss = cb(zt)
        gt = torch.from_numpy(g_nchw).to(dev)
        if layout == "nhwc":
            gt = gt.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        torch.autograd.backward([z_q, loss], [gt, torch.tensor(0.5, device=dev)])
        assert_close(zt.grad.cpu().numpy(), gz_ref, f"grad_z ({layout})")
        assert_close(cb.codebook.weight.grad.cpu().numpy(), gE_ref, f"grad_E ({layout})")
    # loss only (no upstream gradient on z_q), and z_q.sum() (broadcast ones)
    zt = torch.from_numpy(z).to(dev).requires_grad_(True)
    z_q, idx, loss = cb(zt)
    loss.backward()
    gz0, _ = oracle.backward(None, 1.0, z, ref["idx"], E)
    assert_close(zt.grad.cpu().numpy(), gz0, "grad_z (loss only)")


@pytest.mark.parametrize("B,H,W,K", [(1, 1, 1, 1), (1, 1, 1, 7), (5, 1, 3, 257), (1, 3, 32, 513), (2, 2, 48, 1), (3, 7, 11, 1025)])
def test_edge_shapes_vs_oracle(B, H, W, K, vq, oracle):
    """Minimum sizes, K crossing code-tile boundaries, N not a multiple of any tile, vector and scalar tile paths."""
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(1000 + B * 131 + H * 17 + W * 3 + K)
    E = rng.standard_normal((K, 256)).astype(np.float32)
    z = rng.standard_normal((B, 256, H, W)).astype(np.float32)
    g = rng.standard_normal((B, H, W, 256)).astype(np.float32)
    got = run_module(vq, z, E, g)
    ref = oracle.forward(z, E)
    assert np.array_equal(got["idx"], ref["idx"])
    assert np.array_equal(got["hist"], ref["hist"])
    assert np.array_equal(got["zq_rows"], ref["zq_nhwc"])
    assert abs(got["loss"] - float(ref["loss"])) <= 1e-6 * abs(float(ref["loss"]))
    gz, gE = oracle.backward(np.transpose(g, (0, 3, 1, 2)), 1.0, z, ref["idx"], E)
    assert_close(got["grad_z"], gz, "grad_z")
    assert_close(got["grad_E"], gE, "grad_E")


def test_degenerate_codebook_mass_fallback(vq, oracle):
    """Every row overflows its candidate list (identical codes): the exact fallback decides all of them, including
    the worklist tail beyond the split-scan capacity; torch.argmin semantics -> lowest index of the tied minimum."""
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(7)
    K, B, H, W = 300, 2, 64, 72                              # N = 9216 > 4096 split rows; HW % 32 == 0
    base = rng.standard_normal((3, 256)).astype(np.float32)
    E = np.repeat(base, K // 3, axis=0)                       # 3 distinct codes, each repeated 100 times
    z = rng.standard_normal((B, 256, H, W)).astype(np.float32)
    cb = vq.CodeBook(K, 256).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(torch.from_numpy(E))
        z_q, idx, loss = cb(torch.from_numpy(z).to(dev))
        hist = cb.last_histogram.clone()
        z_q2, idx2, loss2 = cb(torch.from_numpy(z).to(dev))
    st = cb.stats_dict()
    ref = oracle.forward(z, E)
    assert np.array_equal(idx.cpu().numpy(), ref["idx"])
    assert set(np.unique(ref["idx"]).tolist()) <= {0, 100, 200}
    assert st["tie_rows"] == ref["tie_rows"] == B * H * W
    assert st["fallback_rows"] == B * H * W
    assert np.array_equal(z_q.permute(0, 2, 3, 1).reshape(-1, 256).cpu().numpy(), ref["zq_nhwc"])
    assert np.array_equal(hist.cpu().numpy(), ref["hist"])
    assert abs(float(loss) - float(ref["loss"])) <= 1e-6 * abs(float(ref["loss"]))
    assert torch.equal(idx, idx2) and torch.equal(z_q, z_q2) and float(loss) == float(loss2)   # reproducible


def test_partial_fallback_split_scan(vq, oracle):
    """A few rows overflow their candidate list (a cluster of identical / nearly identical codes), the rest do not:
    the fallback's split scan (several code blocks per row group, merged by the last arriver) decides them."""
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(11)
    K, B, H, W = 4096, 2, 32, 32
    E = rng.standard_normal((K, 256)).astype(np.float32)
    v = rng.standard_normal(256).astype(np.float32)
    E[100:200] = v                                            # 100 identical codes (lowest index must win) ...
    E[1000:1100] = v + 1e-4 * rng.standard_normal((100, 256)).astype(np.float32)   # ... and 100 near copies
    pick = rng.integers(2000, K, size=B * H * W)              # ordinary rows stay away from the cluster
    zf = E[pick] + 0.3 * rng.standard_normal((B * H * W, 256)).astype(np.float32)
    hot = rng.choice(B * H * W, size=21, replace=False)      # 21 rows next to the cluster: 2 row groups, one ragged
    zf[hot] = v + 0.05 * rng.standard_normal((21, 256)).astype(np.float32)
    z = np.ascontiguousarray(zf.reshape(B, H, W, 256).transpose(0, 3, 1, 2))
    cb = vq.CodeBook(K, 256).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(torch.from_numpy(E))
        z_q, idx, loss = cb(torch.from_numpy(z).to(dev))
        st = cb.stats_dict()
        hist = cb.last_histogram.clone()
        losses = [float(cb(torch.from_numpy(z).to(dev))[2]) for _ in range(3)]
        idx_tok = cb.encode_indices(torch.from_numpy(z).to(dev))
    ref = oracle.forward(z, E)
    assert np.array_equal(idx.cpu().numpy(), ref["idx"])
    assert np.array_equal(idx_tok.cpu().numpy(), ref["idx"])
    assert np.array_equal(z_q.permute(0, 2, 3, 1).reshape(-1, 256).cpu().numpy(), ref["zq_nhwc"])
    assert np.array_equal(hist.cpu().numpy(), ref["hist"])
    assert abs(float(loss) - float(ref["loss"])) <= 1e-6 * abs(float(ref["loss"]))
    assert all(v == float(loss) for v in losses), (float(loss), losses)     # reproducible from call to call
    assert st["fallback_rows"] == 21, st
    assert st["tie_rows"] == ref["tie_rows"]


def test_embed_nchw(vq):
    dev = torch.device("cuda:0")
    W = torch.randn(300, 256, device=dev)
    idx = torch.randint(0, 300, (3 * 5 * 7,), device=dev)
    out = vq.vq_embed_nchw(idx, W, 3, 5, 7)
    exp = W[idx].reshape(3, 5, 7, 256).permute(0, 3, 1, 2)
    assert out.is_contiguous() and torch.equal(out, exp)


@pytest.mark.parametrize("K,dist", [(16384, "init"), (16384, "trained"), (8192, "init")])
def test_full_size_properties(K, dist, vq, oracle):
    """BASELINE.json configs[2]/[3] at full size (B=256, 32x32 -> N=262144): size-independent properties plus an
    oracle spot check on a row sample."""
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1234)
    B, D, H, W = 256, 256, 32, 32
    N = B * H * W
    if dist == "init":
        E = (torch.rand(K, D, device=dev, generator=g) * 2 - 1) / K
        z = torch.randn(B, D, H, W, device=dev, generator=g)
    else:
        E = torch.randn(K, D, device=dev, generator=g)
        pick = torch.randint(0, K, (N,), device=dev, generator=g)
        z = (E[pick] + 0.3 * torch.randn(N, D, device=dev, generator=g)).reshape(B, H, W, D).permute(0, 3, 1, 2).contiguous()
    cb = vq.CodeBook(K, D).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(E)
    zt = z.clone().requires_grad_(True)
    z_q, idx, loss = cb(zt)
    gout = torch.randn(B, H, W, D, device=dev, generator=g).permute(0, 3, 1, 2)
    (loss + (z_q * gout).sum()).backward()
    hist = cb.last_histogram
    stats = cb.stats_dict()
    assert int(hist.sum()) == N and torch.equal(hist, torch.bincount(idx, minlength=K))
    assert int(idx.min()) >= 0 and int(idx.max()) < K
    assert stats["fallback_rows"] <= N // 1000, stats
    zrows = z.permute(0, 2, 3, 1).reshape(N, D)
    e = E[idx]
    # loss identity (1 + beta) * mean((e - z)^2)
    exp_loss = 1.25 * float(((e - zrows).double() ** 2).mean())
    assert abs(float(loss.detach()) - exp_loss) <= 1e-5 * exp_loss
    # z_q value and straight-through gradient identities
    assert torch.equal(z_q.permute(0, 2, 3, 1).reshape(N, D), zrows + (e - zrows))
    exp_gz = gout + (2.0 / (N * D)) * (z - e.reshape(B, H, W, D).permute(0, 3, 1, 2))
    assert rel_err(zt.grad[:8].cpu().numpy(), exp_gz[:8].cpu().numpy()) <= 1e-5
    # codebook gradient: column sums equal beta * 2/(ND) * sum(e - z)  (linearity), and rows of unused codes are zero
    gE = cb.codebook.weight.grad
    exp_colsum = (0.25 * 2.0 / (N * D)) * (e - zrows).double().sum(0)
    assert rel_err(gE.double().sum(0).cpu().numpy(), exp_colsum.cpu().numpy()) <= 1e-4
    assert float(gE[hist == 0].abs().max() if (hist == 0).any() else 0.0) == 0.0
    if dist == "trained":
        assert float((idx == pick).double().mean()) > 0.999      # the planted code wins
    # idempotence: quantising the chosen codes returns codes at distance 0 from them
    with torch.no_grad():
        sub = e[:4096].reshape(4, 32, 32, D).permute(0, 3, 1, 2).contiguous()
        idx2 = cb.encode_indices(sub)
        assert torch.equal(E[idx2], e[:4096])
    # oracle spot check: 1024 sampled rows, full K
    rows = torch.randperm(N, device=dev, generator=g)[:1024].sort().values
    zs = zrows[rows].reshape(1024, 1, 1, D).permute(0, 3, 1, 2).contiguous().cpu().numpy()
    ref = oracle.forward(zs, E.cpu().numpy(), want_zq=False)
    assert np.array_equal(idx[rows].cpu().numpy(), ref["idx"])


# ------------------------------------------------------------------------------------------------------------------
# row-major nearest-code search (vq_argmin_rows) and narrow token dtypes (vq_argmin_narrow)
from cases import NN_CASES, make_nn_inputs  # noqa: E402


@pytest.mark.parametrize("name", sorted(NN_CASES))
def test_nearest_rows_vs_oracle_and_reference(name, vq, oracle):
    spec = NN_CASES[name]
    dev = torch.device("cuda:0")
    x, table = make_nn_inputs(spec)
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    tab = vq.CodeTable(torch.from_numpy(table).to(dev))
    idx = tab.nearest(torch.from_numpy(x).to(dev))
    assert idx.shape == (spec["B"], spec["L"]) and idx.dtype == torch.int64
    z = np.ascontiguousarray(x.reshape(-1, spec["D"], 1, 1))
    ref = oracle.forward(z, table, want_zq=False)                       # the oracle at the table's own width
    got = idx.reshape(-1).cpu().numpy()
    assert np.array_equal(got, ref["idx"]), "zero-padded CUDA search must equal the oracle at the native width"
    cls = classify_index_mismatches(z, table, got, gold["idx"].reshape(-1).astype(np.int64), pair_dist=oracle.pair_dist)
    assert cls["real"] == 0, cls
    # non-contiguous input (the (B, D, L) layout of distruibute_dim == 1, viewed (B, L, D)) and the one-shot form
    xt = torch.from_numpy(x).to(dev).permute(0, 2, 1).contiguous().permute(0, 2, 1)
    assert torch.equal(vq.nearest_indices(xt, torch.from_numpy(table).to(dev)), idx)
    for dt in (torch.int32, torch.int16):
        assert torch.equal(tab.nearest(torch.from_numpy(x).to(dev), dtype=dt).to(torch.int64), idx)


def test_nearest_rows_large_ragged(vq, oracle):
    """N not a multiple of any tile, K crossing code tiles, D = 256 rows: the 16-byte row path of prep / select."""
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(77)
    N, K = 4099, 1300
    table = rng.standard_normal((K, 256)).astype(np.float32)
    x = table[rng.integers(0, K, N)] + 0.5 * rng.standard_normal((N, 256)).astype(np.float32)
    idx = vq.nearest_indices(torch.from_numpy(x).to(dev), torch.from_numpy(table).to(dev))
    ref = oracle.forward(np.ascontiguousarray(x.reshape(N, 256, 1, 1)), table, want_zq=False)
    assert np.array_equal(idx.cpu().numpy(), ref["idx"])
    # the same rows through the NCHW entry point (H*W == 1: generic tile path) agree
    cb = vq.CodeBook(K, 256).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(torch.from_numpy(table))
        assert torch.equal(cb.encode_indices(torch.from_numpy(x).to(dev).reshape(N, 256, 1, 1)), idx)


def test_narrow_token_dtypes(vq, oracle):
    spec = CASES["cfg2s_init"]
    z, E, _ = make_inputs(spec)
    dev = torch.device("cuda:0")
    cb = vq.CodeBook(spec["K"], spec["D"]).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(torch.from_numpy(E))
    zt = torch.from_numpy(z).to(dev)
    ref = oracle.forward(z, E, want_zq=False)["idx"]
    for dt in (torch.int64, torch.int32, torch.int16, torch.uint16):
        idx = cb.encode_indices(zt, dtype=dt)
        assert idx.dtype == dt and idx.shape == ref.shape
        assert np.array_equal(idx.cpu().numpy().astype(np.int64), ref)
    with pytest.raises(ValueError):
        vq.CodeBook(40000, 256).to(dev).encode_indices(torch.zeros(1, 256, 1, 1, device=dev), dtype=torch.int16)
    with pytest.raises(ValueError):
        cb.encode_indices(zt, dtype=torch.float32)


# ------------------------------------------------------------------------------------------------------------------
# Drop-in through a VQVAE-shaped caller (network/vqvae/vqvae.py:116-146: encoder -> quant_conv -> CodeBook ->
# post_quant_conv -> decoder) and CUDA-graph capture of the hot path.
class _EagerCodeBook(torch.nn.Module):
    """Test-local restatement of codebook.py:47-111 in eager PyTorch (the reference cannot travel to the GPU box)."""

    def __init__(self, K, D, beta=0.25):
        super().__init__()
        self.beta = beta
        self.codebook = torch.nn.Embedding(K, D)

    def forward(self, z):
        zp = z.permute(0, 2, 3, 1).contiguous()
        zf = zp.view(-1, zp.shape[-1])
        w = self.codebook.weight
        d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(w ** 2, dim=1) - 2 * torch.matmul(zf, w.t())
        idx = torch.argmin(d, dim=1)
        z_q = self.codebook(idx).view(zp.shape)
        loss = torch.mean((z_q.detach() - zp) ** 2 + self.beta * torch.mean((z_q - zp.detach()) ** 2))
        z_q = zp + (z_q - zp).detach()
        return z_q.permute(0, 3, 1, 2), idx, loss


class _TinyVQVAE(torch.nn.Module):
    def __init__(self, codebook):
        super().__init__()
        self.encoder = torch.nn.Sequential(torch.nn.Conv2d(3, 64, 4, 2, 1), torch.nn.SiLU(), torch.nn.Conv2d(64, 256, 4, 2, 1))
        self.quant_conv = torch.nn.Conv2d(256, 256, 1)
        self.codebook = codebook
        self.post_quant_conv = torch.nn.Conv2d(256, 256, 1)
        self.decoder = torch.nn.Sequential(torch.nn.ConvTranspose2d(256, 64, 4, 2, 1), torch.nn.SiLU(),
                                           torch.nn.ConvTranspose2d(64, 3, 4, 2, 1))

    def forward(self, x):                                    # vqvae.py:116-137
        quant_x = self.quant_conv(self.encoder(x))
        z_q, idx, q_loss = self.codebook(quant_x)
        return self.decoder(self.post_quant_conv(z_q)), idx, q_loss


def test_drop_in_through_vqvae_shaped_caller(vq):
    import copy
    dev = torch.device("cuda:0")
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(3)
        K = 512
        ref = _TinyVQVAE(_EagerCodeBook(K, 256)).to(dev)
        with torch.no_grad():
            ref.codebook.codebook.weight.normal_(0, 0.5)
        ours = copy.deepcopy(ref)
        ours.codebook = vq.CodeBook(K, 256).to(dev)
        ours.codebook.load_state_dict(ref.codebook.state_dict())          # same key: codebook.weight
        x = torch.randn(4, 3, 64, 64, device=dev)                       # -> 16x16 latents
        outs = []
        for m in (ref, ours):
            dec, idx, q_loss = m(x)
            (torch.nn.functional.l1_loss(dec, x) + q_loss).backward()
            outs.append((dec.detach(), idx, q_loss.detach(), {n: p.grad.detach().clone() for n, p in m.named_parameters()}))
        (dec_r, idx_r, ql_r, g_r), (dec_o, idx_o, ql_o, g_o) = outs
        assert torch.equal(idx_r, idx_o), int((idx_r != idx_o).sum())
        assert rel_err(dec_o.cpu().numpy(), dec_r.cpu().numpy()) <= 1e-5
        assert abs(float(ql_o) - float(ql_r)) <= 1e-5 * abs(float(ql_r))
        assert set(g_r) == set(g_o)
        for n in g_r:
            assert rel_err(g_o[n].cpu().numpy(), g_r[n].cpu().numpy()) <= 2e-5, n
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_cuda_graph_capture_and_replay(vq, oracle):
    """The C-ABI never allocates or synchronises, so a whole call sequence is capturable; replays follow new input
    contents written into the static buffers."""
    dev = torch.device("cuda:0")
    spec = CASES["cfg2s_trained"]
    z_np, E_np, _ = make_inputs(spec)
    cb = vq.CodeBook(spec["K"], spec["D"]).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(torch.from_numpy(E_np))
        z_static = torch.zeros(z_np.shape, device=dev)
        cb.encode_indices(z_static)                                     # warm-up: one-time attribute setup, workspace
        cb(z_static)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            idx_static = cb.encode_indices(z_static)
            zq_static, idx2_static, loss_static = cb(z_static)
        for shift in (0.0, 0.25):
            z_now = (z_np + np.float32(shift)).astype(np.float32)
            z_static.copy_(torch.from_numpy(z_now))
            graph.replay()
            torch.cuda.synchronize()
            ref = oracle.forward(z_now, E_np)
            assert np.array_equal(idx_static.cpu().numpy(), ref["idx"])
            assert np.array_equal(idx2_static.cpu().numpy(), ref["idx"])
            assert np.array_equal(zq_static.permute(0, 2, 3, 1).reshape(-1, spec["D"]).cpu().numpy(), ref["zq_nhwc"])
            assert abs(float(loss_static) - float(ref["loss"])) <= 1e-6 * float(ref["loss"])


def test_nan_and_inf_rows_follow_the_oracle(vq, oracle):
    """Rows containing NaN / Inf get NaN / inf distances; torch.argmin (and the oracle) return the first NaN.  Such rows take
    the exact fallback and must not disturb their neighbours."""
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(21)
    K, B, H, W = 700, 2, 8, 16
    E = rng.standard_normal((K, 256)).astype(np.float32)
    z = rng.standard_normal((B, 256, H, W)).astype(np.float32)
    z[0, 5, 1, 3] = np.nan
    z[1, 200, 7, 15] = np.inf
    z[1, 17, 0, 0] = -np.inf
    cb = vq.CodeBook(K, 256).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(torch.from_numpy(E))
        z_q, idx, loss = cb(torch.from_numpy(z).to(dev))
        idx_tok = cb.encode_indices(torch.from_numpy(z).to(dev))
    ref = oracle.forward(z, E)
    bad_rows = [0 * H * W + 1 * W + 3, 1 * H * W + 7 * W + 15, 1 * H * W + 0]
    assert ref["idx"][bad_rows[0]] == 0                              # all distances NaN: the first code
    assert np.array_equal(idx.cpu().numpy(), ref["idx"])
    assert np.array_equal(idx_tok.cpu().numpy(), ref["idx"])
    assert cb.stats_dict()["fallback_rows"] == 3
    assert not np.isfinite(float(loss))
    assert ref["idx"][bad_rows[1]] == int(np.argmax(E[:, 200] > 0)) and ref["idx"][bad_rows[2]] == int(np.argmax(E[:, 17] < 0))
    # a NaN inside the codebook makes that code every row's argmin (torch.argmin: the first NaN wins)
    E2 = E.copy()
    E2[433, 7] = np.nan
    zc = np.nan_to_num(z, nan=0.0, posinf=1.0, neginf=-1.0)
    with torch.no_grad():
        cb.codebook.weight.copy_(torch.from_numpy(E2))
        idx2 = cb.encode_indices(torch.from_numpy(zc).to(dev))
    assert (idx2 == 433).all()
    assert np.array_equal(idx2.cpu().numpy(), oracle.forward(zc, E2, want_zq=False)["idx"])


from cases import DIFFSQ_CASES, make_diffsq_inputs  # noqa: E402


@pytest.mark.parametrize("name", sorted(DIFFSQ_CASES))
def test_nearest_rows_diffsq_vs_oracle_and_reference(name, vq, oracle):
    """The broadcast-difference recipe of V_VQDiffusion.sample (v_vq_diffusion.py:114-123) on the same GEMM + exact stage."""
    spec = DIFFSQ_CASES[name]
    dev = torch.device("cuda:0")
    x, E = make_diffsq_inputs(spec)
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    tab = vq.CodeTable(torch.from_numpy(E).to(dev))
    idx = tab.nearest(torch.from_numpy(x).to(dev), recipe="diffsq")
    assert idx.shape == (spec["B"], spec["L"])
    ref = oracle.nearest_diffsq(x, E)
    got = idx.reshape(-1).cpu().numpy()
    assert np.array_equal(got, ref["idx"])
    st = dict(zip(vq._native.VQ_STAT_NAMES, tab.last_stats.tolist()))
    assert st["tie_rows"] == ref["tie_rows"]
    gidx = gold["idx"].reshape(-1).astype(np.int64)
    rows = x.reshape(-1, 256).astype(np.float64)
    for r in np.nonzero(got != gidx)[0]:
        da = ((rows[r] - E[got[r]].astype(np.float64)) ** 2).sum()
        db = ((rows[r] - E[gidx[r]].astype(np.float64)) ** 2).sum()
        assert abs(da - db) <= 70 * 2.0 ** -24 * max(da, db), (r, da, db)
    with pytest.raises(ValueError):
        tab.nearest(torch.from_numpy(x).to(dev), recipe="cosine")


def test_diffsq_large_and_fallback(vq, oracle):
    """Ragged N, duplicated codes (ties -> lowest index) and a cluster of near-identical codes (exact fallback) under the
    broadcast-difference recipe."""
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(31)
    N, K = 2051, 1500
    E = rng.standard_normal((K, 256)).astype(np.float32)
    E[700:760] = E[40]                                                # 60 copies of code 40
    v = rng.standard_normal(256).astype(np.float32)
    E[1000:1100] = v + 1e-4 * rng.standard_normal((100, 256)).astype(np.float32)
    x = E[rng.integers(0, K, N)] + 0.4 * rng.standard_normal((N, 256)).astype(np.float32)
    x[:7] = v + 0.05 * rng.standard_normal((7, 256)).astype(np.float32)
    x[7:12] = E[40]
    tab = vq.CodeTable(torch.from_numpy(E).to(dev))
    idx = tab.nearest(torch.from_numpy(x).to(dev), recipe="diffsq")
    ref = oracle.nearest_diffsq(x, E)
    assert np.array_equal(idx.cpu().numpy(), ref["idx"])
    assert (idx[7:12] == 40).all()
    st = dict(zip(vq._native.VQ_STAT_NAMES, tab.last_stats.tolist()))
    assert st["fallback_rows"] >= 7 and st["tie_rows"] == ref["tie_rows"]


@pytest.mark.parametrize("D,K", [(64, 300), (96, 1000), (128, 64), (3, 5)])
def test_narrow_latent_dim_runs_zero_padded(D, K, vq, oracle):
    """latent_dim < 256 (not used by any reference config, but a legal constructor argument, codebook.py:30-32): the module
    zero-pads to the kernels' 256 channels.  Everything must equal the oracle evaluated at the NATIVE width."""
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(500 + D)
    B, H, W = 2, 8, 8
    E = rng.standard_normal((K, D)).astype(np.float32)
    z = (E[rng.integers(0, K, B * H * W)] + 0.4 * rng.standard_normal((B * H * W, D)).astype(np.float32))
    z = np.ascontiguousarray(z.reshape(B, H, W, D).transpose(0, 3, 1, 2))
    g = rng.standard_normal((B, H, W, D)).astype(np.float32)
    cb = vq.CodeBook(K, D).to(dev)
    assert cb.codebook.weight.shape == (K, D)
    with torch.no_grad():
        cb.codebook.weight.copy_(torch.from_numpy(E))
    zt = torch.from_numpy(z).to(dev).requires_grad_(True)
    z_q, idx, loss = cb(zt)
    assert z_q.shape == (B, D, H, W) and z_q.stride() == (H * W * D, 1, W * D, D)
    gt = torch.from_numpy(g).to(dev).permute(0, 3, 1, 2)
    (loss + (z_q * gt).sum()).backward()
    ref = oracle.forward(z, E)
    assert np.array_equal(idx.cpu().numpy(), ref["idx"])
    assert np.array_equal(cb.last_histogram.cpu().numpy(), ref["hist"])
    assert np.array_equal(z_q.detach().permute(0, 2, 3, 1).reshape(-1, D).cpu().numpy(), ref["zq_nhwc"])
    assert abs(float(loss.detach()) - float(ref["loss"])) <= 1e-6 * float(ref["loss"])
    gz, gE = oracle.backward(np.transpose(g, (0, 3, 1, 2)), 1.0, z, ref["idx"], E)
    assert zt.grad.shape == (B, D, H, W) and cb.codebook.weight.grad.shape == (K, D)
    assert_close(zt.grad.cpu().numpy(), gz, "grad_z")
    assert_close(cb.codebook.weight.grad.cpu().numpy(), gE, "grad_E")
    with torch.no_grad():
        assert np.array_equal(cb.encode_indices(zt.detach()).cpu().numpy(), ref["idx"])
    with pytest.raises(ValueError):
        vq.CodeBook(16, 512).to(dev)(torch.zeros(1, 512, 2, 2, device=dev))


# ------------------------------------------------------------------ token-stream formats
from cases import TOKEN_BLEND_CASES, TOKEN_ONEHOT_CASES, make_token_indices  # noqa: E402

ULP_AT_69 = 2.0 ** -17          # float32 spacing at |log(1e-30)| = 69.08


def reference_log_onehot(x, num_classes):
    """network/vq_diffusion/vq_diffusion.py:29-35, the reference's own lines (any device)."""
    x_onehot = torch.nn.functional.one_hot(x, num_classes)
    permute_order = (0, -1) + tuple(range(1, len(x.size())))
    x_onehot = x_onehot.permute(permute_order)
    return torch.log(x_onehot.float().clamp(min=1e-30))


@pytest.mark.parametrize("name", list(TOKEN_ONEHOT_CASES))
def test_index_to_log_onehot(name, vq):
    """vq_index_to_log_onehot against the reference's lines on the same GPU (bit-exact: same logf), against the CPU
    oracle / golden output of the reference (0 exactly; log(1e-30) within one float32 ulp -- device logf vs host libm)."""
    from oracle.vq_oracle import index_to_log_onehot_np
    dev = torch.device("cuda:0")
    spec = TOKEN_ONEHOT_CASES[name]
    x_np = make_token_indices(spec)
    x = torch.from_numpy(x_np).to(dev)
    out = vq.index_to_log_onehot(x, spec["num_classes"])
    exp = reference_log_onehot(x, spec["num_classes"])
    assert out.dtype == torch.float32 and out.shape == exp.shape and out.is_contiguous()
    assert torch.equal(out, exp)
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))["out"]
    orc = index_to_log_onehot_np(x_np, spec["num_classes"])
    got = out.cpu().numpy()
    assert np.array_equal(orc, gold)
    assert np.array_equal(got == 0.0, gold == 0.0)                       # the one-hot pattern itself: exact
    assert float(np.abs(got - gold).max()) <= ULP_AT_69
    assert torch.equal(out.argmax(1), x)                                 # log_onehot_to_index (vq_diffusion.py:37-38) inverts it


def test_index_to_log_onehot_sizes_and_errors(vq):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    # the VQ-Diffusion shape: 1024 tokens, K + 1 classes; a non-contiguous view; an unaligned-length batch
    for shape, C in (((8, 1024), 1025), ((5, 33), 7), ((1, 4), 1)):
        x = torch.randint(0, C, shape, device=dev, generator=g)
        assert torch.equal(vq.index_to_log_onehot(x, C), reference_log_onehot(x, C))
    xt = torch.randint(0, 40, (16, 12), device=dev, generator=g).t()     # (12, 16) view with strides (1, 12)
    assert torch.equal(vq.index_to_log_onehot(xt, 40), reference_log_onehot(xt, 40))
    assert vq.index_to_log_onehot(torch.zeros((0, 8), dtype=torch.int64, device=dev), 5).shape == (0, 5, 8)
    with pytest.raises(RuntimeError, match="smaller than num_classes"):
        vq.index_to_log_onehot(torch.tensor([[0, 5]], device=dev), 5)
    with pytest.raises(RuntimeError, match="non-negative"):
        vq.index_to_log_onehot(torch.tensor([[0, -1]], device=dev), 5)
    with pytest.raises(RuntimeError, match="LongTensor"):
        vq.index_to_log_onehot(torch.zeros((2, 2), dtype=torch.int32, device=dev), 5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        vq.index_to_log_onehot(torch.zeros((2, 2), dtype=torch.int64), 5)


@pytest.mark.parametrize("name", list(TOKEN_BLEND_CASES))
def test_blend_with_sos_golden(name, vq):
    """vq_mask_replace on the reference's stored draws == the reference's new_indices (vqTransformer.py:117-141)."""
    from oracle.vq_oracle import blend_with_sos_np
    dev = torch.device("cuda:0")
    spec = TOKEN_BLEND_CASES[name]
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    idx_np = make_token_indices(spec)
    out = vq.blend_with_sos(torch.from_numpy(idx_np).to(dev), torch.from_numpy(gold["mask"]).to(dev),
                            torch.from_numpy(gold["random_indices"]).to(dev), spec["sos_token"])
    assert out.dtype == torch.int64 and out.shape == (idx_np.shape[0], idx_np.shape[1] + 1)
    assert np.array_equal(out.cpu().numpy(), gold["new_indices"])
    assert np.array_equal(out.cpu().numpy(), blend_with_sos_np(idx_np, gold["mask"], gold["random_indices"], spec["sos_token"]))


def test_mask_and_replace_consumes_the_generator_like_the_reference(vq):
    """Same seed -> same corrupted tokens as the reference's lines run on the same device (the draws stay in torch)."""
    dev = torch.device("cuda:0")
    indices = torch.randint(0, 1024, (6, 256), device=dev)
    pkeep, vocab, sos = 0.5, 1024, 0
    torch.manual_seed(77)
    got = vq.mask_and_replace(indices, pkeep, vocab, sos)
    after_ours = torch.rand(4, device=dev)
    torch.manual_seed(77)
    sos_tokens = (torch.ones(indices.shape[0], 1) * sos).long().to(dev)                       # vqTransformer.py:117-118
    mask = torch.bernoulli(pkeep * torch.ones(indices.shape, device=indices.device))          # :121-123
    mask = mask.round().to(dtype=torch.int64)                                                 # :124
    random_indices = torch.randint_like(indices, high=vocab)                                  # :127-129
    exp = torch.cat((sos_tokens, mask * indices + (1 - mask) * random_indices), dim=1)        # :138-141
    after_ref = torch.rand(4, device=dev)
    assert torch.equal(got, exp)
    assert torch.equal(after_ours, after_ref)                                                 # generator left in the same state
    assert 0.3 < float((got[:, 1:] == indices).double().mean()) < 0.7


def test_log_onehot_to_index(vq):
    """log_onehot_to_index(log_x) = log_x.argmax(1) (vq_diffusion.py:37-38) against the oracle and torch.argmax on the GPU:
    round trip with index_to_log_onehot, vector and scalar paths, ties (first maximum), NaN (counts as the maximum), -inf rows."""
    from oracle.vq_oracle import log_onehot_to_index_np
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(8)
    for shape, C in (((4, 64), 45), ((3, 37), 33), ((2, 6, 10), 17), ((2, 1024), 1025)):
        x = torch.randint(0, C, shape, device=dev, generator=g)
        assert torch.equal(vq.log_onehot_to_index(vq.index_to_log_onehot(x, C)), x)
        v = torch.randn((shape[0], C) + tuple(shape[1:]), device=dev, generator=g)
        v[0, 3] = v[0, 1]                                   # ties along the class axis: first wins
        v[-1, C // 2].view(-1)[0] = float("nan")
        v[-1, C - 1].view(-1)[0] = float("nan")             # first NaN wins
        v[0, :].view(C, -1)[:, -1] = float("-inf")
        got = vq.log_onehot_to_index(v)
        assert got.dtype == torch.int64 and got.shape == (shape[0],) + tuple(shape[1:])
        assert torch.equal(got, v.argmax(1))
        assert np.array_equal(got.cpu().numpy(), log_onehot_to_index_np(v.cpu().numpy()))
    with pytest.raises(RuntimeError):
        vq.log_onehot_to_index(torch.zeros(2, 3))


def test_cuda_graph_fast_path_matches_eager(vq, oracle):
    """CodeBook.use_cuda_graphs: forward / backward replay captured graphs on static buffers.  Same results as the eager path over
    several optimizer steps (the weight changes between replays), in frozen / no_grad mode, with "static" outputs, and the
    guard against a backward whose buffers a later forward overwrote."""
    dev = torch.device("cuda:0")
    spec = CASES["cfg2s_trained"]
    z_np, E_np, g_np = make_inputs(spec)
    K, D = spec["K"], spec["D"]
    mods = []
    for graphs in (False, True):
        cb = vq.CodeBook(K, D).to(dev)
        with torch.no_grad():
            cb.codebook.weight.copy_(torch.from_numpy(E_np))
        cb.use_cuda_graphs = graphs
        mods.append(cb)
    eager, graphed = mods
    opts = [torch.optim.SGD(m.parameters(), lr=0.5) for m in mods]
    g = torch.from_numpy(g_np).to(dev).permute(0, 3, 1, 2)
    for step in range(4):
        z = torch.from_numpy(z_np).to(dev) * (1.0 + 0.1 * step)
        outs = []
        for m, opt in zip(mods, opts):
            opt.zero_grad(set_to_none=True)
            zt = z.clone().requires_grad_(True)
            z_q, idx, loss = m(zt)
            (loss + (z_q * g).sum()).backward()
            outs.append((z_q.detach().clone(), idx.clone(), loss.detach().clone(), zt.grad.clone(), m.codebook.weight.grad.clone(),
                         m.last_histogram.clone()))
            opt.step()
        a, b = outs
        assert b[0].stride() == a[0].stride() and torch.equal(a[0], b[0]), f"z_q step {step}"
        assert torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and torch.equal(a[5], b[5])
        assert torch.equal(a[3], b[3]), "grad_z"
        assert rel_err(b[4].cpu().numpy(), a[4].cpu().numpy()) <= 1e-5      # scatter-add order differs
        with torch.no_grad():                                              # keep the two in lockstep bit for bit (the in-place
            graphed.codebook.weight.copy_(eager.codebook.weight)           # copy also bumps the version, like the optimizer)
    # frozen codebook under no_grad (the tokenisers' use): cached derived state follows the weight version
    for m in mods:
        m.codebook.weight.requires_grad_(False)
    z = torch.from_numpy(z_np).to(dev)
    with torch.no_grad():
        ra, rb = eager(z), graphed(z)
        assert torch.equal(ra[1], rb[1]) and torch.equal(ra[0], rb[0]) and torch.equal(ra[2], rb[2])
        graphed.codebook.weight.mul_(1.5)
        eager.codebook.weight.mul_(1.5)
        ra, rb = eager(z), graphed(z)
        assert torch.equal(ra[1], rb[1]) and torch.equal(ra[0], rb[0])
    # static outputs alias the graph's buffers; a second forward before the first one's backward is refused
    for m in mods:
        m.codebook.weight.requires_grad_(True)
    graphed.graph_outputs = "static"
    z1 = z.clone().requires_grad_(True)
    q1, _, l1 = graphed(z1)
    q2, _, l2 = graphed(z.clone().requires_grad_(True))
    assert q1.data_ptr() == q2.data_ptr()
    with pytest.raises(RuntimeError, match="overwritten"):
        l1.backward()


# ------------------------------------------------------------------------------------------------------------------
# SURVEY.md 8(f) n2, third site: VQGaussianDiffusion3DWrapper.gaussian_to_indices (normalise + cdist + argmin), and the
# native table widths of all three recipes
from cases import CDIST_CASES, make_cdist_inputs  # noqa: E402
from test_oracle_golden import classify_cdist  # noqa: E402


@pytest.mark.parametrize("name", sorted(CDIST_CASES))
def test_gaussian_to_indices_3d_vs_oracle_and_reference(name, vq, oracle):
    spec = CDIST_CASES[name]
    dev = torch.device("cuda:0")
    x, table = make_cdist_inputs(spec)
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    tab = vq.CodeTable(torch.from_numpy(table).to(dev))
    xt = torch.from_numpy(x).to(dev)
    idx = vq.gaussian_to_indices(xt, tab)
    assert idx.shape == (spec["B"], spec["L"]) and idx.dtype == torch.int64
    assert torch.equal(vq.gaussian_to_indices(xt.unsqueeze(1), tab), idx)            # (B, 1, L, D) is squeezed like :547-548
    ref = oracle.nearest_cdist(x, table)
    got = idx.reshape(-1).cpu().numpy()
    assert np.array_equal(got, ref["idx"]), np.nonzero(got != ref["idx"])[0][:8]
    st = dict(zip(vq._native.VQ_STAT_NAMES, tab.last_stats.tolist()))
    assert st["tie_rows"] == ref["tie_rows"], (st, ref["tie_rows"])
    cls = classify_cdist(x, table, got, gold["idx"].reshape(-1).astype(np.int64))
    assert cls["real"] == 0, cls
    # the other two recipes at the same (native) width
    xr = x.reshape(-1, spec["D"])
    z_grid = np.ascontiguousarray(xr.reshape(-1, 1, 1, spec["D"]).transpose(0, 3, 1, 2))
    assert np.array_equal(tab.nearest(xt, recipe="expanded").reshape(-1).cpu().numpy(), oracle.forward(z_grid, table, want_zq=False)["idx"])
    assert np.array_equal(tab.nearest(xt, recipe="diffsq").reshape(-1).cpu().numpy(), oracle.nearest_diffsq(xr, table)["idx"])
    assert torch.equal(tab.nearest(xt, recipe="cdist_normalized", dtype=torch.int32).long(), idx)


@pytest.mark.parametrize("D,K,N", [(512, 1024, 4096), (96, 2048, 5003), (33, 70, 300), (64, 256, 257), (128, 4100, 2000), (500, 300, 129)])
def test_rows_path_native_widths_large_ragged(D, K, N, vq, oracle):
    """Every contraction width (1, 2, 4, 8 chunks of 64), ragged N / K / D, all three recipes; NaN / Inf rows and a cluster of
    near-identical table rows (more than 64 candidates -> exact-scan fallback) follow the oracle."""
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(1000 + D)
    table = rng.standard_normal((K, D)).astype(np.float32)
    if K >= 300:
        v = rng.standard_normal(D).astype(np.float32)
        table[100:200] = v + 1e-5 * rng.standard_normal((100, D)).astype(np.float32)      # 100 near-identical rows
        table[250:260] = table[7]                                                          # exact duplicates of row 7
    x = table[rng.integers(0, K, N)] + 0.3 * rng.standard_normal((N, D)).astype(np.float32)
    if K >= 300:
        x[:5] = v + 0.01 * rng.standard_normal((5, D)).astype(np.float32)
        x[5:9] = table[7]
    x[9, D // 2] = np.nan
    x[10, 0] = np.inf
    x[11] = 0.0
    tab = vq.CodeTable(torch.from_numpy(table).to(dev))
    xt = torch.from_numpy(x).to(dev)
    for recipe in ("expanded", "diffsq", "cdist_normalized"):
        idx = tab.nearest(xt, recipe=recipe).cpu().numpy()
        if recipe == "expanded":
            ref = oracle.forward(np.ascontiguousarray(x.reshape(N, 1, 1, D).transpose(0, 3, 1, 2)), table, want_zq=False, fast=True)
        elif recipe == "diffsq":
            ref = oracle.nearest_diffsq(x, table)
        else:
            ref = oracle.nearest_cdist(x, table)
        bad = np.nonzero(idx != ref["idx"])[0]
        assert bad.size == 0, (recipe, bad[:8], idx[bad[:8]], ref["idx"][bad[:8]])
        st = dict(zip(vq._native.VQ_STAT_NAMES, tab.last_stats.tolist()))
        assert st["tie_rows"] == ref["tie_rows"], (recipe, st, ref["tie_rows"])
        if K >= 300:
            assert st["fallback_rows"] >= 2, (recipe, st)          # the NaN / Inf rows at least


def test_normalize_rows_matches_oracle(vq, oracle):
    """vq_normalize_rows == F.normalize in the canonical order (bit-exact against the oracle's normalised table)."""
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(3)
    for K, D in ((300, 96), (1024, 512), (17, 5), (64, 256)):
        t = rng.standard_normal((K, D)).astype(np.float32) * 5
        t[3] = 0.0
        tab = vq.CodeTable(torch.from_numpy(t).to(dev))
        E_hat = tab._prepare(True)[0]
        ref = oracle.nearest_cdist(t[:4], t)
        assert np.array_equal(E_hat.cpu().numpy(), ref["table_hat"])


# ------------------------------------------------------------------------------------------------------------------
# SURVEY.md 8(f) n1 (decoder side): post_quant_conv folded into a codebook-sized lookup (postconv.py)
@pytest.mark.parametrize("name,bias", [("cfg2s_trained", True), ("cfg2s_init", True), ("ragged_trained", False)])
def test_folded_post_quant_conv_matches_fp32_conv(name, bias, vq):
    """FoldedPostQuant(codebook, post_quant_conv)(z) against the reference composition post_quant_conv(codebook(z)[0])
    (vqvae.py:131-133) with an fp32 (TF32 off) Conv2d: output, indices, loss and all four gradients within 1e-5."""
    dev = torch.device("cuda:0")
    spec = CASES[name]
    z_np, E_np, g_np = make_inputs(spec)
    K, D = spec["K"], spec["D"]
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(5)
        conv = torch.nn.Conv2d(D, D, 1, bias=bias).to(dev)
        cb = vq.CodeBook(K, D).to(dev)
        with torch.no_grad():
            cb.codebook.weight.copy_(torch.from_numpy(E_np))
        fused = vq.FoldedPostQuant(cb, conv)
        g = torch.from_numpy(g_np).to(dev).permute(0, 3, 1, 2).contiguous()

        def run(folded):
            for p in list(cb.parameters()) + list(conv.parameters()):
                p.grad = None
            z = torch.from_numpy(z_np).to(dev).requires_grad_(True)
            if folded:
                y, idx, loss = fused(z)
            else:
                z_q, idx, loss = cb(z)
                y = conv(z_q)
            (loss + (y * g).sum()).backward()
            return (y.detach(), idx, loss.detach(), z.grad, cb.codebook.weight.grad.clone(), conv.weight.grad.clone(),
                    conv.bias.grad.clone() if bias else None)

        ref, got = run(False), run(True)
        assert got[0].shape == ref[0].shape and got[0].is_contiguous()
        assert torch.equal(got[1], ref[1]) and torch.equal(got[2], ref[2])
        # float64 evaluation of what both compute: y = W_p e[idx] + b_p, dW_p = sum_n g[n] (x) e[idx[n]], db_p = sum_n g[n],
        # grad_z = g W_p + 2 (z - e) / (N D) (straight-through), grad_E = the loss gradient only
        idx = ref[1]
        N = idx.numel()
        e64 = cb.codebook.weight.detach().double()[idx]                                  # (N, D)
        w64 = conv.weight.detach().double().reshape(D, D)
        g64 = g.double().permute(0, 2, 3, 1).reshape(N, D)
        z64 = torch.from_numpy(z_np).to(dev).double().permute(0, 2, 3, 1).reshape(N, D)
        y64 = e64 @ w64.t() + (conv.bias.detach().double() if bias else 0.0)
        truth = {0: y64.reshape(g.shape[0], g.shape[2], g.shape[3], D).permute(0, 3, 1, 2),
                 3: (g64 @ w64 + 2.0 * (z64 - e64) / (N * D)).reshape(g.shape[0], g.shape[2], g.shape[3], D).permute(0, 3, 1, 2),
                 5: (g64.t() @ e64).reshape(conv.weight.shape), 6: g64.sum(0) if bias else None}
        for i, what in ((0, "post_quant_x"), (3, "grad_z"), (5, "grad conv weight"), (6, "grad conv bias")):
            if truth[i] is None:
                continue
            t = truth[i].cpu().numpy()
            err_fold, err_ref = rel_err(got[i].cpu().numpy(), t), rel_err(ref[i].cpu().numpy(), t)
            assert err_fold <= 1e-5, (what, err_fold)
            # the reference composition carries the ulp(|z|) rounding of z + (e - z): up to ~1e-4 of |e| at the init codebook
            assert err_ref <= (2e-4 if spec["dist"] == "init" else 1e-5), (what, err_ref)
        assert rel_err(got[4].cpu().numpy(), ref[4].cpu().numpy()) <= 1e-5, "grad_E (loss gradient only, scatter-add order varies)"
        # under no_grad with a frozen codebook (the tokenisers' / inference use)
        with torch.no_grad():
            y2, idx2, _ = fused(torch.from_numpy(z_np).to(dev))
        assert torch.equal(idx2, ref[1]) and rel_err(y2.cpu().numpy(), ref[0].cpu().numpy()) <= 1e-5
    finally:
        torch.backends.cudnn.allow_tf32 = old
    with pytest.raises(ValueError):
        vq.FoldedPostQuant(cb, torch.nn.Conv2d(D, D, 3, padding=1))


def test_unaligned_and_noncontiguous_codebook_weights(vq, oracle):
    """a weight that is a view at a storage offset that is not 16-byte aligned (a slice of a flat parameter buffer)
    or non-contiguous must give the same results as a dense copy (the module feeds the kernels an aligned clone), and the
    C-ABI refuses misaligned pointers instead of faulting; vq_embed_nchw checks its table and handles narrow ones."""
    dev = torch.device("cuda:0")
    spec = CASES["small_trained"]
    z_np, E_np, _ = make_inputs(spec)
    K, D = spec["K"], spec["D"]
    flat = torch.zeros(K * D + 1, device=dev)
    flat[1:].copy_(torch.from_numpy(E_np).reshape(-1).to(dev))
    z = torch.from_numpy(z_np).to(dev)
    ref = oracle.forward(z_np, E_np)
    for w in (flat[1:].view(K, D), torch.from_numpy(np.ascontiguousarray(E_np.T)).to(dev).t()):
        assert w.data_ptr() % 16 != 0 or not w.is_contiguous()
        cb = vq.CodeBook(K, D).to(dev)
        cb.codebook.weight = torch.nn.Parameter(w)
        zt = z.clone().requires_grad_(True)
        z_q, idx, loss = cb(zt)
        loss.backward()
        assert np.array_equal(idx.cpu().numpy(), ref["idx"])
        assert np.array_equal(z_q.detach().permute(0, 2, 3, 1).reshape(-1, D).cpu().numpy(), ref["zq_nhwc"])
        assert cb.codebook.weight.grad.shape == (K, D)
        assert np.array_equal(cb.encode_indices(z).cpu().numpy(), ref["idx"])
    L = vq._native.lib()
    ws = torch.empty(vq._native.workspace_bytes(32, K, D), dtype=torch.uint8, device=dev)
    dense = torch.from_numpy(E_np).to(dev)
    cb = vq.CodeBook(K, D).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(dense)
    E_h, e2, cbs = cb._derived(cb.codebook.weight)
    idx = torch.empty(32, dtype=torch.int64, device=dev)
    rc = L.vq_argmin(z.data_ptr(), 2, 16, D, flat[1:].data_ptr(), E_h.data_ptr(), e2.data_ptr(), cbs.data_ptr(), K, idx.data_ptr(), None,
                     ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == -1 and b"aligned" in L.vq_last_error()
    # embedding lookup: narrow table, wrong dtype
    t96 = torch.randn(50, 96, device=dev)
    ii = torch.randint(0, 50, (2 * 3 * 5,), device=dev)
    assert torch.equal(vq.vq_embed_nchw(ii, t96, 2, 3, 5), t96[ii].reshape(2, 3, 5, 96).permute(0, 3, 1, 2))
    with pytest.raises(RuntimeError):
        vq.vq_embed_nchw(ii, t96.half(), 2, 3, 5)


def test_codebook_gradient_from_forward_time_sums(vq, oracle):
    """module.scatter_in_forward (what DataParallelVQ's overlapped exchange uses): vq_forward_ex accumulates sum (e - z) per
    code, vq_backward_ex(code_diff_sum=...) turns it into grad_E with one scaling pass.  Same results as the scatter-add path and
    the oracle; grad_z untouched; a forward-only call or a frozen codebook does not accumulate anything."""
    dev = torch.device("cuda:0")
    for name in ("cfg2s_init", "ragged_trained", "dup_rows"):
        spec = CASES[name]
        z_np, E_np, g_np = make_inputs(spec)
        K, D = spec["K"], spec["D"]
        ref = oracle.forward(z_np, E_np)
        gz_o, gE_o = oracle.backward(np.transpose(g_np, (0, 3, 1, 2)), 0.7, z_np, ref["idx"], E_np, beta=0.25)
        cb = vq.CodeBook(K, D).to(dev)
        with torch.no_grad():
            cb.codebook.weight.copy_(torch.from_numpy(E_np))
        cb.scatter_in_forward = True
        z = torch.from_numpy(z_np).to(dev).requires_grad_(True)
        z_q, idx, loss = cb(z)
        g = torch.from_numpy(g_np).to(dev).permute(0, 3, 1, 2)
        (0.7 * loss + (z_q * g).sum()).backward()
        assert np.array_equal(idx.cpu().numpy(), ref["idx"])
        assert np.array_equal(z_q.detach().permute(0, 2, 3, 1).reshape(-1, D).cpu().numpy(), ref["zq_nhwc"])
        assert_close(z.grad.cpu().numpy(), gz_o, "grad_z")
        assert_close(cb.codebook.weight.grad.cpu().numpy(), gE_o, "grad_E from the forward-time sums")
        with torch.no_grad():
            assert torch.equal(cb(z.detach())[1], idx)
