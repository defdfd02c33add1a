"""GPU parity of the folded quant_conv (encoder side): ``FoldedQuantConv(quant_conv, codebook)(h)`` against
the reference composition ``codebook(quant_conv(h))`` (/root/reference/network/vqvae/vqvae.py:128-131).

Bars: the convolution's output z within 1e-5 (max-norm relative) of the fp32 CPU convolution (oracle/vq_oracle.py:
quant_conv_fp32) and of a float64 evaluation; everything downstream computed from EXACTLY that z -- indices, histogram and z_q
bit-exact against the C oracle fed with the returned z, and bit-identical to the unfused CodeBook run on it (same operand image,
norms and scales); loss within 1e-6; all gradients within 1e-5 of a float64 evaluation.
"""
import numpy as np
import pytest
import torch

from parity import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vq():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vq_vae_gan_diffusion_b200 as m
    m.build()
    m._native.check(m._native.lib().vq_device_check(), "vq_device_check")
    return m


def make_case(B, H, W, K, dist, seed, h_scale=1.0):
    rng = np.random.default_rng(seed)
    D = 256
    Wc = (rng.uniform(-1, 1, (D, D)) / 16).astype(np.float32)
    bc = (h_scale * rng.uniform(-1, 1, D) / 16).astype(np.float32)
    h = (h_scale * rng.standard_normal((B, D, H, W))).astype(np.float32)
    if dist == "init":
        E = rng.uniform(-1.0 / K, 1.0 / K, (K, D)).astype(np.float32)
    else:
        # codes scattered around the convolution's outputs: every latent has a near code, as after training
        zr = (np.matmul(Wc.astype(np.float64)[None], h.astype(np.float64).reshape(B, D, H * W)) + bc.astype(np.float64)[None, :, None])
        zr = zr.transpose(0, 2, 1).reshape(-1, D)
        E = (zr[rng.integers(0, zr.shape[0], K)] + 0.3 * zr.std() * rng.standard_normal((K, D))).astype(np.float32)
    g = rng.standard_normal((B, H, W, D)).astype(np.float32)
    return h, Wc, bc, E, g


CASES = [
    # B, H, W, K, dist, bias, h_scale
    (2, 16, 16, 1024, "trained", True, 1.0),          # 4 row tiles
    (3, 32, 32, 2048, "init", True, 1.0),             # 24 row tiles
    (1, 8, 16, 300, "trained", False, 1.0),           # one row tile, ragged K, no bias
    (40, 32, 32, 1024, "trained", True, 1.0),         # 320 row tiles: several per CTA, both accumulators, ring wrap-around
    (2, 16, 32, 512, "trained", True, 1e-4),          # small-magnitude activations: the per-row operand scale at work
    (2, 16, 32, 512, "init", True, 3e3),              # large ones
]


@pytest.mark.parametrize("B,H,W,K,dist,bias,h_scale", CASES)
def test_folded_quant_conv(B, H, W, K, dist, bias, h_scale, vq, oracle):
    from oracle.vq_oracle import quant_conv_fp32
    dev = torch.device("cuda:0")
    D = 256
    h_np, W_np, b_np, E_np, g_np = make_case(B, H, W, K, dist, 1000 + K + H, h_scale)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        conv = torch.nn.Conv2d(D, D, 1, bias=bias).to(dev)
        cb = vq.CodeBook(K, D).to(dev)
        with torch.no_grad():
            conv.weight.copy_(torch.from_numpy(W_np).reshape(D, D, 1, 1))
            if bias:
                conv.bias.copy_(torch.from_numpy(b_np))
            cb.codebook.weight.copy_(torch.from_numpy(E_np))
        fused = vq.FoldedQuantConv(conv, cb)
        h = torch.from_numpy(h_np).to(dev).requires_grad_(True)
        assert fused.fusable(h)
        z_q, idx, loss = fused(h)
        g = torch.from_numpy(g_np).to(dev).permute(0, 3, 1, 2)
        (loss + (z_q * g).sum()).backward()
        torch.cuda.synchronize()
        z = fused.last_z
        z_np = z.cpu().numpy()
        N = B * H * W

        # 1. the convolution: fp32 CPU oracle and float64 truth
        z64 = np.matmul(W_np.astype(np.float64)[None], h_np.astype(np.float64).reshape(B, D, H * W))
        if bias:
            z64 = z64 + b_np.astype(np.float64)[None, :, None]
        z64 = z64.reshape(B, D, H, W)
        z_cpu = quant_conv_fp32(h_np, W_np, b_np if bias else None)
        e64, e_cpu = rel_err(z_np, z64), rel_err(z_np, z_cpu)
        assert e64 <= 1e-5 and e_cpu <= 1e-5, (e64, e_cpu)
        assert rel_err(z_cpu, z64) <= 1e-5                       # the yardstick itself

        # 2. downstream of z: the C oracle on the returned z, bit for bit
        ref = oracle.forward(z_np, E_np, 0.25)
        assert np.array_equal(idx.cpu().numpy(), ref["idx"]), "indices differ from the oracle on the fused kernel's z"
        assert np.array_equal(cb.last_histogram.cpu().numpy(), ref["hist"])
        assert np.array_equal(z_q.detach().permute(0, 2, 3, 1).reshape(-1, D).cpu().numpy(), ref["zq_nhwc"])
        assert abs(float(loss.detach()) - float(ref["loss"])) <= 1e-6 * abs(float(ref["loss"]))

        # 3. ... and the unfused CodeBook on the same z: identical operand image / norms / scales -> identical everything
        cb2 = vq.CodeBook(K, D).to(dev)
        with torch.no_grad():
            cb2.codebook.weight.copy_(torch.from_numpy(E_np))
            zq2, idx2, loss2 = cb2(z)
        assert torch.equal(idx2, idx) and torch.equal(zq2, z_q.detach()) and torch.equal(loss2, loss.detach())
        assert cb2.stats_dict() == cb.stats_dict()

        # 4. gradients against float64: grad_z = g + 2 (z - e) / (N D), then the convolution's backward
        e_sel = E_np.astype(np.float64)[ref["idx"]]
        zr64 = z64.transpose(0, 2, 3, 1).reshape(N, D)
        gz64 = g_np.astype(np.float64).reshape(N, D) + 2.0 * (zr64 - e_sel) / (N * D)
        hr64 = h_np.astype(np.float64).transpose(0, 2, 3, 1).reshape(N, D)
        truth_h = (gz64 @ W_np.astype(np.float64)).reshape(B, H, W, D).transpose(0, 3, 1, 2)
        truth_w = gz64.T @ hr64
        truth_b = gz64.sum(0)
        gE64 = np.zeros((K, D))
        np.add.at(gE64, ref["idx"], 0.25 * 2.0 * (e_sel - zr64) / (N * D))
        assert rel_err(h.grad.cpu().numpy(), truth_h) <= 1e-5
        assert rel_err(conv.weight.grad.reshape(D, D).cpu().numpy(), truth_w) <= 1e-5
        if bias:
            assert rel_err(conv.bias.grad.cpu().numpy(), truth_b) <= 1e-5
        assert rel_err(cb.codebook.weight.grad.cpu().numpy(), gE64) <= 1e-5

        # 5. call-to-call reproducibility and the no_grad / frozen path (cached weight images)
        with torch.no_grad():
            zq3, idx3, loss3 = fused(torch.from_numpy(h_np).to(dev))
            zq4, idx4, loss4 = fused(torch.from_numpy(h_np).to(dev))
        assert torch.equal(idx3, idx) and torch.equal(zq3, z_q.detach()) and torch.equal(loss3, loss.detach())
        assert torch.equal(idx4, idx) and torch.equal(zq4, zq3)
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_folded_quant_conv_other_shapes_run_the_composition(vq, oracle):
    """H * W not a multiple of 128: the reference's composition (library convolution, then the CodeBook)."""
    dev = torch.device("cuda:0")
    D, K = 256, 256
    torch.manual_seed(3)
    conv = torch.nn.Conv2d(D, D, 1).to(dev)
    cb = vq.CodeBook(K, D).to(dev)
    fused = vq.FoldedQuantConv(conv, cb)
    h = torch.randn(2, D, 5, 7, device=dev)
    assert not fused.fusable(h)
    z_q, idx, loss = fused(h)
    z_q2, idx2, loss2 = cb(conv(h))
    assert torch.equal(idx, idx2) and torch.equal(z_q, z_q2) and torch.equal(loss, loss2)
    with pytest.raises(ValueError):
        vq.FoldedQuantConv(torch.nn.Conv2d(D, D, 3, padding=1), cb)
    # the C-ABI refuses what the kernel cannot tile
    L = vq._native.lib()
    rc = L.vq_forward_qconv(h.data_ptr(), 2, 35, D, h.data_ptr(), h.data_ptr(), 0, h.data_ptr(), h.data_ptr(), h.data_ptr(), h.data_ptr(),
                            h.data_ptr(), K, 0.25, 0, h.data_ptr(), h.data_ptr(), 0, 0, h.data_ptr(), 1 << 20, 0)
    assert rc != 0 and b"HW" in L.vq_last_error()


@pytest.mark.parametrize("B,H,W,K,dist", [(3, 16, 32, 1024, "trained"), (2, 32, 32, 512, "init")])
def test_folded_vq_both_convolutions(B, H, W, K, dist, vq, oracle):
    """FoldedVQ(quant_conv, codebook, post_quant_conv)(h) against the reference composition post_quant_conv(codebook(quant_conv(h))[0])
    (vqvae.py:128-133): same indices / loss as FoldedQuantConv, output and all six gradients within 1e-5 of a float64 evaluation."""
    dev = torch.device("cuda:0")
    D = 256
    h_np, Wq, bq, E_np, g_np = make_case(B, H, W, K, dist, 77 + K)
    rng = np.random.default_rng(5)
    Wp = (rng.uniform(-1, 1, (D, D)) / 16).astype(np.float32)
    bp = (rng.uniform(-1, 1, D) / 16).astype(np.float32)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        qconv, pconv = torch.nn.Conv2d(D, D, 1).to(dev), torch.nn.Conv2d(D, D, 1).to(dev)
        cb = vq.CodeBook(K, D).to(dev)
        with torch.no_grad():
            qconv.weight.copy_(torch.from_numpy(Wq).reshape(D, D, 1, 1)); qconv.bias.copy_(torch.from_numpy(bq))
            pconv.weight.copy_(torch.from_numpy(Wp).reshape(D, D, 1, 1)); pconv.bias.copy_(torch.from_numpy(bp))
            cb.codebook.weight.copy_(torch.from_numpy(E_np))
        fused = vq.FoldedVQ(qconv, cb, pconv)
        h = torch.from_numpy(h_np).to(dev).requires_grad_(True)
        y, idx, loss = fused(h)
        assert y.shape == (B, D, H, W) and y.is_contiguous()
        g = torch.from_numpy(g_np).to(dev).permute(0, 3, 1, 2).contiguous()        # gradient on post_quant_x, NCHW
        (loss + (y * g).sum()).backward()
        torch.cuda.synchronize()
        N = B * H * W

        # same quantiser decisions as the encoder-side fold alone
        with torch.no_grad():
            zq1, idx1, loss1 = vq.FoldedQuantConv(qconv, cb)(torch.from_numpy(h_np).to(dev))
        assert torch.equal(idx1, idx) and torch.equal(loss1, loss.detach())
        idx_np = idx.cpu().numpy()
        ref = oracle.forward(fused.pre.last_z.cpu().numpy(), E_np, 0.25)
        assert np.array_equal(idx_np, ref["idx"])

        # float64 truth on those indices
        hr = h_np.astype(np.float64).transpose(0, 2, 3, 1).reshape(N, D)
        z64 = hr @ Wq.astype(np.float64).T + bq
        e = E_np.astype(np.float64)[idx_np]
        y64 = e @ Wp.astype(np.float64).T + bp
        gy = g_np.astype(np.float64).reshape(N, D)                                 # rows of the NCHW gradient
        g_zq = gy @ Wp.astype(np.float64)
        gz = g_zq + 2.0 * (z64 - e) / (N * D)
        nchw = lambda a: a.reshape(B, H, W, D).transpose(0, 3, 1, 2)
        assert rel_err(y.detach().cpu().numpy(), nchw(y64)) <= 1e-5
        assert rel_err(h.grad.cpu().numpy(), nchw(gz @ Wq.astype(np.float64))) <= 1e-5
        assert rel_err(qconv.weight.grad.reshape(D, D).cpu().numpy(), gz.T @ hr) <= 1e-5
        assert rel_err(qconv.bias.grad.cpu().numpy(), gz.sum(0)) <= 1e-5
        assert rel_err(pconv.weight.grad.reshape(D, D).cpu().numpy(), gy.T @ e) <= 1e-5
        assert rel_err(pconv.bias.grad.cpu().numpy(), gy.sum(0)) <= 1e-5
        gE = np.zeros((K, D))
        np.add.at(gE, idx_np, 0.25 * 2.0 * (e - z64) / (N * D))                    # the loss gradient only, as in the reference
        assert rel_err(cb.codebook.weight.grad.cpu().numpy(), gE) <= 1e-5
    finally:
        torch.backends.cudnn.allow_tf32 = old
