"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/vq_b200.h declares; host-side argument validation works without a GPU; the module refuses to run on CPU."""
import ctypes
import os
import re

import pytest
import torch

import vq_vae_gan_diffusion_b200 as vq
from vq_vae_gan_diffusion_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vq_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vq_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    path = _native.build()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/vq_b200.h but not exported"
    assert set(syms) == set(_native.EXPORTED_SYMBOLS), "ctypes signatures out of sync with the header"
    assert _native.lib().vq_abi_version() == 1


def test_host_side_validation_without_gpu():
    L = _native.lib()
    assert L.vq_padded_codes(1) == 256 and L.vq_padded_codes(256) == 256 and L.vq_padded_codes(257) == 512
    out = ctypes.c_size_t(0)
    assert L.vq_workspace_bytes(1000, 1024, 256, ctypes.byref(out)) == 0 and out.value > 1000 * 256 * 2
    small = out.value
    assert L.vq_workspace_bytes(100000, 1024, 256, ctypes.byref(out)) == 0 and out.value > small
    # the row-major searches take any width up to 512 (narrower contraction -> smaller operand image); beyond that: unsupported
    assert L.vq_workspace_bytes(1000, 1024, 96, ctypes.byref(out)) == 0 and out.value < small
    assert L.vq_workspace_bytes(1000, 1024, 512, ctypes.byref(out)) == 0 and out.value > small
    assert L.vq_workspace_bytes(1000, 1024, 513, ctypes.byref(out)) == -2          # VQ_E_UNSUPPORTED
    assert L.vq_workspace_bytes(1000, 1024, 0, ctypes.byref(out)) == -2
    assert b"256" in L.vq_last_error()
    assert L.vq_workspace_bytes(-1, 1024, 256, ctypes.byref(out)) == -1            # VQ_E_INVALID
    # compute entry points reject bad arguments before touching the device
    assert L.vq_argmin(None, 1, 1, 64, None, None, None, None, 16, None, None, None, 0, None) == -2
    assert L.vq_forward(None, 4, 4, 256, None, None, None, None, 16, 0.25, None, None, None, None, None, None, 0, None) == -1


def test_token_entry_points_validate_without_gpu():
    L = _native.lib()
    assert L.vq_index_to_log_onehot(None, -1, 4, 8, 1e-30, None, None) == -1      # VQ_E_INVALID: negative shape
    assert L.vq_index_to_log_onehot(None, 2, 4, 0, 1e-30, None, None) == -1       # num_classes < 1
    assert b"num_classes" in L.vq_last_error()
    assert L.vq_index_to_log_onehot(None, 0, 4, 8, 1e-30, None, None) == 0        # empty batch: nothing to launch
    assert L.vq_index_to_log_onehot(None, 2, 4, 8, 1e-30, None, None) == -1       # null pointers
    assert L.vq_mask_replace(None, None, None, 0, -1, 4, None, None) == -1
    assert L.vq_mask_replace(None, None, None, 0, 0, 4, None, None) == 0
    assert L.vq_mask_replace(None, None, None, 0, 2, 4, None, None) == -1
    with pytest.raises(RuntimeError, match="no CPU path"):
        vq.index_to_log_onehot(torch.zeros((2, 3), dtype=torch.int64), 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        vq.mask_and_replace(torch.zeros((2, 3), dtype=torch.int64), 0.5, 16, 0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        vq.blend_with_sos(torch.zeros((2, 3), dtype=torch.int64), torch.ones(2, 3), torch.zeros((2, 3), dtype=torch.int64), 0)


def test_module_has_no_cpu_path():
    cb = vq.CodeBook(32, 256)
    assert list(cb.state_dict().keys()) == ["codebook.weight"]
    with pytest.raises(RuntimeError, match="no CPU path"):
        cb(torch.zeros(1, 256, 2, 2))
    with pytest.raises(ValueError):
        cb(torch.zeros(1, 256, 2))


def test_constructor_matches_reference_rng_consumption():
    """Same seed -> same weights as the reference constructor (nn.Embedding init, then uniform_(-1/K, 1/K))."""
    torch.manual_seed(123)
    cb = vq.CodeBook(num_codebook_vectors=48, latent_dim=256, beta=0.3)
    torch.manual_seed(123)
    emb = torch.nn.Embedding(48, 256)
    emb.weight.data.uniform_(-1 / 48, 1 / 48)
    assert torch.equal(cb.codebook.weight, emb.weight)
    assert (cb.num_codebook_vectors, cb.latent_dim, cb.beta) == (48, 256, 0.3)


def test_install_registers_reference_module_path():
    import sys
    vq.install()
    from network.vqvae.submodule.codebook import CodeBook  # noqa: the reference's import line (vqvae.py:17)
    assert CodeBook is vq.CodeBook
    del sys.modules["network.vqvae.submodule.codebook"]


def test_build_is_a_noop_when_sources_are_unchanged():
    """The build stamp is a content hash of the sources (the tree is copied between machines, so file times are useless)."""
    path = _native.build()
    before = os.path.getmtime(path)
    assert _native.build() == path and os.path.getmtime(path) == before
    assert open(path + ".src").read().strip() == _native._source_digest()


def test_nearest_search_host_side_validation():
    from vq_vae_gan_diffusion_b200 import nearest
    assert nearest._index_bits(torch.int64, 10 ** 6) == 64 and nearest._index_bits(torch.int32, 10 ** 6) == 32
    assert nearest._index_bits(torch.int16, 32768) == 16 and nearest._index_bits(torch.uint16, 65536) == 16
    with pytest.raises(ValueError):
        nearest._index_bits(torch.int16, 32769)
    with pytest.raises(ValueError):
        nearest._index_bits(torch.uint16, 65537)
    with pytest.raises(ValueError):
        nearest._index_bits(torch.float32, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        vq.CodeTable(torch.zeros(4, 96))                       # CPU table
    with pytest.raises(ValueError):
        vq.CodeTable(torch.zeros(4, 600))                      # wider than the kernels (D <= 512)
    L = _native.lib()
    # row-major / narrow entry points validate before touching the device: width beyond 512, unknown recipe, bad index width
    assert L.vq_argmin_rows(None, 0, 513, None, None, None, None, 16, 0, None, 64, None, None, 0, None) == -2
    assert L.vq_argmin_rows(None, 0, 128, None, None, None, None, 16, 7, None, 64, None, None, 0, None) == -1
    assert L.vq_argmin_rows(None, 0, 128, None, None, None, None, 16, 2, None, 24, None, None, 0, None) == -1
    assert L.vq_normalize_rows(None, 4, 600, None, None) == -2
    assert L.vq_argmin_narrow(None, 1, 1, 64, None, None, None, None, 16, None, 32, None, None, 0, None) == -2


def test_narrow_latent_dim_is_accepted_and_wide_rejected():
    cb = vq.CodeBook(8, 64)
    assert cb.codebook.weight.shape == (8, 64)
    with pytest.raises(RuntimeError, match="no CPU path"):
        cb(torch.zeros(1, 64, 2, 2))
    with pytest.raises(ValueError):
        vq.CodeBook(8, 512)(torch.zeros(1, 512, 2, 2))


def test_folded_quant_conv_entry_points_validate_without_gpu():
    """vq_forward_qconv / vq_prepare_quant_conv / vq_pack_stats / vq_allreduce_multimem refuse bad arguments before any device
    work, and the Python modules over them have no CPU path (the CodeBook they end in raises)."""
    L = _native.lib()
    one = ctypes.c_void_p(16)                                    # a non-null, 16-byte aligned dummy (never dereferenced)
    args_tail = (one, one, one, one, 16, 0.25, None, one, one, None, None, one, 1 << 20, None)
    assert L.vq_forward_qconv(one, 2, 35, 256, one, one, None, one, *args_tail) == -2          # HW % 128 != 0: VQ_E_UNSUPPORTED
    assert b"HW" in L.vq_last_error()
    assert L.vq_forward_qconv(one, 2, 128, 64, one, one, None, one, *args_tail) == -2          # only 256 -> 256
    assert L.vq_forward_qconv(None, 2, 128, 256, one, one, None, one, *args_tail) == -1        # null activations
    assert L.vq_forward_qconv(ctypes.c_void_p(20), 2, 128, 256, one, one, None, one, *args_tail) == -1   # misaligned
    assert L.vq_prepare_quant_conv(None, one, one, None) == -1
    assert L.vq_pack_stats(None, one, 16, one, None) == -1
    assert L.vq_pack_stats(one, one, 0, one, None) == -1
    assert L.vq_allreduce_multimem(one, one, 0, 1, 64, one, None) == -1                        # world < 2
    assert L.vq_allreduce_multimem(one, one, 0, 2, 60, one, None) == -1                        # not a multiple of 4 * world
    assert L.vq_allreduce_multimem(one, one, 0, 2, 64, None, None) == -1                       # no local_sync words
    conv = torch.nn.Conv2d(256, 256, 1)
    cb = vq.CodeBook(32, 256)
    fused = vq.FoldedQuantConv(conv, cb)
    h = torch.zeros(1, 256, 8, 16)
    assert not fused.fusable(h)                                  # CPU tensor: the composition, whose CodeBook refuses it
    with pytest.raises(RuntimeError, match="no CPU path"):
        fused(h)
    with pytest.raises(RuntimeError, match="no CPU path"):
        vq.FoldedVQ(conv, cb, torch.nn.Conv2d(256, 256, 1))(h)
    with pytest.raises(ValueError):
        vq.FoldedQuantConv(torch.nn.Conv2d(256, 256, 3, padding=1), cb)
    with pytest.raises(ValueError):
        vq.FoldedQuantConv(torch.nn.Conv2d(256, 128, 1), cb)


def test_conv_oracle_matches_torch_conv2d():
    """oracle/vq_oracle.py: quant_conv_fp32 restates nn.Conv2d(C, C, 1) (vqvae.py:83) on the CPU."""
    import numpy as np
    from oracle.vq_oracle import quant_conv_fp32
    torch.manual_seed(1)
    conv = torch.nn.Conv2d(256, 256, 1)
    h = torch.randn(2, 256, 4, 8)
    with torch.no_grad():
        want = conv(h).numpy()
    got = quant_conv_fp32(h.numpy(), conv.weight.detach().numpy(), conv.bias.detach().numpy())
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    got_nb = quant_conv_fp32(h.numpy(), conv.weight.detach().numpy().reshape(256, 256))
    assert np.abs(got_nb + conv.bias.detach().numpy()[None, :, None, None] - want).max() <= 1e-5 * np.abs(want).max()


def test_split_precision_model_of_the_folded_convolution():
    """The arithmetic vq_qconv_prep_kernel runs, modelled in numpy: operands split hi + lo in fp16 after an exact power-of-two
    scaling (per row for h, per tensor for W), the three products hi hi + lo hi + hi lo accumulated in fp32 sixteen
    contraction elements at a time -- pessimistically with TRUNCATION after every accumulation, which is the worst the tensor
    core's fp32 accumulator is known to do -- then z = fl(acc / (s_h s_W) + b).  The model must sit inside the 1e-5 bar the GPU
    test holds the kernel to, next to the fp32 CPU convolution (the oracle), for O(1), tiny and large activations."""
    import numpy as np
    from oracle.vq_oracle import quant_conv_fp32
    rng = np.random.default_rng(0)
    D, N = 256, 512

    def split16(x, sc):
        xs = (x * sc).astype(np.float32)
        hi = xs.astype(np.float16)
        lo = (xs - hi.astype(np.float32)).astype(np.float16)
        return hi.astype(np.float64), lo.astype(np.float64)

    def trunc32(x):
        y = x.astype(np.float32)
        bad = np.abs(y.astype(np.float64)) > np.abs(x)
        y[bad] = np.nextafter(y[bad], np.float32(0))
        return y

    for h_scale in (1.0, 1e-4, 3e3):
        h = (h_scale * rng.standard_normal((N, D))).astype(np.float32)
        W = (rng.uniform(-1, 1, (D, D)) / 16).astype(np.float32)
        b = (h_scale * rng.uniform(-1, 1, D) / 16).astype(np.float32)
        z64 = h.astype(np.float64) @ W.astype(np.float64).T + b
        s_h = np.ldexp(1.0, 15 - np.frexp(np.abs(h).max(1))[1])[:, None].astype(np.float32)
        s_w = np.float32(np.ldexp(1.0, 15 - np.frexp(np.abs(W).max())[1]))
        hh, hl = split16(h, s_h)
        wh, wl = split16(W, s_w)
        assert np.abs(hh + hl - (h * s_h).astype(np.float64)).max() <= 2.0 ** -8           # <= 2^-23 of the row's top binade (2^15)
        acc = np.zeros((N, D), np.float32)
        for dc in range(4):
            for A, B in ((hh, wh), (hl, wh), (hh, wl)):
                for k in range(4):
                    s = slice(64 * dc + 16 * k, 64 * dc + 16 * k + 16)
                    acc = trunc32(acc.astype(np.float64) + A[:, s] @ B[:, s].T)
        z = (acc.astype(np.float64) / s_h.astype(np.float64) / float(s_w) + b).astype(np.float32)
        err = np.abs(z - z64).max() / np.abs(z64).max()
        z_cpu = quant_conv_fp32(h.T.reshape(1, D, N, 1).copy(), W, b)[0, :, :, 0].T
        err_cpu = np.abs(z_cpu - z64).max() / np.abs(z64).max()
        assert err <= 5e-6 and err_cpu <= 5e-6, (h_scale, err, err_cpu)
