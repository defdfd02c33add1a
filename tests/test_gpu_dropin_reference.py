"""Drop-in test with the REAL reference models (``pytest -m gpu``): the unmodified reference tree staged under git-ignored
``baseline/_ref/`` (tools/stage_reference.py; /root/reference itself does not exist on the GPU box) is imported twice --
once as shipped, once after ``vq_vae_gan_diffusion_b200.install()`` has put the B200 CodeBook under the reference's module
path -- and the reference's own callers of the hot path are run on identical weights and inputs:

  VQVAE.forward / VQVAE.encode              network/vqvae/vqvae.py:116-146
  VQTransformer.encode_to_z / z_to_image    network/vqTransformer/vqTransformer.py:64-103
  VQDiffusion.encode_to_z                   network/vqDiffusion/vqDiffusion.py:140-156
  VQVAE forward + backward (the training step's use, worker/vqganVqvaeWorker.py:181,246-247)

Everything around the CodeBook (encoder, 1x1 convs, decoder) is the same PyTorch code in both runs, so the latents entering
the quantiser are bit-identical and the comparison isolates the swapped class.
"""
import copy
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from parity import classify_index_mismatches, rel_err

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _purge_reference_modules():
    for name in [m for m in sys.modules if m == "network" or m.startswith("network.")]:
        del sys.modules[name]


@pytest.fixture(scope="module")
def env():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from stage_reference import staged_root
    root = staged_root()
    if root is None:
        pytest.skip("no reference tree (baseline/_ref is staged by __graft_entry__.build() in the build container)")
    import yaml
    import vq_vae_gan_diffusion_b200 as vq
    vq.build()
    cfg = yaml.load(open(os.path.join(root, "configs", "training_config_small.yml")), Loader=yaml.FullLoader)
    a = cfg["architecture"]["vqvae"]
    # a narrower encoder / decoder keeps the test fast; the quantiser's interface (256 channels, K = 1024) is the config's own
    a["intermediate_channels"] = [32, 32, 64, 64, 128]
    a["num_residual_blocks_encoder"] = 1
    a["num_residual_blocks_decoder"] = 1
    cfg["dataset"]["dataset_name"] = "Oxford102Flower"          # 3 x 256 x 256 images -> 16 x 16 latents
    t = cfg["architecture"]["vqvae_transformer"]
    t["n_layer"], t["n_head"], t["n_embd"] = 1, 2, 32
    if root not in sys.path:
        sys.path.insert(0, root)

    def build_models(ours: bool):
        """Import the reference packages fresh (with or without install()) and build VQVAE under a fixed seed."""
        _purge_reference_modules()
        if ours:
            vq.install()
        vqvae_mod = importlib.import_module("network.vqvae.vqvae")
        cls = vqvae_mod.CodeBook
        assert (cls is vq.CodeBook) == ours, "install() must decide which CodeBook network.vqvae.vqvae imports"
        torch.manual_seed(0)
        model = vqvae_mod.VQVAE(config=copy.deepcopy(cfg)).cuda().eval()
        vt_mod = importlib.import_module("network.vqTransformer.vqTransformer")
        vd_mod = importlib.import_module("network.vqDiffusion.vqDiffusion")
        mods = dict(vqvae=model, VQTransformer=vt_mod.VQTransformer, VQDiffusion=vd_mod.VQDiffusion)
        _purge_reference_modules()
        if vq.REFERENCE_MODULE in sys.modules:
            del sys.modules[vq.REFERENCE_MODULE]
        return mods

    ref = build_models(False)
    ours = build_models(True)
    sd_r, sd_o = ref["vqvae"].state_dict(), ours["vqvae"].state_dict()
    assert list(sd_r) == list(sd_o) and "codebook.codebook.weight" in sd_o
    assert all(torch.equal(sd_r[k], sd_o[k]) for k in sd_r), "same seed, same RNG consumption -> same weights"
    torch.backends.cudnn.deterministic = True
    return dict(vq=vq, cfg=cfg, ref=ref, ours=ours)


def _images(n, seed=3):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.rand(n, 3, 256, 256, device="cuda", generator=g) * 2 - 1


def _plant_codebook(env, x):
    """Give BOTH models a data-like codebook (rows = real encoder outputs + noise), so that nearest codes are well separated
    and the two implementations must agree on every index (the init codebook case is covered separately, with ties counted)."""
    ref, ours = env["ref"]["vqvae"], env["ours"]["vqvae"]
    with torch.no_grad():
        q = ref.quant_conv(ref.encoder(x))
        rows = q.permute(0, 2, 3, 1).reshape(-1, q.shape[1])
        g = torch.Generator(device="cuda").manual_seed(5)
        K = ref.codebook.codebook.weight.shape[0]
        pick = torch.randint(0, rows.shape[0], (K,), device="cuda", generator=g)
        W = rows[pick] + 0.05 * rows.std() * torch.randn(K, rows.shape[1], device="cuda", generator=g)
        ref.codebook.codebook.weight.copy_(W)
        ours.codebook.codebook.weight.copy_(W)
    ours.codebook.refresh_codebook()


def _latents(model, x):
    with torch.no_grad():
        return model.quant_conv(model.encoder(x))


def test_vqvae_encode_and_forward_init_codebook(env, oracle):
    """Constructor-initialised codebook U(-1/K, 1/K): near-tie heavy; disagreements must be ties / rounding band."""
    ref, ours = env["ref"]["vqvae"], env["ours"]["vqvae"]
    x = _images(4)
    with torch.no_grad():
        zq_r, idx_r, loss_r = ref.encode(x)
        zq_o, idx_o, loss_o = ours.encode(x)
        dec_r, idx_r2, _ = ref(x)
        dec_o, idx_o2, _ = ours(x)
    assert zq_o.shape == zq_r.shape and zq_o.stride() == zq_r.stride() and zq_o.dtype == zq_r.dtype
    assert idx_o.shape == idx_r.shape and idx_o.dtype == idx_r.dtype and loss_o.shape == loss_r.shape
    assert torch.equal(idx_o, idx_o2)
    z = _latents(ref, x)
    assert torch.equal(z, _latents(ours, x)), "the latents entering the quantiser must be identical in both runs"
    E = ref.codebook.codebook.weight.detach()
    cls = classify_index_mismatches(z.cpu().numpy(), E.cpu().numpy(), idx_o.cpu().numpy(), idx_r.cpu().numpy(), pair_dist=oracle.pair_dist)
    print("VQVAE.encode, init codebook:", cls)
    assert cls["real"] == 0, cls
    assert abs(float(loss_o) - float(loss_r)) <= 1e-5 * abs(float(loss_r))
    same = (idx_o == idx_r).view(4, -1).all(dim=1)              # batch items whose indices all agree: identical downstream
    if same.any():
        assert rel_err(zq_o[same].cpu().numpy(), zq_r[same].cpu().numpy()) <= 1e-5
        assert rel_err(dec_o[same].cpu().numpy(), dec_r[same].cpu().numpy()) <= 1e-4


def test_vqvae_forward_backward_planted_codebook(env):
    """The training step's use: decoded, indices, q_loss, then backward through decoder -> CodeBook -> encoder."""
    ref, ours = env["ref"]["vqvae"], env["ours"]["vqvae"]
    x = _images(4, seed=9)
    _plant_codebook(env, x)
    outs = {}
    for tag, m in (("ref", ref), ("ours", ours)):
        m.zero_grad(set_to_none=True)
        dec, idx, q_loss = m(x)
        (torch.nn.functional.l1_loss(dec, x) + q_loss).backward()
        outs[tag] = (dec.detach(), idx, q_loss.detach(),
                     m.codebook.codebook.weight.grad.clone(), m.quant_conv.weight.grad.clone(),
                     next(m.encoder.parameters()).grad.clone(),
                     m.post_quant_conv.weight.grad.clone())
    r, o = outs["ref"], outs["ours"]
    assert torch.equal(o[1], r[1]), "planted codebook: every index must agree"
    assert abs(float(o[2]) - float(r[2])) <= 1e-5 * abs(float(r[2]))
    assert rel_err(o[0].cpu().numpy(), r[0].cpu().numpy()) <= 1e-4
    for i, what in ((3, "codebook.weight.grad"), (4, "quant_conv.weight.grad"), (5, "first encoder weight grad"),
                    (6, "post_quant_conv.weight.grad")):
        err = rel_err(o[i].cpu().numpy(), r[i].cpu().numpy())
        assert err <= 2e-4, (what, err)                          # cuDNN / cuBLAS kernels in between: not our arithmetic


def test_vqtransformer_and_vqdiffusion_encode_to_z(env):
    ref, ours = env["ref"], env["ours"]
    x = _images(2, seed=21)
    _plant_codebook(env, x)
    cfg_t = copy.deepcopy(env["cfg"])
    cfg_t["architecture"]["model_name"] = "vqvae_transformer"
    res = {}
    for tag, side in (("ref", ref), ("ours", ours)):
        vt = side["VQTransformer"](side["vqvae"], device="cuda", config=copy.deepcopy(cfg_t)).cuda()
        zq, idx = vt.encode_to_z(x)                              # vqTransformer.py:64-81 (prints x.shape like the reference)
        img = vt.z_to_image(idx)                                 # vqTransformer.py:83-103: uses codebook.codebook(indices)
        # VQDiffusion.encode_to_z (vqDiffusion.py:140-156) only touches self.vqvae: run the reference's method body on a
        # minimal holder instead of building its U-Net
        holder = type("Holder", (), {"vqvae": side["vqvae"]})()
        zq_d, idx_d = side["VQDiffusion"].encode_to_z(holder, x)
        res[tag] = (zq, idx, img, zq_d, idx_d)
    r, o = res["ref"], res["ours"]
    assert o[1].shape == r[1].shape == (2, 256) and o[1].dtype == torch.int64
    assert torch.equal(o[1], r[1]) and torch.equal(o[4], r[4]) and torch.equal(o[1], o[4])
    assert o[0].stride() == r[0].stride()
    assert rel_err(o[0].cpu().numpy(), r[0].cpu().numpy()) <= 1e-5
    assert rel_err(o[3].cpu().numpy(), r[3].cpu().numpy()) <= 1e-5
    assert rel_err(o[2].cpu().numpy(), r[2].cpu().numpy()) <= 1e-4


def test_frozen_vqvae_and_state_dict_roundtrip(env):
    """Stage-2 use: a frozen VQVAE (vqvae.py:103-104 sets requires_grad False on the codebook) loaded from a checkpoint."""
    ref, ours = env["ref"]["vqvae"], env["ours"]["vqvae"]
    x = _images(2, seed=33)
    _plant_codebook(env, x)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    with torch.no_grad():
        ours.codebook.codebook.weight.zero_()
    ours.load_state_dict(sd)                                     # in-place copy_: bumps the version counter -> derived state refreshed
    for p in ours.parameters():
        p.requires_grad = False
    with torch.no_grad():
        _, idx_r, _ = ref.encode(x)
        _, idx_o, _ = ours.encode(x)
    assert torch.equal(idx_o, idx_r)
    for p in ours.parameters():
        p.requires_grad = True


def test_vqvae_forward_backward_with_both_convolutions_folded(env):
    """The reference VQVAE's own forward (vqvae.py:116-137) with quant_conv -> CodeBook -> post_quant_conv replaced by the opt-in
    FoldedVQ: decoded images, indices, q_loss and the gradients of encoder / both 1x1 convolutions / codebook
    against the unmodified reference model on identical weights and inputs."""
    vq = env["vq"]
    ref, ours = env["ref"]["vqvae"], env["ours"]["vqvae"]
    x = _images(4, seed=41)
    _plant_codebook(env, x)
    fused = vq.FoldedVQ(ours.quant_conv, ours.codebook, ours.post_quant_conv)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                      # the reference's convolutions in fp32, like the fold's arithmetic
    try:
        ref.zero_grad(set_to_none=True)
        dec_r, idx_r, loss_r = ref(x)
        (torch.nn.functional.l1_loss(dec_r, x) + loss_r).backward()
        ours.zero_grad(set_to_none=True)
        h = ours.encoder(x)
        assert fused.pre.fusable(h), "16 x 16 latents of 256 channels are what the fused kernel takes"
        post_quant_x, idx_o, loss_o = fused(h)                   # in place of vqvae.py:128-133
        dec_o = ours.decoder(post_quant_x)
        (torch.nn.functional.l1_loss(dec_o, x) + loss_o).backward()
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert torch.equal(idx_o, idx_r), "planted codebook: every index must agree"
    assert abs(float(loss_o) - float(loss_r)) <= 1e-5 * abs(float(loss_r))
    assert rel_err(dec_o.detach().cpu().numpy(), dec_r.detach().cpu().numpy()) <= 1e-4
    for what, po, pr in (("codebook.weight.grad", ours.codebook.codebook.weight, ref.codebook.codebook.weight),
                         ("quant_conv.weight.grad", ours.quant_conv.weight, ref.quant_conv.weight),
                         ("quant_conv.bias.grad", ours.quant_conv.bias, ref.quant_conv.bias),
                         ("post_quant_conv.weight.grad", ours.post_quant_conv.weight, ref.post_quant_conv.weight),
                         ("first encoder weight grad", next(ours.encoder.parameters()), next(ref.encoder.parameters()))):
        err = rel_err(po.grad.cpu().numpy(), pr.grad.cpu().numpy())
        assert err <= 2e-4, (what, err)                          # cuDNN / cuBLAS kernels in between: not our arithmetic
