"""Multi-rank parity of the CUDA path (``pytest -m gpu``): batch-sharded CodeBook + DataParallelVQ must give the
single-device results on the concatenated batch -- same indices, histogram, global loss and codebook gradient (<= 1e-5) --
while every rank's own loss and grad_z are the local-mean quantities (the DDP convention, vq_vae_gan_diffusion_b200/dist.py);
and a CodeBook inside a DistributedDataParallel-wrapped model needs nothing special.

Two launch modes, same worker:
* ``gloo`` with both ranks on cuda:0 -- runs on a one-GPU box (gloo all-reduces CUDA tensors through host staging; the
  ranks never wait on each other on the device);
* ``nccl`` with one GPU per rank -- runs when at least two GPUs are visible (the production transport).
(tools/dist_check.py runs the same worker under torchrun on N GPUs.)
"""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def run_checks(rank, world, dev, K=4096, Bl=4, H=32, W=32, deterministic=False, overlap=True, collective="nccl"):
    """The checks of one rank; returns a dict of error figures.  Needs an initialised process group."""
    import vq_vae_gan_diffusion_b200 as vq
    from vq_vae_gan_diffusion_b200.dist import DataParallelVQ
    D = 256
    g = torch.Generator(device=dev).manual_seed(99)           # same seed on every rank: identical global tensors
    E = torch.randn(K, D, device=dev, generator=g)
    B = Bl * world
    z = (E[torch.randint(0, K, (B * H * W,), device=dev, generator=g)] + 0.5 * torch.randn(B * H * W, D, device=dev, generator=g))
    z = z.reshape(B, H, W, D).permute(0, 3, 1, 2).contiguous()
    gout = torch.randn(B, H, W, D, device=dev, generator=g).permute(0, 3, 1, 2)
    sl = slice(rank * Bl, (rank + 1) * Bl)
    one = torch.ones((), device=dev)

    def fresh():
        m = vq.CodeBook(K, D).to(dev)
        m.deterministic = deterministic
        with torch.no_grad():
            m.codebook.weight.copy_(E)
        return m

    # --- stand-alone wrapper: ONE all-reduce of [grad_E / W | hist | loss | 1]
    cb = fresh()
    dp = DataParallelVQ(cb, overlap=overlap, collective=collective)
    zl = z[sl].clone().requires_grad_(True)
    z_q, idx, loss = dp(zl)
    assert dp.collective in ("nccl", "multimem")
    assert dp._step_overlapped == (overlap and not deterministic and world > 1 and dp.collective == "nccl")
    torch.autograd.backward([z_q, loss], [gout[sl], one])
    dp.wait()

    ref = fresh()                                               # single-device run on the concatenated batch
    zf = z.clone().requires_grad_(True)
    zq_f, idx_f, loss_f = ref(zf)
    torch.autograd.backward([zq_f, loss_f], [gout, one])
    loc = fresh()                                               # single-device run on this rank's shard alone
    zs = z[sl].clone().requires_grad_(True)
    zq_s, idx_s, loss_s = loc(zs)
    torch.autograd.backward([zq_s, loss_s], [gout[sl], one])
    torch.cuda.synchronize()

    n = Bl * H * W
    out = dict(
        idx_ok=bool(torch.equal(idx, idx_f[rank * n:(rank + 1) * n])),
        hist_ok=bool(torch.equal(dp.global_histogram, ref.last_histogram)),
        loss_global=abs(float(dp.global_loss) - float(loss_f)) / float(loss_f),
        loss_local=abs(float(loss.detach()) - float(loss_s.detach())) / float(loss_s.detach()),
        grad_E=rel(cb.codebook.weight.grad, ref.codebook.weight.grad),
        grad_z_local=rel(zl.grad, zs.grad),
    )

    # --- inside DistributedDataParallel with a small encoder in front
    torch.manual_seed(11)
    enc = torch.nn.Conv2d(D, D, 1).to(dev)
    cb2 = fresh()

    class Net(torch.nn.Module):
        def __init__(self, e, c):
            super().__init__()
            self.enc, self.cb = e, c

        def forward(self, x):
            return self.cb(self.enc(x))

    ddp = torch.nn.parallel.DistributedDataParallel(Net(enc, cb2), device_ids=[dev.index] if dist.get_backend() == "nccl" else None)
    xq, _, l2 = ddp(z[sl].clone())
    (l2 + (xq * gout[sl]).sum()).backward()
    torch.manual_seed(11)
    enc1 = torch.nn.Conv2d(D, D, 1).to(dev)
    cb1 = fresh()
    q1, _, l1 = cb1(enc1(z.clone()))
    (l1 + (q1 * gout).sum() / world).backward()                # the average of the per-rank objectives
    torch.cuda.synchronize()
    out["ddp_enc_grad"] = rel(enc.weight.grad, enc1.weight.grad)
    out["ddp_grad_E"] = rel(cb2.codebook.weight.grad, cb1.codebook.weight.grad)
    return out


def _worker(rank, world, backend, init_file, out_file):
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    kw = dict(device_id=dev) if backend == "nccl" else {}
    dist.init_process_group(backend, init_method=f"file://{init_file}", rank=rank, world_size=world, **kw)
    try:
        res = run_checks(rank, world, dev)
        res_det = run_checks(rank, world, dev, K=1024, Bl=2, deterministic=True)
        res_hook = run_checks(rank, world, dev, K=2048, Bl=2, overlap=False)
        results = [res, res_det, res_hook]
        if backend == "nccl":
            # the library's own NVLS all-reduce (csrc/vq_allreduce.cuh) instead of NCCL's, two steps on the persistent buffer
            results.append(run_checks(rank, world, dev, K=4096, Bl=2, overlap=False, collective="multimem"))
            results.append(run_checks(rank, world, dev, K=1000, Bl=2, overlap=False, collective="multimem"))
            # "auto" settles on the NVLS kernel where the box has multicast, on NCCL elsewhere -- the same on every rank
            results.append(run_checks(rank, world, dev, K=2048, Bl=2, overlap=False, collective="auto"))
        gathered = [None] * world
        dist.all_gather_object(gathered, tuple(results))
        if rank == 0:
            np.save(out_file, np.array(gathered, dtype=object), allow_pickle=True)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _check(results):
    for rank, pair in enumerate(results):
        for res in pair:
            assert res["idx_ok"] and res["hist_ok"], (rank, res)
            for key in ("loss_global", "loss_local", "grad_E", "grad_z_local"):
                assert res[key] <= 1e-5, (rank, key, res)
            # the encoder in front runs through cuBLAS / cuDNN (TF32 off): a slightly wider band than our own kernels
            assert res["ddp_enc_grad"] <= 1e-4 and res["ddp_grad_E"] <= 1e-4, (rank, res)


def _spawn(world, backend):
    with tempfile.TemporaryDirectory() as td:
        init_file, out_file = os.path.join(td, "rdzv"), os.path.join(td, "out.npy")
        mp.spawn(_worker, args=(world, backend, init_file, out_file), nprocs=world, join=True)
        return list(np.load(out_file, allow_pickle=True))


def test_two_ranks_on_one_gpu_gloo():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    _check(_spawn(2, "gloo"))


def test_ranks_on_separate_gpus_nccl():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs (the one-GPU variant above covers the logic with gloo)")
    _check(_spawn(min(torch.cuda.device_count(), 8), "nccl"))
