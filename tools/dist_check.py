"""Multi-GPU parity check (run under torchrun on N GPUs): batch-sharded CodeBook + DataParallelVQ over NCCL must give
the single-device results on the concatenated batch -- same indices, histogram, loss, weight.grad (<= 1e-5), and each
rank's grad_z equal to its slice of the single-device grad_z.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py
"""
from __future__ import annotations

import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import vq_vae_gan_diffusion_b200 as vq  # noqa: E402
from vq_vae_gan_diffusion_b200.dist import DataParallelVQ  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    K, D, H, W, Bl = 4096, 256, 32, 32, 8
    g = torch.Generator(device=dev).manual_seed(99)           # same seed on every rank: identical global tensors
    E = torch.randn(K, D, device=dev, generator=g)
    B = Bl * world
    z = (E[torch.randint(0, K, (B * H * W,), device=dev, generator=g)] + 0.5 * torch.randn(B * H * W, D, device=dev, generator=g))
    z = z.reshape(B, H, W, D).permute(0, 3, 1, 2).contiguous()
    gout = torch.randn(B, H, W, D, device=dev, generator=g).permute(0, 3, 1, 2)
    sl = slice(rank * Bl, (rank + 1) * Bl)

    cb = vq.CodeBook(K, D).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(E)
    dp = DataParallelVQ(cb)
    zl = z[sl].clone().requires_grad_(True)
    z_q, idx, loss = dp(zl)
    torch.autograd.backward([z_q, loss], [gout[sl], torch.ones((), device=dev)])
    dp.wait()
    torch.cuda.synchronize()

    ref = vq.CodeBook(K, D).to(dev)                            # single-device run on the concatenated batch
    with torch.no_grad():
        ref.codebook.weight.copy_(E)
    zf = z.clone().requires_grad_(True)
    zq_f, idx_f, loss_f = ref(zf)
    torch.autograd.backward([zq_f, loss_f], [gout, torch.ones((), device=dev)])
    torch.cuda.synchronize()

    n = Bl * H * W
    ok = True
    ok &= torch.equal(idx, idx_f[rank * n:(rank + 1) * n])
    ok &= torch.equal(dp.global_histogram, ref.last_histogram)
    e_loss = abs(float(dp.global_loss) - float(loss_f)) / float(loss_f)
    e_gE = rel(cb.codebook.weight.grad, ref.codebook.weight.grad)
    e_gz = rel(zl.grad, zf.grad[sl])
    ok &= e_loss < 1e-5 and e_gE < 1e-5 and e_gz < 1e-5
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"dist_check world={world}: loss_err={e_loss:.2e} gradE_err={e_gE:.2e} gradz_err={e_gz:.2e} "
              f"-> {'OK' if int(flag) else 'FAILED'}")
    dist.destroy_process_group()
    return 0 if int(flag) else 1


if __name__ == "__main__":
    sys.exit(main())
