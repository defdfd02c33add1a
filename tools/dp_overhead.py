"""Where does the data-parallel step lose time against the single-GPU step?  (torchrun, N >= 2.)

Same cfg4 step as bench.py, timed with CUDA events (max over ranks) in several variants:
  plain       the CodeBook alone on every rank, no wrapper, no collective (N concurrent single-GPU runs)
  no_sync     DataParallelVQ with the gradient exchange switched off (wrapper + hooks only)
  nccl        DataParallelVQ, NCCL all-reduce of the flat buffer
  multimem    DataParallelVQ, the library's NVLS all-reduce kernel
  ar_nccl / ar_multimem   the all-reduce of the 16.9 MB buffer alone, back to back

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dp_overhead.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
import vq_vae_gan_diffusion_b200 as vq  # noqa: E402
from vq_vae_gan_diffusion_b200 import _native  # noqa: E402
from vq_vae_gan_diffusion_b200.dist import DataParallelVQ  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    steps = int(os.environ.get("STEPS", "30"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    wl = bench.WORKLOADS["cfg4"]
    B, H, W, K = wl["B"], wl["H"], wl["W"], wl["K"]
    D = bench.D
    E, z, g_out = bench.make_latents(torch, dev, B, H, W, K, os.environ.get("DIST", "init"), 1234 + rank)
    g_loss = torch.ones((), device=dev)
    out = {"world": world, "steps": steps}

    def timed(fn, n):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def variant(name, wrap):
        cb = vq.CodeBook(K, D, 0.25).to(dev)
        with torch.no_grad():
            cb.codebook.weight.copy_(E)
        dp = wrap(cb)
        zr = z.clone().requires_grad_(True)

        def step():
            cb.refresh_codebook()
            cb.codebook.weight.grad = None
            zr.grad = None
            z_q, idx, loss = (dp or cb)(zr)
            torch.autograd.backward([z_q, loss], [g_out, g_loss])
            if dp is not None:
                dp.wait()

        try:
            out[name + "_ms"] = timed(step, steps)
        except Exception as e:  # noqa: BLE001
            out[name + "_err"] = f"{type(e).__name__}: {str(e)[:300]}"
        if dp is not None:
            dp._hook.remove()
        return cb, dp

    variant("plain", lambda cb: None)

    def no_sync(cb):
        dp = DataParallelVQ(cb)
        dp.sync_grads = False
        return dp

    variant("no_sync", no_sync)
    variant("nccl", lambda cb: DataParallelVQ(cb, collective="nccl"))
    _, dpm = variant("multimem", lambda cb: DataParallelVQ(cb, collective="multimem"))
    variant("nccl_overlap", lambda cb: DataParallelVQ(cb, collective="nccl", overlap=True))
    variant("plain_again", lambda cb: None)

    # the collective alone
    n = K * D + 2 * K + 2
    x = torch.randn(n, device=dev)
    out["ar_nccl_ms"] = timed(lambda: dist.all_reduce(x), 50)
    if dpm is not None and dpm._symm is not None:
        buf, hdl, n_pad = dpm._symm
        st = torch.cuda.current_stream(dev).cuda_stream

        def ar():
            rc = _native.lib().vq_allreduce_multimem(int(hdl.multicast_ptr), int(hdl.signal_pad_ptrs_dev), int(hdl.rank),
                                                     int(hdl.world_size), n_pad, dpm._symm_sync.data_ptr(), int(st))
            _native.check(rc, "vq_allreduce_multimem")

        buf.zero_()
        out["ar_multimem_ms"] = timed(ar, 50)
        for nb in (8, 16, 64, 128):
            os.environ["VQ_AR_MAX_BLOCKS"] = str(nb)
            buf.zero_()
            out[f"ar_multimem_{nb}blocks_ms"] = timed(ar, 50)
        os.environ.pop("VQ_AR_MAX_BLOCKS")
        # correctness of the stand-alone kernel: every rank contributes rank + 1
        buf.fill_(float(rank + 1))
        torch.cuda.synchronize()
        dist.barrier()
        ar()
        torch.cuda.synchronize()
        want = world * (world + 1) / 2
        out["ar_multimem_ok"] = bool((buf == want).all().item())
        dist.barrier()
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
