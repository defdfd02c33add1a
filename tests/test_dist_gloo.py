"""World-size-2 CPU (gloo) tests of the data-parallel path: batch-sharded latents, replicated codebook, ONE exchange
step (a SUM all-reduce of the flat buffer [grad_E / W | hist | loss | 1]).  The compute on each rank is the CPU oracle
(the CUDA module has no CPU path); what is under test is the host logic of vq_vae_gan_diffusion_b200/dist.py and its
gradient convention: local-mean loss and grad_z per rank (what DDP-averaged upstream layers expect), codebook gradient
averaged over ranks == single-device gradient on the concatenated batch -- stand-alone (DataParallelVQ) and inside a
DistributedDataParallel-wrapped model.  The same checks run on the CUDA kernels in tests/test_gpu_dist.py.
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


class OracleCodeBook(torch.nn.Module):
    """CPU stand-in with the CodeBook module's interface towards dist.py, computed by the oracle (test infrastructure only)."""

    def __init__(self, E, beta=0.25):
        super().__init__()
        from oracle.vq_oracle import COracle
        self.oracle = COracle()
        K, D = E.shape
        self.codebook = torch.nn.Embedding(K, D)
        with torch.no_grad():
            self.codebook.weight.copy_(torch.from_numpy(E))
        self.beta = beta
        self.grad_scale = 1.0
        self.grad_alloc = None
        self.scatter_in_forward = False
        self.scatter_alloc = None
        self.scatter_ready = None
        self.deterministic = False
        self.last_histogram = None

    def forward(self, z):
        mod = self

        class Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, z, w):
                out = mod.oracle.forward(z.detach().numpy(), w.detach().numpy(), mod.beta)
                B, D, H, W = z.shape
                ctx.save_for_backward(z, w)
                ctx.idx = out["idx"]
                ctx.scat = None
                if mod.scatter_in_forward:                    # vq_forward_ex: per-code sums of (e - z), where the wrapper wants them
                    zr = np.ascontiguousarray(np.moveaxis(z.detach().numpy().reshape(B, D, -1), 1, 2)).reshape(-1, D)
                    S = np.zeros_like(w.detach().numpy())
                    np.add.at(S, out["idx"], w.detach().numpy()[out["idx"]] - zr)
                    ctx.scat = mod.scatter_alloc(S.shape[0], S.shape[1], w.device) if mod.scatter_alloc is not None else torch.empty(S.shape)
                    ctx.scat.copy_(torch.from_numpy(S))
                    ctx.scat_ready = mod.scatter_ready
                mod.last_histogram = torch.from_numpy(out["hist"])
                idx = torch.from_numpy(out["idx"])
                ctx.mark_non_differentiable(idx)
                zq = torch.from_numpy(out["zq_nhwc"]).reshape(B, H, W, D).permute(0, 3, 1, 2)
                return zq, idx, torch.tensor(float(out["loss"]))

            @staticmethod
            def backward(ctx, g_zq, _gi, g_loss):
                z, w = ctx.saved_tensors
                g = None if g_zq is None else np.ascontiguousarray(g_zq.numpy())
                gz, gE = mod.oracle.backward(g, float(g_loss), z.detach().numpy(), ctx.idx, w.detach().numpy(), mod.beta)
                if ctx.scat is not None:                       # vq_backward_ex(code_diff_sum=...): one scaling pass
                    if ctx.scat_ready is not None:
                        ctx.scat_ready(ctx.scat)
                    coef = np.float32(2.0 * float(g_loss) / (ctx.idx.size * z.shape[1]))
                    return torch.from_numpy(gz), ctx.scat * float(np.float32(mod.beta) * coef * np.float32(mod.grad_scale))
                gE_t = torch.from_numpy(gE) * mod.grad_scale          # vq_backward_ex's grad_E_scale
                if mod.grad_alloc is not None:                        # ... written where the wrapper wants it
                    out = mod.grad_alloc(gE_t.shape[0], gE_t.shape[1], gE_t.device)
                    out.copy_(gE_t)
                    gE_t = out
                return torch.from_numpy(gz), gE_t

        return Fn.apply(z, self.codebook.weight)


def _worker(rank, world, init_file, out_file):
    from cases import CASES, make_inputs
    from vq_vae_gan_diffusion_b200.dist import DataParallelVQ
    from oracle.vq_oracle import COracle
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        spec = CASES["small_trained"]                      # B = 2 -> one batch item per rank
        z, E, g = make_inputs(spec)
        B = spec["B"]
        assert B % world == 0
        sl = slice(rank * (B // world), (rank + 1) * (B // world))
        orc = COracle()

        # --- 1. the stand-alone wrapper, ONE async all-reduce per step: post-backward (hook) and overlapped (forward-time sums)
        gt = torch.from_numpy(np.ascontiguousarray(np.transpose(g[sl], (0, 3, 1, 2))))
        cb0 = OracleCodeBook(E)
        dp0 = DataParallelVQ(cb0, overlap=False, collective="auto")
        assert dp0.collective == "nccl"       # "auto" only picks the NVLS kernel for CUDA ranks on the nccl backend
        z0 = torch.from_numpy(z[sl].copy()).requires_grad_(True)
        q0, _, l0 = dp0(z0)
        torch.autograd.backward([q0, l0], [gt, torch.tensor(1.0)])
        dp0.wait()
        gE_hook = cb0.codebook.weight.grad.numpy().copy()
        hist_hook, loss_hook = dp0.global_histogram.numpy(), float(dp0.global_loss)
        dp0._hook.remove()

        cb = OracleCodeBook(E)
        dp = DataParallelVQ(cb, overlap=True)
        zt = torch.from_numpy(z[sl].copy()).requires_grad_(True)
        z_q, idx, loss = dp(zt)
        assert dp._step_overlapped and dp._fwd_work is not None
        torch.autograd.backward([z_q, loss], [gt, torch.tensor(1.0)])
        dp.wait()
        gE_wrap = cb.codebook.weight.grad.numpy().copy()
        hist_wrap, loss_wrap = dp.global_histogram.numpy(), float(dp.global_loss)

        # gradient accumulation, DDP style: micro-steps under no_sync() accumulate LOCAL gradients, the last one exchanges
        # the accumulated .grad (which is not aliased to the flat buffer then: the wrapper exchanges a packed copy)
        cb.codebook.weight.grad = None
        with dp.no_sync():
            z_q, _, loss2 = dp(zt)
            torch.autograd.backward([z_q, loss2], [gt, torch.tensor(1.0)])
        z_q, _, loss2 = dp(zt)
        assert not dp._step_overlapped                      # a .grad is being accumulated: the post-backward path takes over
        torch.autograd.backward([z_q, loss2], [gt, torch.tensor(1.0)])
        dp.wait()
        gE_accum = cb.codebook.weight.grad.numpy().copy()

        # evaluation step (no backward): histogram / loss still become global on demand
        with torch.no_grad():
            dp(zt)
        hist_eval, loss_eval = dp.global_histogram.numpy(), float(dp.global_loss)

        # --- 2. inside DistributedDataParallel, with a small encoder in front: nothing special is needed
        torch.manual_seed(7)
        D = spec["D"]
        enc = torch.nn.Conv2d(D, D, 1)
        cb2 = OracleCodeBook(E)

        class Net(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.enc, self.cb = enc, cb2

            def forward(self, x):
                return self.cb(self.enc(x))

        net = Net()
        ddp = torch.nn.parallel.DistributedDataParallel(net)
        xq, _, l2 = ddp(torch.from_numpy(z[sl].copy()))
        (l2 + (xq * gt).sum()).backward()
        ddp_gw = enc.weight.grad.numpy().copy()
        ddp_gE = cb2.codebook.weight.grad.numpy().copy()

        if rank == 0:
            full = orc.forward(z, E, 0.25)
            gz_full, gE_full = orc.backward(np.transpose(g, (0, 3, 1, 2)), 1.0, z, full["idx"], E, 0.25)
            loc = orc.forward(z[sl], E, 0.25)
            gz_loc, _ = orc.backward(np.transpose(g[sl], (0, 3, 1, 2)), 1.0, z[sl], loc["idx"], E, 0.25)
            # single-device run of the same net on the concatenated batch: objective = global-mean loss + (1/W) sum(z_q g),
            # i.e. the average of the per-rank objectives
            torch.manual_seed(7)
            enc1 = torch.nn.Conv2d(D, D, 1)
            cb1 = OracleCodeBook(E)
            q1, _, l1 = cb1(enc1(torch.from_numpy(z.copy())))
            g_all = torch.from_numpy(np.ascontiguousarray(np.transpose(g, (0, 3, 1, 2))))
            (l1 + (q1 * g_all).sum() / world).backward()
            np.savez(out_file, gE_wrap=gE_wrap, hist_wrap=hist_wrap, loss_wrap=loss_wrap, gz_wrap=zt.grad.numpy() / 3,
                     gE_hook=gE_hook, hist_hook=hist_hook, loss_hook=loss_hook,
                     gE_accum=gE_accum, hist_eval=hist_eval, loss_eval=loss_eval,
                     gE_full=gE_full, hist_full=full["hist"], loss_full=float(full["loss"]), gz_loc=gz_loc,
                     loss_local=float(loss.detach()), loss_loc_ref=float(loc["loss"]),
                     idx_ok=np.array_equal(idx.numpy(), full["idx"][: idx.numel()]),
                     ddp_gw=ddp_gw, ddp_gE=ddp_gE, one_gw=enc1.weight.grad.numpy(), one_gE=cb1.codebook.weight.grad.numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_exchange():
    from parity import assert_close
    with tempfile.TemporaryDirectory() as td:
        init_file = os.path.join(td, "rdzv")
        out_file = os.path.join(td, "out.npz")
        mp.spawn(_worker, args=(2, init_file, out_file), nprocs=2, join=True)
        r = np.load(out_file)
    # stand-alone wrapper: averaged shard gradients == single-device gradient on the concatenated batch
    assert_close(r["gE_wrap"], r["gE_full"], "averaged shard grad_E (overlapped exchange of the per-code sums)")
    assert_close(r["gE_hook"], r["gE_full"], "averaged shard grad_E (post-backward exchange)")
    assert_close(r["gE_accum"], 2 * r["gE_full"], "grad_E after a second accumulated micro-step")
    for tag in ("wrap", "hook", "eval"):
        assert np.array_equal(r[f"hist_{tag}"], r["hist_full"])
        # equal shard sizes: the mean of the per-rank losses is the global loss
        assert abs(float(r[f"loss_{tag}"]) - float(r["loss_full"])) <= 1e-6 * abs(float(r["loss_full"]))
    # each rank's loss and grad_z are the LOCAL-mean quantities (three backward passes were accumulated into zt.grad)
    assert abs(float(r["loss_local"]) - float(r["loss_loc_ref"])) <= 1e-6 * abs(float(r["loss_loc_ref"]))
    assert_close(r["gz_wrap"], r["gz_loc"], "rank-0 grad_z (local mean)")
    assert bool(r["idx_ok"])
    # inside DDP: encoder and codebook gradients equal the single-device ones
    assert_close(r["ddp_gw"], r["one_gw"], "DDP encoder weight gradient", rtol=2e-5)
    assert_close(r["ddp_gE"], r["one_gE"], "DDP codebook gradient", rtol=2e-5)


def test_pack_unpack_exact_counts():
    from vq_vae_gan_diffusion_b200.dist import pack_hist, unpack_hist
    hist = torch.tensor([0, 1, 2 ** 39 + 3, 7, 65535, 65536, 2 ** 24 + 1], dtype=torch.int64)
    buf = torch.zeros(2 * hist.numel())
    pack_hist(hist, buf)
    assert torch.equal(unpack_hist(buf * 3), 3 * hist)       # what a 3-rank SUM of equal histograms delivers
