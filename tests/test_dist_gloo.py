"""World-size-2 CPU (gloo) tests of the data-parallel path: batch-sharded latents, replicated codebook, ONE exchange
step (SUM all-reduce of grad_E and of [hist | loss | 1]).  The compute on each rank is the CPU oracle (the CUDA
module has no CPU path); what is under test is the host logic of vq_vae_gan_diffusion_b200/dist.py and the
n_global convention of the backward: summed shard gradients == single-device gradient on the concatenated batch.
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
    """CPU stand-in with the CodeBook module's interface, computed by the oracle (test infrastructure only)."""

    def __init__(self, E, beta=0.25):
        super().__init__()
        from oracle.vq_oracle import COracle
        self.oracle = COracle()
        K, D = E.shape
        self.codebook = torch.nn.Embedding(K, D)
        with torch.no_grad():
            self.codebook.weight.copy_(torch.from_numpy(E))
        self.beta = beta
        self.grad_world_size = 1
        self.grad_hook = None
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
                mod.last_histogram = torch.from_numpy(out["hist"])
                idx = torch.from_numpy(out["idx"])
                ctx.mark_non_differentiable(idx)
                zq = torch.from_numpy(out["zq_nhwc"]).reshape(B, H, W, D).permute(0, 3, 1, 2)
                return zq, idx, torch.tensor(float(out["loss"]))

            @staticmethod
            def backward(ctx, g_zq, _gi, g_loss):
                z, w = ctx.saved_tensors
                n_global = ctx.idx.size * mod.grad_world_size
                g = None if g_zq is None else np.ascontiguousarray(g_zq.numpy())
                gz, gE = mod.oracle.backward(g, float(g_loss), z.detach().numpy(), ctx.idx, w.detach().numpy(), mod.beta,
                                             n_global=n_global)
                return torch.from_numpy(gz), torch.from_numpy(gE)

        return Fn.apply(z, self.codebook.weight)


def _worker(rank, world, init_file, out_file):
    from cases import CASES, make_inputs
    from vq_vae_gan_diffusion_b200.dist import DataParallelVQ, allreduce_codebook, pack_stats, unpack_stats
    from oracle.vq_oracle import COracle
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        spec = CASES["small_trained"]                      # B = 2 -> one batch item per rank
        z, E, g = make_inputs(spec)
        B = spec["B"]
        assert B % world == 0
        sl = slice(rank * (B // world), (rank + 1) * (B // world))
        orc = COracle()

        # --- 1. the raw protocol: local oracle results -> one exchange -> global results
        loc = orc.forward(z[sl], E, 0.25)
        n_total = B * spec["H"] * spec["W"]
        gz_loc, gE_loc = orc.backward(np.transpose(g[sl], (0, 3, 1, 2)), 1.0, z[sl], loc["idx"], E, 0.25, n_global=n_total)
        gE_t = torch.from_numpy(gE_loc.copy())
        stats = pack_stats(torch.from_numpy(loc["hist"]), torch.tensor(float(loc["loss"])))
        allreduce_codebook(gE_t, stats)
        hist_g, loss_g = unpack_stats(stats)

        # --- 2. the wrapper: hooks, async all-reduce, wait()
        cb = OracleCodeBook(E)
        dp = DataParallelVQ(cb)
        zt = torch.from_numpy(z[sl].copy()).requires_grad_(True)
        z_q, idx, loss = dp(zt)
        gt = torch.from_numpy(np.ascontiguousarray(np.transpose(g[sl], (0, 3, 1, 2))))
        torch.autograd.backward([z_q, loss], [gt, torch.tensor(1.0)])
        dp.wait()

        if rank == 0:
            full = orc.forward(z, E, 0.25)
            gz_full, gE_full = orc.backward(np.transpose(g, (0, 3, 1, 2)), 1.0, z, full["idx"], E, 0.25)
            np.savez(out_file, gE_proto=gE_t.numpy(), hist_proto=hist_g.numpy(), loss_proto=float(loss_g),
                     gE_wrap=cb.codebook.weight.grad.numpy(), hist_wrap=dp.global_histogram.numpy(),
                     loss_wrap=float(dp.global_loss), gz_wrap=zt.grad.numpy(),
                     gE_full=gE_full, hist_full=full["hist"], loss_full=float(full["loss"]), gz_full=gz_full[sl],
                     idx_ok=np.array_equal(idx.numpy(), full["idx"][: idx.numel()]))
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
    for tag in ("proto", "wrap"):
        assert_close(r[f"gE_{tag}"], r["gE_full"], f"summed shard grad_E ({tag})")
        assert np.array_equal(r[f"hist_{tag}"], r["hist_full"])
        # equal shard sizes: the mean of the per-rank losses is the global loss
        assert abs(float(r[f"loss_{tag}"]) - float(r["loss_full"])) <= 1e-6 * abs(float(r["loss_full"]))
    assert_close(r["gz_wrap"], r["gz_full"], "rank-0 grad_z slice")
    assert bool(r["idx_ok"])


def test_pack_unpack_exact_counts():
    from vq_vae_gan_diffusion_b200.dist import pack_stats, unpack_stats
    hist = torch.tensor([0, 1, 2 ** 40 + 3, 7], dtype=torch.int64)
    buf = pack_stats(hist, torch.tensor(0.5)) + pack_stats(hist, torch.tensor(1.5))
    h, l = unpack_stats(buf)
    assert torch.equal(h, 2 * hist) and abs(float(l) - 1.0) < 1e-7
