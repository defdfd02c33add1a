"""Data-parallel plumbing for the VQ hot path (new in this build; the reference is single-device).

Latents shard by batch across ranks (one process per GPU), the codebook is replicated.  Per training step exactly
one exchange happens, a SUM all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests) of

    grad_E   (K, D) fp32     codebook gradient          -- launched from inside backward
    hist     (K)    int64    usage histogram             -- launched right after forward, overlaps backward
    loss     (1)    fp32     sum of the per-rank losses  -- idem

Everything else is local.  With ``n_global = sum of shard sizes`` passed to the backward, the SUM of the per-rank
codebook gradients equals the single-device gradient on the concatenated batch, and the per-rank ``grad_z`` already
is the corresponding slice of the global-batch gradient; no 1/W rescale is needed.

The gradient all-reduce is launched from a post-accumulate-grad hook on the codebook weight, i.e. as soon as the
scatter-add kernel has been enqueued; NCCL runs it on its own stream, so the rest of the backward pass (quant_conv,
encoder) overlaps it.  ``wait()`` joins it before the optimizer step (stream-ordered, the host does not block).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["pack_stats", "unpack_stats", "allreduce_codebook", "DataParallelVQ"]


def pack_stats(hist: torch.Tensor, loss: torch.Tensor) -> torch.Tensor:
    """[hist (K) | loss | 1] as float64 (counts up to 2^53 stay exact under SUM): one buffer, one collective."""
    K = hist.numel()
    buf = torch.empty(K + 2, dtype=torch.float64, device=hist.device)
    buf[:K] = hist
    buf[K] = loss.detach()
    buf[K + 1] = 1.0
    return buf


def unpack_stats(buf: torch.Tensor):
    """-> (global histogram int64 (K), mean of the per-rank losses fp32 0-dim)."""
    K = buf.numel() - 2
    return buf[:K].round().to(torch.int64), (buf[K] / buf[K + 1]).to(torch.float32)


def allreduce_codebook(grad_E: torch.Tensor, stats_buf: torch.Tensor, group=None, async_op: bool = False):
    """The one exchange step: SUM over ranks, in place.  Returns the work handles when async."""
    w1 = dist.all_reduce(grad_E, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    w2 = dist.all_reduce(stats_buf, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return (w1, w2) if async_op else None


class DataParallelVQ(torch.nn.Module):
    """Wraps a CodeBook for batch-sharded training.

        dp = DataParallelVQ(codebook)             # after dist.init_process_group
        z_q, idx, loss = dp(z_local)              # local forward; histogram / loss all-reduce starts
        (loss + downstream(z_q)).backward()       # grad_E all-reduce starts inside backward, overlapped
        dp.wait()                                 # before optimizer.step(): weight.grad is the global gradient
        dp.global_histogram, dp.global_loss

    Shards must have equal size (the loss reported is the mean of the per-rank losses).
    """

    def __init__(self, codebook, group=None):
        super().__init__()
        self.codebook_module = codebook
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        codebook.grad_world_size = self.world_size
        self._pending = []
        self._hist = None
        self._loss_sum = None
        self.sync_grads = True
        self._hook = codebook.codebook.weight.register_post_accumulate_grad_hook(self._on_grad_ready)

    def _on_grad_ready(self, param):
        if self.world_size > 1 and self.sync_grads:
            self._pending.append(dist.all_reduce(param.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def forward(self, z, **kw):
        out = self.codebook_module(z, **kw)
        z_q, idx, loss = out
        cb = self.codebook_module
        if loss is not None and cb.last_histogram is not None:
            # the module hands out fresh tensors every call, so they can be reduced in place
            self._hist = cb.last_histogram
            self._loss_sum = loss.detach().clone()
            if self.world_size > 1:
                self._pending.append(dist.all_reduce(self._hist, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
                self._pending.append(dist.all_reduce(self._loss_sum, op=dist.ReduceOp.SUM, group=self.group,
                                                     async_op=True))
        return out

    def wait(self):
        """Join the outstanding all-reduces (stream-ordered on GPU: the current stream waits, the host does not)."""
        for w in self._pending:
            if w is not None:
                w.wait()
        self._pending.clear()

    @property
    def global_histogram(self):
        return self._hist

    @property
    def global_loss(self):
        return None if self._loss_sum is None else self._loss_sum / self.world_size

    def no_sync(self):
        """Context manager: skip the gradient all-reduce (gradient-accumulation micro-steps)."""
        outer = self

        class _Ctx:
            def __enter__(self):
                outer.sync_grads = False

            def __exit__(self, *a):
                outer.sync_grads = True

        return _Ctx()
