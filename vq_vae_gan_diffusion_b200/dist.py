"""Data-parallel plumbing for the VQ hot path (new in this build; the reference is single-device).

Latents shard by batch across ranks (one process per GPU), the codebook is replicated.  Convention -- the one
``DistributedDataParallel`` uses for every other layer, so a CodeBook works the same inside or outside a DDP-wrapped model:

* every rank's forward returns the mean loss over ITS latents, and its backward returns the gradient of that local mean for
  ``z`` (upstream layers wrapped in DDP get their gradients AVERAGED over ranks, which turns local-mean gradients into
  global-mean gradients);
* the codebook gradient of the global-batch mean is the AVERAGE of the per-rank codebook gradients (equal shard sizes).

Two ways to get that average:

1. The CodeBook sits inside a DDP-wrapped module (the usual case: ``DDP(VQVAE(...))``): nothing to do, DDP's bucket
   all-reduce averages ``codebook.weight.grad`` like any other parameter.  Do NOT also wrap it in :class:`DataParallelVQ`.
2. Stand-alone (bench.py, or a codebook kept out of DDP with ``_ddp_params_and_buffers_to_ignore``):
   :class:`DataParallelVQ` issues exactly ONE collective per training step, a SUM all-reduce (NCCL over NVLink on GPUs,
   gloo in the CPU tests) of one flat fp32 buffer

       [ S or grad_E (K*D) | hist low 16 bits (K) | hist high bits (K) | loss | 1 ]

   * ``overlap=True`` (opt-in): the collective leaves the critical path.  The codebook gradient is linear in the per-code
     sums ``S[k] = sum_{n: idx[n] = k} (e_k - z_n)``, which the FORWARD accumulates (``vq_forward_ex``) straight into the
     head of the flat buffer; the all-reduce starts right after the forward and runs (on NCCL's stream) under whatever
     comes between the forward and the codebook's backward -- decoder, losses, the decoder's backward; in bench.py, where
     nothing comes in between, under the backward kernel itself, which no longer scatters.  The backward waits for it
     (stream-ordered) and turns the summed S into ``weight.grad`` with one scaling pass
     (``vq_backward_ex(code_diff_sum=...)``, scale ``g_loss * beta * 2 / (N D W)``).
   * ``overlap=False`` (default), and automatically for gradient-accumulation steps: the backward writes ``grad_E`` pre-scaled by 1/W
     into the head of the buffer (no packing copy) and a post-accumulate-grad hook starts the all-reduce as soon as the
     scatter-add kernel has been enqueued; ``wait()`` joins it before the optimizer step.

   Histogram counts travel as two exact fp32 words (low 16 bits and the rest: sums stay below 2^24 for up to 256 ranks and
   2^40 latents per code).  ``wait()`` is stream-ordered, the host does not block.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["DataParallelVQ", "pack_hist", "unpack_hist"]


def pack_hist(hist: torch.Tensor, out: torch.Tensor) -> None:
    """int64 counts (K) -> out (2K) fp32: [count & 0xffff | count >> 16], each exactly representable and exactly summable."""
    K = hist.numel()
    out[:K] = (hist & 0xFFFF).to(torch.float32)
    out[K:2 * K] = (hist >> 16).to(torch.float32)


def unpack_hist(buf: torch.Tensor) -> torch.Tensor:
    """(2K) fp32 sums -> int64 counts (K)."""
    K = buf.numel() // 2
    return buf[:K].round().to(torch.int64) + (buf[K:2 * K].round().to(torch.int64) << 16)


class DataParallelVQ(torch.nn.Module):
    """Wraps a stand-alone CodeBook for batch-sharded training (see the module docstring for when NOT to use it).

        dp = DataParallelVQ(codebook)             # after dist.init_process_group
        z_q, idx, loss = dp(z_local)              # local forward (overlap=True: the ONE all-reduce of the step starts here ...
        (loss + downstream(z_q)).backward()       # ... and is joined inside the codebook's backward; else it starts here)
        dp.wait()                                 # before optimizer.step(): weight.grad is the global-batch gradient
        dp.global_histogram, dp.global_loss

    Shards must have equal size (the global loss / gradient are means of the per-rank ones).
    """

    def __init__(self, codebook, group=None, overlap: bool = False, collective: str = "nccl"):
        super().__init__()
        self.codebook_module = codebook
        self.group = group
        self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        # collective = "multimem": the library's own NVLS all-reduce kernel (csrc/vq_allreduce.cuh: multimem.ld_reduce / multimem.st
        # on a symmetric buffer, the NVSwitch adds) instead of NCCL's, in-stream; "nccl" (default) = torch.distributed.all_reduce.
        # "auto": multimem when every rank can set it up (CUDA ranks behind an NVSwitch with multicast support), else nccl.
        if collective not in ("nccl", "multimem", "auto"):
            raise ValueError("collective must be 'nccl', 'multimem' or 'auto'")
        self._symm = None            # (persistent symmetric buffer, handle, padded length) of the multimem path
        self._symm_sync = None       # local_sync words of vq_allreduce_multimem
        if collective == "auto":
            collective = self._probe_multimem(codebook)
        self.collective = collective
        if collective == "multimem":
            overlap = False          # the NVLS kernel runs in-stream on one persistent buffer: nothing to overlap with
        # The overlapped exchange is opt-in: accumulating the sums in the forward costs ~50-70 us per cfg4 step on the rank itself
        # (the atomics are free inside the HBM-bound backward kernel but not inside the latency-bound select kernel:
        # profiles/r2_ab_scatter_in_forward_1gpu.json), about what the hidden all-reduce of a 17 MB buffer takes on 2-8 GPUs -- in
        # bench.py, where nothing runs between forward and backward, it measured neutral (profiles/r2_ab_overlapped_exchange.jsonl).
        # It pays off when the collective is slower (more ranks, larger codebooks, a slower fabric).
        self.overlap = bool(overlap)
        codebook.grad_scale = 1.0 / self.world_size
        codebook.grad_alloc = self._alloc_grad
        codebook.scatter_alloc = self._alloc_scatter
        codebook.scatter_ready = self._scatter_ready
        self._flat = None            # the step's exchange buffer
        self._flat_has_stats = False
        self._work = None            # outstanding collective of the post-backward path: (work, flat, K, D, copy_back)
        self._fwd_work = None        # outstanding collective of the overlapped path: (work, flat, K, D)
        self._step_overlapped = False
        self._reduced = None         # (flat buffer, K, D) of the last completed exchange
        self._local = None           # (hist, loss) of the last forward, not yet exchanged
        self.sync_grads = True
        self._hook = codebook.codebook.weight.register_post_accumulate_grad_hook(self._on_grad_ready)

    def _probe_multimem(self, codebook) -> str:
        """collective="auto": set up the symmetric exchange buffer now (a rendezvous: every rank constructs the wrapper at the
        same point) and agree across ranks on whether it worked; any rank's failure sends everybody to NCCL."""
        w = codebook.codebook.weight
        if self.world_size <= 1 or not w.is_cuda or dist.get_backend(self.group) != "nccl":
            return "nccl"
        self.collective = "multimem"
        ok = 1
        try:
            self._symm_flat(w.shape[0], w.shape[1], w.device)
        except Exception:                                      # noqa: BLE001 -- no symmetric memory / no multicast on this box
            ok = 0
            self._symm = None
        flag = torch.tensor([ok], dtype=torch.int32, device=w.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 1:
            return "multimem"
        self._symm = None
        return "nccl"

    # ---- flat exchange buffer: [S or grad_E (K*D) | hist lo (K) | hist hi (K) | loss | 1]
    def _new_flat(self, K, D, device):
        if self.collective == "multimem" and self.world_size > 1 and device.type == "cuda":
            return self._symm_flat(K, D, device)
        return torch.empty(K * D + 2 * K + 2, dtype=torch.float32, device=device)

    def _symm_flat(self, K, D, device):
        """The multimem path exchanges through ONE persistent symmetric-memory buffer (allocation + handle exchange are a
        rendezvous, far too slow per step): every step reuses it, so what leaves this class is copied out of it."""
        n = K * D + 2 * K + 2
        n_pad = -(-n // (4 * self.world_size)) * (4 * self.world_size)
        if self._symm is None or self._symm[2] != n_pad or self._symm[0].device != device:
            import torch.distributed._symmetric_memory as symm_mem
            buf = symm_mem.empty(n_pad, dtype=torch.float32, device=device)
            grp = self.group if self.group is not None else dist.group.WORLD
            hdl = symm_mem.rendezvous(buf, grp.group_name)
            if int(hdl.multicast_ptr) == 0:
                raise RuntimeError("collective='multimem' needs NVLS multicast support (NVSwitch); use collective='nccl'")
            buf.zero_()
            self._symm = (buf, hdl, n_pad)
            self._symm_sync = torch.zeros(2, dtype=torch.int32, device=device)    # local_sync words of vq_allreduce_multimem
        return self._symm[0][:n]

    def _all_reduce(self, flat):
        """SUM over ranks, in place; returns a work handle to wait on, or None when the collective ran in-stream."""
        if self.collective == "multimem" and self._symm is not None and flat.data_ptr() == self._symm[0].data_ptr():
            from . import _native
            buf, hdl, n_pad = self._symm
            with torch.cuda.device(buf.device):
                rc = _native.lib().vq_allreduce_multimem(int(hdl.multicast_ptr), int(hdl.signal_pad_ptrs_dev), int(hdl.rank),
                                                         int(hdl.world_size), n_pad, self._symm_sync.data_ptr(),
                                                         int(torch.cuda.current_stream(buf.device).cuda_stream))
            _native.check(rc, "vq_allreduce_multimem")
            return None
        return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _alloc_scatter(self, K, D, device):
        """Called by the CodeBook's forward (overlapped path): the per-code sums go into the head of a fresh flat buffer."""
        self._flat = self._new_flat(K, D, device)
        self._flat_has_stats = False
        return self._flat[:K * D].view(K, D)

    def _scatter_ready(self, scat):
        """Called by the CodeBook's backward before it reads the sums: join their all-reduce (the stream waits, not the host)."""
        fw = self._fwd_work
        if fw is not None and fw[1].data_ptr() == scat.data_ptr():
            if fw[0] is not None:
                fw[0].wait()
            self._reduced = (fw[1], fw[2], fw[3])
            self._fwd_work = None

    def _alloc_grad(self, K, D, device):
        """Called by the CodeBook's backward (post-backward path): grad_E is the head of the step's flat buffer (fresh per step,
        because autograd may keep the tensor as ``weight.grad``; allocated -- and its histogram / loss tail filled -- right
        after the forward, so that only the collective itself is left to do when the gradient arrives)."""
        if self._flat is None or self._flat.numel() != K * D + 2 * K + 2 or self._flat.device != device:
            self._flat = self._new_flat(K, D, device)
            self._flat_has_stats = False
        return self._flat[:K * D].view(K, D)

    def _fill_stats(self, flat, K, D):
        hist, loss = self._local
        tail = flat[K * D:]
        if flat.is_cuda and hist.dtype == torch.int64 and hist.is_contiguous() and loss.dtype == torch.float32:
            # one launch (vq_pack_stats) instead of eight small tensor kernels between the forward and the backward
            from . import _native
            from .codebook import _on_device, _stream_ptr
            with _on_device(flat.device):
                rc = _native.lib().vq_pack_stats(hist.data_ptr(), loss.data_ptr(), K, tail.data_ptr(), _stream_ptr(flat.device))
            _native.check(rc, "vq_pack_stats")
            return
        pack_hist(hist, tail)
        tail[2 * K:2 * K + 1].copy_(loss.reshape(1))
        tail[2 * K + 1:].fill_(1.0)

    def _on_grad_ready(self, param):
        if self._step_overlapped or self.world_size <= 1 or not self.sync_grads or self._local is None:
            return
        K, D = param.shape
        flat = self._flat
        aliased = flat is not None and param.grad is not None and param.grad.data_ptr() == flat.data_ptr()
        if not aliased:
            # autograd accumulated into an existing .grad (gradient accumulation) or copied: exchange a packed copy
            flat = self._new_flat(K, D, param.device)
            flat[:K * D].view(K, D).copy_(param.grad)
            self._flat_has_stats = False
        if not self._flat_has_stats:
            self._fill_stats(flat, K, D)
        work = self._all_reduce(flat)
        if self._symm is not None and flat.data_ptr() == self._symm[0].data_ptr():
            # the persistent symmetric buffer is reused next step: weight.grad and the statistics get their own copies
            flat = flat.clone()
            param.grad = flat[:K * D].view(K, D)
            aliased = True
        self._work = (work, flat, K, D, None if aliased else param)
        self._local = None
        self._flat = None
        self._flat_has_stats = False

    def forward(self, z, **kw):
        cb = self.codebook_module
        w = cb.codebook.weight
        training = self.world_size > 1 and self.sync_grads and torch.is_grad_enabled() and w.requires_grad and w.dim() == 2
        # the overlapped exchange needs this step's gradient to be the whole gradient (no accumulation in progress)
        self._step_overlapped = bool(training and self.overlap and w.grad is None and not cb.deterministic
                                     and not kw.get("indices_only", False))
        cb.scatter_in_forward = self._step_overlapped
        try:
            out = cb(z, **kw)
        finally:
            cb.scatter_in_forward = False
        z_q, idx, loss = out
        if loss is not None and cb.last_histogram is not None:
            self._local = (cb.last_histogram, loss.detach())
            self._reduced = None
            if training:
                K, D = w.shape
                if self._step_overlapped and self._flat is not None:
                    # the forward filled the head with the per-code sums: complete the buffer and start the step's collective
                    # (a side stream for the packing kernels + the launch was tried: record_stream on the 17 MB buffer makes the
                    # caching allocator fall back to fresh allocations every step, 1.7 ms per step at 2 GPUs)
                    flat = self._flat
                    self._fill_stats(flat, K, D)
                    work = self._all_reduce(flat)
                    self._fwd_work = (work, flat, K, D)
                    self._local = None
                    self._flat = None
                else:
                    self._step_overlapped = False
                    # post-backward path: the step's exchange buffer, its histogram / loss tail filled now (these small
                    # kernels run ahead of the backward pass instead of between the scatter-add and the collective)
                    self._flat = self._new_flat(K, D, w.device)
                    self._fill_stats(self._flat, K, D)
                    self._flat_has_stats = True
        return out

    def wait(self):
        """Join the outstanding all-reduce (stream-ordered on GPU: the current stream waits, the host does not)."""
        if self._work is not None:
            work, flat, K, D, copy_back = self._work
            if work is not None:
                work.wait()
            if copy_back is not None and copy_back.grad is not None:
                copy_back.grad.copy_(flat[:K * D].view(K, D))
            self._reduced = (flat, K, D)
            self._work = None
        if self._fwd_work is not None:                        # a forward whose backward never ran (or has not run yet)
            work, flat, K, D = self._fwd_work
            if work is not None:
                work.wait()
            self._reduced = (flat, K, D)
            self._fwd_work = None

    def _ensure_stats(self):
        """Histogram / loss exchange for steps without a backward (evaluation): a small collective of its own."""
        self.wait()
        if self._reduced is None and self._local is not None:
            hist, loss = self._local
            K = hist.numel()
            flat = torch.zeros(2 * K + 2, dtype=torch.float32, device=hist.device)
            pack_hist(hist, flat)
            flat[2 * K] = loss
            flat[2 * K + 1] = 1.0
            if self.world_size > 1:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            self._reduced = (flat, K, 0)
            self._local = None

    @property
    def global_histogram(self):
        self._ensure_stats()
        if self._reduced is None:
            return None
        flat, K, D = self._reduced
        return unpack_hist(flat[K * D:K * D + 2 * K])

    @property
    def global_loss(self):
        self._ensure_stats()
        if self._reduced is None:
            return None
        flat, K, D = self._reduced
        return flat[K * D + 2 * K] / flat[K * D + 2 * K + 1]

    def no_sync(self):
        """Context manager: skip the gradient all-reduce (gradient-accumulation micro-steps)."""
        outer = self

        class _Ctx:
            def __enter__(self):
                outer.sync_grads = False

            def __exit__(self, *a):
                outer.sync_grads = True

        return _Ctx()
