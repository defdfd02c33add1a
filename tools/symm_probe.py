"""Probe (multi-GPU, torchrun): does torch's symmetric memory work on this box, is NVLS multicast available, and how long do the
library all-reduces take for the CodeBook's exchange buffer (K*D + 2K + 2 fp32 at K = 16384: 16.9 MB) next to NCCL?

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/symm_probe.py
"""
import os
import sys
import time

import torch
import torch.distributed as dist


def timed(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 16384 * 256 + 2 * 16384 + 2
    n = (n + 1023) // 1024 * 1024
    x = torch.randn(n, device=dev)
    out = {"world": world, "n_floats": n}
    out["nccl_ms"] = timed(lambda: dist.all_reduce(x))
    try:
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty(n, dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(t, dist.group.WORLD.group_name)
        out["symm_mem"] = True
        out["multicast"] = bool(hdl.has_multicast_support) if hasattr(hdl, "has_multicast_support") else None
        out["multicast_ptr_nonzero"] = int(hdl.multicast_ptr) != 0
        out["signal_pad_size"] = int(hdl.signal_pad_size)
        t.copy_(x)
        for name in ("two_shot_all_reduce_", "one_shot_all_reduce", "multimem_all_reduce_"):
            try:
                op = getattr(torch.ops.symm_mem, name)
                if name == "one_shot_all_reduce":
                    fn = lambda: op(t, "sum", dist.group.WORLD.group_name)
                else:
                    fn = lambda: op(t, "sum", dist.group.WORLD.group_name)
                out[name + "_ms"] = timed(fn)
            except Exception as e:                           # noqa: BLE001
                out[name + "_err"] = f"{type(e).__name__}: {str(e)[:200]}"
    except Exception as e:                                   # noqa: BLE001
        out["symm_mem"] = False
        out["symm_mem_err"] = f"{type(e).__name__}: {str(e)[:300]}"
    if rank == 0:
        import json
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
