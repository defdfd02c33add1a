"""Multi-GPU parity check under torchrun on N GPUs: runs the worker of tests/test_gpu_dist.py (batch-sharded CodeBook +
DataParallelVQ over NCCL against the single-device results; CodeBook inside DDP) and prints one line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py
"""
from __future__ import annotations

import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from test_gpu_dist import _check, run_checks
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    res = (run_checks(rank, world, dev), run_checks(rank, world, dev, K=1024, Bl=2, deterministic=True),
           run_checks(rank, world, dev, K=2048, Bl=2, overlap=False),
           run_checks(rank, world, dev, K=4096, Bl=2, overlap=False, collective="multimem"))
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    rc = 0
    if rank == 0:
        try:
            _check(gathered)
            print(f"dist_check world={world}: OK", {k: (f"{v:.2e}" if isinstance(v, float) else v) for k, v in gathered[-1][0].items()})
        except AssertionError as e:
            print(f"dist_check world={world}: FAILED {e}")
            rc = 1
    dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
