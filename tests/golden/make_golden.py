"""Generate tests/golden/*.npz by running the UNMODIFIED reference CodeBook (CPU, fp32).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

For every case of tests/cases.py the reference module
(``/root/reference/network/vqvae/submodule/codebook.py::CodeBook``) is instantiated, its weight is
overwritten with the case's codebook, and ``forward`` + ``backward`` are run exactly the way
``worker/vqganVqvaeWorker.py`` does (loss + a downstream use of z_q).  Stored per case:

  idx            int32  reference indices
  loss           fp32   reference loss
  zq / grad_z / grad_E   whole arrays (small cases) or sampled values at tests.cases.sample_positions
  zq_strides, zq_shape   the reference's returned strides/shape (in elements)
  ref_tie_rows   rows whose minimal fp32 distance occurs >= 2 times in the reference's distance matrix
  ref_ne_fp64    rows where the reference argmin differs from the float64 argmin
  torch_version
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))          # tests/
sys.path.insert(0, "/root/reference")

from cases import CASES, FULL_ARRAY_LIMIT, make_inputs, sample_positions  # noqa: E402
from network.vqvae.submodule.codebook import CodeBook  # noqa: E402  (the reference itself)


def run_case(name: str, spec: dict) -> dict:
    z_np, E_np, g_np = make_inputs(spec)
    K, D = E_np.shape
    torch.manual_seed(0)
    ref = CodeBook(num_codebook_vectors=K, latent_dim=D, beta=0.25)
    with torch.no_grad():
        ref.codebook.weight.copy_(torch.from_numpy(E_np))
    z = torch.from_numpy(z_np).clone().requires_grad_(True)
    g_out = torch.from_numpy(g_np).permute(0, 3, 1, 2)     # NHWC memory viewed NCHW, like z_q itself
    z_q, idx, loss = ref(z)
    (loss + (z_q * g_out).sum()).backward()

    # census on the reference's own distance matrix (codebook.py:70-79 verbatim)
    with torch.no_grad():
        zf = z.detach().permute(0, 2, 3, 1).contiguous().view(-1, D)
        W = ref.codebook.weight
        dist = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(W ** 2, dim=1) - 2 * torch.matmul(zf, W.t())
        assert torch.equal(torch.argmin(dist, dim=1), idx)
        mn = dist.min(dim=1, keepdim=True).values
        ref_tie_rows = int(((dist == mn).sum(dim=1) > 1).sum())
        zd, Wd = zf.double(), W.double()
        dist64 = (zd ** 2).sum(1, keepdim=True) + (Wd ** 2).sum(1) - 2 * zd @ Wd.t()
        ref_ne_fp64 = int((dist64.argmin(1) != idx).sum())

    out = dict(
        idx=idx.numpy().astype(np.int32),
        loss=np.float32(loss.item()),
        zq_shape=np.array(z_q.shape, np.int64),
        zq_strides=np.array(z_q.stride(), np.int64),
        ref_tie_rows=np.int64(ref_tie_rows),
        ref_ne_fp64=np.int64(ref_ne_fp64),
        torch_version=np.array(torch.__version__),
    )
    zq_nhwc = z_q.detach().permute(0, 2, 3, 1).contiguous().numpy().reshape(-1, D)
    grad_z = z.grad.numpy()
    grad_E = ref.codebook.weight.grad.numpy()
    N = zq_nhwc.shape[0]
    if N * D <= FULL_ARRAY_LIMIT:
        out.update(zq=zq_nhwc, grad_z=grad_z, grad_E=grad_E)
    else:
        for key, arr in (("zq", zq_nhwc), ("grad_z", grad_z), ("grad_E", grad_E)):
            pos = sample_positions(arr.size, spec["seed"])
            out[key + "_samples"] = arr.reshape(-1)[pos]
            out[key + "_absmax"] = np.float32(np.abs(arr).max())
            out[key + "_sum64"] = np.float64(arr.astype(np.float64).sum())
    return out


def main():
    for name, spec in CASES.items():
        out = run_case(name, spec)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name:16s} N={spec['B']*spec['H']*spec['W']:6d} K={spec['K']:6d} loss={out['loss']:.6g} "
              f"ties={int(out['ref_tie_rows'])} ne_fp64={int(out['ref_ne_fp64'])} "
              f"-> {os.path.getsize(path)/1024:.1f} KiB")


if __name__ == "__main__":
    main()
