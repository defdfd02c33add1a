"""Single-GPU A/B of the two ways to get the codebook gradient (cfg4): scatter-add in the backward kernel (default) against
per-code sums accumulated by the forward + one scaling pass (CodeBook.scatter_in_forward, what DataParallelVQ's overlapped
exchange uses).  Prints forward / backward / step milliseconds for both."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vq_vae_gan_diffusion_b200 as vq  # noqa: E402
from bench import WORKLOADS, make_latents  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    wl = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg4"]
    B, H, W, K = wl["B"], wl["H"], wl["W"], wl["K"]
    E, z, g_out = make_latents(torch, dev, B, H, W, K, "init", 1234)
    cb = vq.CodeBook(K, 256).to(dev)
    with torch.no_grad():
        cb.codebook.weight.copy_(E)
    zr = z.clone().requires_grad_(True)
    one = torch.ones((), device=dev)
    out = {}
    for mode in (False, True, False, True):
        cb.scatter_in_forward = mode
        tf, tb = [], []
        for i in range(25):
            cb.codebook.weight.grad = None
            zr.grad = None
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            z_q, idx, loss = cb(zr)
            e[1].record()
            torch.autograd.backward([z_q, loss], [g_out, one])
            e[2].record()
            torch.cuda.synchronize()
            if i >= 5:
                tf.append(e[0].elapsed_time(e[1]))
                tb.append(e[1].elapsed_time(e[2]))
        out.setdefault("scatter_in_forward" if mode else "scatter_in_backward", []).append(
            {"forward_ms": statistics.median(tf), "backward_ms": statistics.median(tb), "step_ms": statistics.median(tf) + statistics.median(tb)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
