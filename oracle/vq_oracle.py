"""Python face of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; nothing under ``vq_vae_gan_diffusion_b200/`` does.

Two restatements of ``/root/reference/network/vqvae/submodule/codebook.py::CodeBook``:

* :class:`COracle` -- ctypes binding of ``oracle/vq_oracle.c`` (canonical accumulation order; the
  checker the CUDA indices are compared against bit for bit).
* :func:`forward_blas` / :func:`backward_blas` -- a numpy line-by-line port whose ``z @ E.T`` runs in
  the BLAS numpy links (OpenBLAS, all host threads).  It has the same cost profile as the
  reference's CPU path (one sgemm + elementwise passes over the (N, K) matrix) and is what the
  benchmark times as ``cpu_baseline`` (``kind: "port"``).

Parity pinning: both are checked against outputs of the reference itself (``tests/golden``, made
by ``tests/golden/make_golden.py`` in the build container, where ``/root/reference`` exists).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libvq_oracle.so")


def build(force: bool = False) -> str:
    """Compile oracle/vq_oracle.c with the committed Makefile; returns the .so path."""
    src = os.path.join(_HERE, "vq_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB


_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)
_u64p = ctypes.POINTER(ctypes.c_uint64)


def _p(a, typ):
    return a.ctypes.data_as(typ) if a is not None else typ()


class COracle:
    """ctypes binding of oracle/vq_oracle.c."""

    def __init__(self):
        self.lib = ctypes.CDLL(build())
        L = self.lib
        L.vq_oracle_num_threads.restype = ctypes.c_int
        L.vq_oracle_row_norms.argtypes = [_f32p, ctypes.c_int64, ctypes.c_int, _f32p]
        L.vq_oracle_row_norms.restype = None
        L.vq_oracle_forward.argtypes = [_f32p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, _f32p, ctypes.c_int,
                                        ctypes.c_float, _f32p, _i64p, _f32p, _i64p, _f32p, _u64p, ctypes.c_int]
        L.vq_oracle_forward.restype = ctypes.c_int
        L.vq_oracle_has_fast_path.restype = ctypes.c_int
        L.vq_oracle_pair_dist.argtypes = [_f32p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, _f32p, _i64p, _i64p,
                                          ctypes.c_int64, _f32p]
        L.vq_oracle_pair_dist.restype = None
        L.vq_oracle_nearest_diffsq.argtypes = [_f32p, ctypes.c_int64, ctypes.c_int, _f32p, ctypes.c_int, _i64p, _f32p, _u64p]
        L.vq_oracle_nearest_diffsq.restype = ctypes.c_int
        L.vq_oracle_nearest_cdist.argtypes = [_f32p, ctypes.c_int64, ctypes.c_int, _f32p, ctypes.c_int, _i64p, _f32p, _u64p, _f32p]
        L.vq_oracle_nearest_cdist.restype = ctypes.c_int
        L.vq_oracle_backward.argtypes = [_f32p, _i64p, ctypes.c_float, _f32p, _i64p, _f32p, ctypes.c_int64,
                                         ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int64,
                                         _f32p, _f32p]
        L.vq_oracle_backward.restype = ctypes.c_int

    @property
    def num_threads(self) -> int:
        return int(self.lib.vq_oracle_num_threads())

    def row_norms(self, X: np.ndarray) -> np.ndarray:
        X = np.ascontiguousarray(X, dtype=np.float32)
        out = np.empty(X.shape[0], np.float32)
        self.lib.vq_oracle_row_norms(_p(X, _f32p), X.shape[0], X.shape[1], _p(out, _f32p))
        return out

    @property
    def has_fast_path(self) -> bool:
        return bool(self.lib.vq_oracle_has_fast_path())

    def forward(self, z: np.ndarray, E: np.ndarray, beta: float = 0.25, want_zq: bool = True, fast: bool = False):
        """z: (B, D, H, W) fp32, E: (K, D) fp32 -> dict(zq_nhwc (N,D), idx, loss, hist, dist_min, tie_rows).

        ``fast`` runs the argmin stage vectorised over codes (AVX2 + FMA, same arithmetic per pair, bit-identical
        results -- tests/test_oracle_golden.py); used for the full-size GPU parity tests."""
        z = np.ascontiguousarray(z, dtype=np.float32)
        E = np.ascontiguousarray(E, dtype=np.float32)
        B, D = z.shape[0], z.shape[1]
        HW = int(np.prod(z.shape[2:])) if z.ndim > 2 else 1
        K = E.shape[0]
        assert E.shape[1] == D
        N = B * HW
        zq = np.empty((N, D), np.float32) if want_zq else None
        idx = np.empty(N, np.int64)
        loss = np.zeros(1, np.float32)
        hist = np.zeros(K, np.int64)
        dmin = np.empty(N, np.float32)
        ties = ctypes.c_uint64(0)
        rc = self.lib.vq_oracle_forward(_p(z, _f32p), B, HW, D, _p(E, _f32p), K, beta, _p(zq, _f32p), _p(idx, _i64p),
                                        _p(loss, _f32p), _p(hist, _i64p), _p(dmin, _f32p), ctypes.byref(ties),
                                        1 if fast else 0)
        if rc != 0:
            raise RuntimeError(f"vq_oracle_forward rc={rc}")
        return dict(zq_nhwc=zq, idx=idx, loss=np.float32(loss[0]), hist=hist, dist_min=dmin, tie_rows=int(ties.value))

    def nearest_diffsq(self, x: np.ndarray, E: np.ndarray):
        """x: (..., D) fp32 rows, E: (K, D) -> dict(idx (N,), dist_min, tie_rows) under the sum((x - e)**2) recipe of
        v_vq_diffusion.py:114-123 (canonical order)."""
        E = np.ascontiguousarray(E, dtype=np.float32)
        rows = np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(-1, E.shape[1]))
        N = rows.shape[0]
        idx = np.empty(N, np.int64)
        dmin = np.empty(N, np.float32)
        ties = ctypes.c_uint64(0)
        rc = self.lib.vq_oracle_nearest_diffsq(_p(rows, _f32p), N, E.shape[1], _p(E, _f32p), E.shape[0], _p(idx, _i64p),
                                               _p(dmin, _f32p), ctypes.byref(ties))
        if rc != 0:
            raise RuntimeError(f"vq_oracle_nearest_diffsq rc={rc}")
        return dict(idx=idx, dist_min=dmin, tie_rows=int(ties.value))

    def nearest_cdist(self, x: np.ndarray, table: np.ndarray):
        """x: (..., D) fp32 rows, table: (K, D) raw -> dict(idx (N,), dist_min, tie_rows, table_hat) under the recipe of
        VQGaussianDiffusion3DWrapper.gaussian_to_indices (diffusion_gaussian3d.py:543-570): both sides L2-normalised, torch.cdist's
        augmented matrix product, clamp_min(0).sqrt(), first minimum (canonical order)."""
        table = np.ascontiguousarray(table, dtype=np.float32)
        rows = np.ascontiguousarray(np.asarray(x, dtype=np.float32).reshape(-1, table.shape[1]))
        N = rows.shape[0]
        idx = np.empty(N, np.int64)
        dmin = np.empty(N, np.float32)
        th = np.empty_like(table)
        ties = ctypes.c_uint64(0)
        rc = self.lib.vq_oracle_nearest_cdist(_p(rows, _f32p), N, table.shape[1], _p(table, _f32p), table.shape[0], _p(idx, _i64p),
                                              _p(dmin, _f32p), ctypes.byref(ties), _p(th, _f32p))
        if rc != 0:
            raise RuntimeError(f"vq_oracle_nearest_cdist rc={rc}")
        return dict(idx=idx, dist_min=dmin, tie_rows=int(ties.value), table_hat=th)

    def pair_dist(self, z: np.ndarray, E: np.ndarray, rows, codes) -> np.ndarray:
        z = np.ascontiguousarray(z, dtype=np.float32)
        E = np.ascontiguousarray(E, dtype=np.float32)
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        codes = np.ascontiguousarray(codes, dtype=np.int64)
        B, D = z.shape[0], z.shape[1]
        HW = int(np.prod(z.shape[2:])) if z.ndim > 2 else 1
        out = np.empty(rows.shape[0], np.float32)
        self.lib.vq_oracle_pair_dist(_p(z, _f32p), B, HW, D, _p(E, _f32p), _p(rows, _i64p), _p(codes, _i64p),
                                     rows.shape[0], _p(out, _f32p))
        return out

    def backward(self, gout, g_loss: float, z: np.ndarray, idx: np.ndarray, E: np.ndarray, beta: float = 0.25,
                 n_global: int = 0):
        """gout: (B, D, H, W) fp32 with any (hw-flattenable) strides, or None -> (grad_z NCHW, grad_E)."""
        z = np.ascontiguousarray(z, dtype=np.float32)
        E = np.ascontiguousarray(E, dtype=np.float32)
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        B, D = z.shape[0], z.shape[1]
        HW = int(np.prod(z.shape[2:])) if z.ndim > 2 else 1
        K = E.shape[0]
        gs = None
        if gout is not None:
            assert gout.dtype == np.float32 and gout.shape == z.shape
            g3 = gout.reshape(B, D, HW) if gout.flags.c_contiguous else None
            if g3 is None:
                # (b, d, hw) element strides of an hw-flattenable view (e.g. NHWC memory viewed NCHW)
                it = gout.itemsize
                sb, sd = gout.strides[0] // it, gout.strides[1] // it
                sw = gout.strides[-1] // it if gout.ndim > 2 else 1
                if gout.ndim == 4:
                    assert gout.strides[2] // it == gout.shape[3] * sw or gout.shape[2] == 1
                gs = np.array([sb, sd, sw], np.int64)
            else:
                gs = np.array([D * HW, HW, 1], np.int64)
        grad_z = np.empty((B, D) + tuple(z.shape[2:]), np.float32)
        grad_E = np.empty((K, D), np.float32)
        rc = self.lib.vq_oracle_backward(_p(gout, _f32p), _p(gs, _i64p), g_loss, _p(z, _f32p), _p(idx, _i64p),
                                         _p(E, _f32p), B, HW, D, K, beta, n_global, _p(grad_z, _f32p),
                                         _p(grad_E, _f32p))
        if rc != 0:
            raise RuntimeError(f"vq_oracle_backward rc={rc}")
        return grad_z, grad_E


# ----------------------------------------------------------------------------------------------
# numpy/BLAS line-by-line port (timed as the CPU baseline)
# ----------------------------------------------------------------------------------------------

def forward_blas(z: np.ndarray, E: np.ndarray, beta: float = 0.25, indices_only: bool = False):
    """Port of codebook.py:62-111 on numpy; the matmul runs in numpy's BLAS (all host threads)."""
    B, D = z.shape[0], z.shape[1]
    zf = np.ascontiguousarray(np.moveaxis(z.reshape(B, D, -1), 1, 2)).reshape(-1, D)   # codebook.py:62-66
    dist = (np.sum(zf ** 2, axis=1, keepdims=True, dtype=np.float32)                  # codebook.py:70-79
            + np.sum(E ** 2, axis=1, dtype=np.float32)
            - np.float32(2) * (zf @ E.T))
    idx = np.argmin(dist, axis=1).astype(np.int64)                                     # codebook.py:82
    if indices_only:
        return None, idx, None
    e = E[idx]                                                                         # codebook.py:85
    diff = e - zf
    m = np.mean(diff ** 2, dtype=np.float32)
    loss = np.float32(np.mean(diff ** 2 + np.float32(beta) * m, dtype=np.float32))     # codebook.py:96-103
    zq = zf + diff                                                                     # codebook.py:106
    return zq, idx, loss


def backward_blas(gout_nhwc, g_loss: float, z: np.ndarray, idx: np.ndarray, E: np.ndarray, beta: float = 0.25):
    """Autograd backward of the port: returns (grad_z as (N, D) rows, grad_E)."""
    B, D = z.shape[0], z.shape[1]
    zf = np.ascontiguousarray(np.moveaxis(z.reshape(B, D, -1), 1, 2)).reshape(-1, D)
    N = zf.shape[0]
    coef = np.float32(2.0 * g_loss / (N * D))
    diff = zf - E[idx]
    grad_z = coef * diff
    if gout_nhwc is not None:
        grad_z += gout_nhwc
    grad_E = np.zeros_like(E)
    np.add.at(grad_E, idx, (-np.float32(beta) * coef) * diff)
    return grad_z, grad_E


# ----------------------------------------------------------------------------------------------
# torch-CPU port (what bench.py times as the CPU baseline): the same ATen op sequence the reference
# issues -- one sgemm, elementwise passes over the (N, K) matrix, argmin, embedding lookup, two means,
# autograd backward -- so that its cost on the host cores is the reference's cost.
# ----------------------------------------------------------------------------------------------

def torch_cpu_step(z, E, g_out=None, beta: float = 0.25, indices_only: bool = False):
    """z (B, D, H, W) and E (K, D) CPU torch tensors.  Returns (z_q NCHW view, idx, loss, grad_z, grad_E);
    the last two are None without g_out.  Restates codebook.py:62-111 + loss.backward()."""
    import torch
    D = E.shape[1]
    need_grad = g_out is not None and not indices_only
    zin = z.detach().clone().requires_grad_(need_grad)
    W = E.detach().clone().requires_grad_(need_grad)
    with torch.set_grad_enabled(need_grad):
        rows = zin.permute(0, 2, 3, 1).contiguous()                  # codebook.py:62
        flat = rows.view(-1, D)                                      # codebook.py:64-66
        dist = (flat ** 2).sum(dim=1, keepdim=True) + (W ** 2).sum(dim=1) - 2 * torch.matmul(flat, W.t())   # :70-79
        idx = torch.argmin(dist, dim=1)                              # codebook.py:82
        if indices_only:
            return None, idx, None, None, None
        q = torch.nn.functional.embedding(idx, W).view(rows.shape)   # codebook.py:85
        loss = torch.mean((q.detach() - rows) ** 2 + beta * torch.mean((q - rows.detach()) ** 2))   # :96-103
        q = rows + (q - rows).detach()                               # codebook.py:106
        z_q = q.permute(0, 3, 1, 2)                                  # codebook.py:109
    if not need_grad:
        return z_q, idx, loss, None, None
    (loss + (z_q * g_out).sum()).backward()
    return z_q.detach(), idx, loss.detach(), zin.grad, W.grad


# ----------------------------------------------------------------------------------------------
# Token-stream formats after the tokeniser; pinned by tests/golden/tok_*.npz
# (tests/golden/make_golden_tokens.py runs the reference's own functions / lines).
# ----------------------------------------------------------------------------------------------

def quant_conv_fp32(h: np.ndarray, W: np.ndarray, b=None) -> np.ndarray:
    """The reference's ``quant_conv`` -- ``nn.Conv2d(C, C, 1)``, /root/reference/network/vqvae/vqvae.py:83, applied at
    vqvae.py:128 -- in fp32 on the CPU: ``z[n, :, hw] = W h[n, :, hw] + bias`` (one sgemm per image, BLAS accumulation order;
    a 1 x 1 convolution IS that matrix product).  h (B, Ci, H, W) fp32, W (Co, Ci) or (Co, Ci, 1, 1) fp32, b (Co) or None."""
    h = np.ascontiguousarray(h, np.float32)
    W2 = np.ascontiguousarray(np.asarray(W, np.float32).reshape(W.shape[0], W.shape[1]))
    B, Ci, H, Wd = h.shape
    z = np.matmul(W2[None], h.reshape(B, Ci, H * Wd))                         # (B, Co, HW) fp32
    if b is not None:
        z = z + np.asarray(b, np.float32)[None, :, None]
    return z.reshape(B, W2.shape[0], H, Wd).astype(np.float32, copy=False)


def index_to_log_onehot_np(x: np.ndarray, num_classes: int) -> np.ndarray:
    """network/vq_diffusion/vq_diffusion.py:29-35 (== diffusion_vq_official.py:53-60): x (B, ...) int64 ->
    log(clamp(one_hot(x), min=1e-30)) as (B, num_classes, ...) fp32, the class axis moved to position 1."""
    x = np.asarray(x, dtype=np.int64)
    if x.min(initial=0) < 0 or x.max(initial=0) >= num_classes:
        raise RuntimeError("Class values must be in [0, num_classes)")                   # F.one_hot raises
    onehot = (x[..., None] == np.arange(num_classes, dtype=np.int64)).astype(np.float32)  # F.one_hot(...).float()  :31/:56
    order = (0, x.ndim) + tuple(range(1, x.ndim))                                         # permute_order  :32/:57
    return np.log(np.maximum(onehot.transpose(order), np.float32(1e-30)))                 # log(clamp(min=1e-30))  :34/:59


def log_onehot_to_index_np(log_x: np.ndarray) -> np.ndarray:
    """network/vq_diffusion/vq_diffusion.py:37-38: log_x.argmax(1) -- first maximal class, a NaN counts as the maximum
    (torch.argmax and numpy.argmax agree on both rules)."""
    return np.argmax(np.asarray(log_x, dtype=np.float32), axis=1).astype(np.int64)


def blend_with_sos_np(indices: np.ndarray, mask: np.ndarray, random_indices: np.ndarray, sos_token: int) -> np.ndarray:
    """network/vqTransformer/vqTransformer.py:117-118, 124, 138-141: mask is the fp32 Bernoulli draw (:121-123), random_indices the
    randint_like draw (:127-129).  Returns cat(sos, mask.round().long() * indices + (1 - mask.round().long()) * random)."""
    m = np.rint(np.asarray(mask, dtype=np.float32)).astype(np.int64)                      # .round().to(int64)  :124
    new = m * np.asarray(indices, np.int64) + (1 - m) * np.asarray(random_indices, np.int64)   # :138
    sos = np.full((new.shape[0], 1), sos_token, dtype=np.int64)                            # :117-118
    return np.concatenate([sos, new], axis=1)                                              # :141
