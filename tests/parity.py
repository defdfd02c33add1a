"""Parity rules of SURVEY.md section 8(c), used by the oracle tests (CPU) and the GPU parity tests.

Indices: bit-exact, except that a disagreeing row is tolerated -- and COUNTED -- when it is
  * an exact tie: both candidates have the same fp32 distance under the reference formula
    (evaluated in the oracle's canonical order), or
  * inside the rounding band: the float64 distances of the two candidates differ by less than
    4 ulp_fp32(|z|^2 + |e|^2), i.e. the reference's own fp32 rounding decides the winner and a different
    (equally legitimate) sgemm accumulation order can flip it.
Anything else is a real mismatch and fails.

Floats (z_q, loss, grad_z, grad_E): relative error <= 1e-5 (north_star), measured as
max|a-b| / max|b|, plus an elementwise allclose(rtol=1e-5, atol=1e-7*scale).
"""
from __future__ import annotations

import numpy as np

FLOAT_RTOL = 1e-5        # north_star: "z_q, loss and gradients must match within 1e-5 relative error in fp32"


def _rows(z: np.ndarray) -> np.ndarray:
    B, D = z.shape[0], z.shape[1]
    return np.ascontiguousarray(np.moveaxis(z.reshape(B, D, -1), 1, 2)).reshape(-1, D)


def classify_index_mismatches(z, E, idx_a, idx_b, pair_dist=None):
    """Compare two index vectors on inputs (z NCHW, E). -> dict(n, mismatch, tie, band, real, real_rows)."""
    idx_a = np.asarray(idx_a, np.int64).reshape(-1)
    idx_b = np.asarray(idx_b, np.int64).reshape(-1)
    assert idx_a.shape == idx_b.shape
    rows = np.nonzero(idx_a != idx_b)[0]
    res = dict(n=int(idx_a.size), mismatch=int(rows.size), tie=0, band=0, real=0, real_rows=[])
    if rows.size == 0:
        return res
    zf = _rows(z)[rows].astype(np.float64)
    ea = E[idx_a[rows]].astype(np.float64)
    eb = E[idx_b[rows]].astype(np.float64)
    da = ((zf - ea) ** 2).sum(1)
    db = ((zf - eb) ** 2).sum(1)
    mag = (zf ** 2).sum(1) + np.maximum((ea ** 2).sum(1), (eb ** 2).sum(1))
    band = 4.0 * np.spacing(mag.astype(np.float32)).astype(np.float64)
    if pair_dist is not None:      # canonical-order fp32 distances: exact-tie test
        fa = pair_dist(z, E, rows, idx_a[rows])
        fb = pair_dist(z, E, rows, idx_b[rows])
        tie = fa == fb
    else:
        tie = np.zeros(rows.size, bool)
    in_band = np.abs(da - db) < band
    res["tie"] = int(tie.sum())
    res["band"] = int((~tie & in_band).sum())
    real = ~tie & ~in_band
    res["real"] = int(real.sum())
    res["real_rows"] = rows[real][:16].tolist()
    return res


def rel_err(a, b) -> float:
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    scale = np.abs(b).max() if b.size else 0.0
    if scale == 0.0:
        return float(np.abs(a - b).max()) if a.size else 0.0
    return float(np.abs(a - b).max() / scale)


def assert_close(a, b, what: str, rtol: float = FLOAT_RTOL):
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    scale = float(np.abs(b).max()) if b.size else 0.0
    err = rel_err(a, b)
    assert err <= rtol, f"{what}: max-norm relative error {err:.3e} > {rtol:.1e}"
    assert np.allclose(a, b, rtol=rtol, atol=1e-7 * max(scale, 1e-30) + 1e-37), f"{what}: elementwise allclose failed"
    return err
