"""Pin the CPU oracle (oracle/vq_oracle.c and the numpy/BLAS port) against outputs of the reference itself.

The golden files were produced by tests/golden/make_golden.py from the unmodified reference CodeBook
(/root/reference/network/vqvae/submodule/codebook.py).  CPU-only; runs in seconds.
"""
import os

import numpy as np
import pytest

from cases import CASES, FULL_ARRAY_LIMIT, make_inputs, sample_positions
from parity import assert_close, classify_index_mismatches

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def compare_floats(gold, key, arr, spec, mask_rows=None, D=None):
    """arr against the golden whole array or its samples; rows in mask_rows (index disagreements) are skipped."""
    arr = np.asarray(arr)
    if key in gold.files:
        ref = gold[key].reshape(arr.shape)
        if mask_rows is not None and len(mask_rows):
            keep = np.ones(arr.shape[0], bool)
            keep[mask_rows] = False
            arr, ref = arr[keep], ref[keep]
        return assert_close(arr, ref, key)
    pos = sample_positions(arr.size, spec["seed"])
    got = arr.reshape(-1)[pos]
    ref = gold[key + "_samples"]
    scale = float(gold[key + "_absmax"])
    if mask_rows is not None and len(mask_rows):
        keep = ~np.isin(pos // D, mask_rows)
        got, ref = got[keep], ref[keep]
    err = float(np.abs(got.astype(np.float64) - ref).max() / max(scale, 1e-30))
    assert err <= 1e-5, f"{key}: sampled relative error {err:.3e}"
    return err


@pytest.mark.parametrize("name", list(CASES))
def test_c_oracle_matches_reference(name, oracle):
    spec = CASES[name]
    gold = load(name)
    z, E, g = make_inputs(spec)
    D = spec["D"]
    out = oracle.forward(z, E, beta=0.25)
    ref_idx = gold["idx"].astype(np.int64)

    cls = classify_index_mismatches(z, E, out["idx"], ref_idx, pair_dist=oracle.pair_dist)
    assert cls["real"] == 0, f"real index mismatches vs reference: {cls}"
    # disagreements can only come from rows the reference itself cannot decide exactly
    assert cls["mismatch"] <= int(gold["ref_tie_rows"]) + int(gold["ref_ne_fp64"]) + 2, cls
    bad = np.nonzero(out["idx"] != ref_idx)[0]

    assert np.array_equal(out["hist"], np.bincount(out["idx"], minlength=spec["K"]))
    # loss: every row contributes, mismatching rows are near-ties so the value agrees to ~1e-7
    assert abs(float(out["loss"]) - float(gold["loss"])) <= 1e-5 * max(abs(float(gold["loss"])), 1e-30) + 1e-12
    compare_floats(gold, "zq", out["zq_nhwc"], spec, bad, D)

    g_view = np.transpose(g, (0, 3, 1, 2))                       # NHWC memory viewed NCHW (like the reference test)
    grad_z, grad_E = oracle.backward(g_view, 1.0, z, ref_idx, E, beta=0.25)   # reference indices: float parity
    B, H, W = spec["B"], spec["H"], spec["W"]
    gz_rows = np.moveaxis(grad_z.reshape(B, D, H * W), 1, 2).reshape(-1, D)
    if "grad_z" in gold.files:
        assert_close(grad_z, gold["grad_z"], "grad_z")
        assert_close(grad_E, gold["grad_E"], "grad_E")
    else:
        compare_floats(gold, "grad_z", grad_z, spec)
        compare_floats(gold, "grad_E", grad_E, spec)
    assert gz_rows.shape == (B * H * W, D)


def test_known_answer_k4_d2(oracle):
    """Hand-computed: codes on the unit square corners."""
    spec = CASES["k4_d2"]
    z, E, g = make_inputs(spec)
    out = oracle.forward(z, E, beta=0.25)
    assert out["idx"].tolist() == [0, 1, 3]
    # sum (e-z)^2 = (0.01+0.04) + (0.01+0.04) + (0.16+0.01) = 0.27 ; mean over 6 = 0.045 ; *(1+0.25)
    assert abs(float(out["loss"]) - 0.05625) < 1e-7
    assert out["hist"].tolist() == [1, 1, 0, 1]
    gold = load("k4_d2")
    assert gold["idx"].tolist() == [0, 1, 3] and abs(float(gold["loss"]) - 0.05625) < 1e-7


def test_duplicate_rows_lowest_index(oracle):
    spec = CASES["dup_rows"]
    z, E, _ = make_inputs(spec)
    out = oracle.forward(z, E)
    assert out["idx"].max() < spec["K"] // 3          # the first copy of every duplicated row wins
    assert out["tie_rows"] == out["idx"].size
    assert np.array_equal(out["idx"], load("dup_rows")["idx"])


def test_exact_hit_zero_distance(oracle):
    spec = CASES["exact_hit"]
    z, E, _ = make_inputs(spec)
    out = oracle.forward(z, E)
    assert float(out["loss"]) == 0.0
    rows = np.moveaxis(z.reshape(spec["B"], spec["D"], -1), 1, 2).reshape(-1, spec["D"])
    assert np.array_equal(out["zq_nhwc"], rows)
    assert np.array_equal(E[out["idx"]], rows)


def test_zero_codebook_all_tie_index0(oracle):
    spec = CASES["zero_codebook"]
    z, E, _ = make_inputs(spec)
    out = oracle.forward(z, E)
    assert (out["idx"] == 0).all() and out["tie_rows"] == out["idx"].size
    assert (load("zero_codebook")["idx"] == 0).all()


@pytest.mark.parametrize("name", ["small_init", "cfg1_init", "cfg2s_trained", "k8192_init"])
def test_blas_port_matches_reference(name, oracle):
    """The numpy/BLAS port (the timed CPU baseline) follows the reference too."""
    from oracle.vq_oracle import backward_blas, forward_blas
    spec = CASES[name]
    gold = load(name)
    z, E, g = make_inputs(spec)
    zq, idx, loss = forward_blas(z, E, 0.25)
    cls = classify_index_mismatches(z, E, idx, gold["idx"].astype(np.int64), pair_dist=oracle.pair_dist)
    assert cls["real"] == 0, cls
    assert abs(float(loss) - float(gold["loss"])) <= 1e-5 * abs(float(gold["loss"]))
    _, idx_only, _ = forward_blas(z, E, 0.25, indices_only=True)
    assert np.array_equal(idx, idx_only)
    gz, gE = backward_blas(g.reshape(-1, spec["D"]), 1.0, z, gold["idx"].astype(np.int64), E, 0.25)
    gz_c, gE_c = oracle.backward(np.transpose(g, (0, 3, 1, 2)), 1.0, z, gold["idx"].astype(np.int64), E, 0.25)
    B, D = spec["B"], spec["D"]
    assert_close(gz, np.moveaxis(gz_c.reshape(B, D, -1), 1, 2).reshape(-1, D), "grad_z port vs C")
    assert_close(gE, gE_c, "grad_E port vs C")


def test_oracle_backward_n_global(oracle):
    """Sharded backward with n_global equals the single-device gradient on the concatenated batch."""
    spec = CASES["small_trained"]
    z, E, g = make_inputs(spec)
    out = oracle.forward(z, E)
    gv = np.transpose(g, (0, 3, 1, 2))
    gz, gE = oracle.backward(gv, 1.0, z, out["idx"], E)
    N = out["idx"].size
    half = spec["B"] // 2
    hw = spec["H"] * spec["W"]
    parts = []
    gE_sum = np.zeros_like(gE)
    for sl in (slice(0, half), slice(half, spec["B"])):
        rows = slice(sl.start * hw, sl.stop * hw)
        gzi, gEi = oracle.backward(gv[sl], 1.0, z[sl], out["idx"][rows], E, n_global=N)
        parts.append(gzi)
        gE_sum += gEi
    assert_close(np.concatenate(parts, 0), gz, "sharded grad_z")
    assert_close(gE_sum, gE, "sharded grad_E")


# ------------------------------------------------------------------------------------------------------------------
# Row-major nearest-code search: the same oracle, fed (N, D) rows as an (N, D, 1, 1) grid, against the
# outputs of the reference's GaussianDiffusion2D.gaussian_to_indices (tests/golden/make_golden_nn.py).
from cases import NN_CASES, make_nn_inputs  # noqa: E402


def _rows_as_grid(x):
    rows = x.reshape(-1, x.shape[-1])
    return np.ascontiguousarray(rows.reshape(rows.shape[0], rows.shape[1], 1, 1))


@pytest.mark.parametrize("name", sorted(NN_CASES))
def test_oracle_matches_reference_nearest_rows(name, oracle):
    spec = NN_CASES[name]
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    x, table = make_nn_inputs(spec)
    z = _rows_as_grid(x)
    ref = oracle.forward(z, table, want_zq=False)
    assert tuple(gold["idx_shape"]) == (spec["B"], spec["L"])
    cls = classify_index_mismatches(z, table, ref["idx"], gold["idx"].reshape(-1).astype(np.int64), pair_dist=oracle.pair_dist)
    assert cls["real"] == 0, cls
    assert cls["mismatch"] <= int(gold["ref_tie_rows"]) + int(gold["ref_ne_fp64"]) + 2, cls


@pytest.mark.parametrize("name", ["nn_g96_clean", "nn_g96_noisy"])
def test_zero_padding_changes_no_distance(name, oracle):
    """The CUDA path runs 96-wide tables zero-padded to 256 columns: in the canonical order a zero column adds
    fma(0, 0, p) == p, so indices, ties and minimal distances are bit-identical."""
    spec = NN_CASES[name]
    x, table = make_nn_inputs(spec)
    z = _rows_as_grid(x)
    a = oracle.forward(z, table, want_zq=False)
    zp = np.zeros((z.shape[0], 256, 1, 1), np.float32)
    zp[:, : z.shape[1]] = z
    tp = np.zeros((table.shape[0], 256), np.float32)
    tp[:, : table.shape[1]] = table
    b = oracle.forward(zp, tp, want_zq=False)
    assert np.array_equal(a["idx"], b["idx"]) and a["tie_rows"] == b["tie_rows"]
    assert np.array_equal(a["dist_min"], b["dist_min"])


def test_nan_and_inf_follow_torch_argmin(oracle):
    """torch.argmin treats a NaN distance as the minimum and returns the first one (codebook.py:82); a row with +-inf
    gets inf / NaN distances depending on the sign of the code's element there.  The oracle must reproduce exactly
    that -- checked against the torch-CPU port of the reference on the same inputs (no BLAS ambiguity: every affected
    distance is inf or NaN)."""
    import torch
    from oracle.vq_oracle import torch_cpu_step
    rng = np.random.default_rng(21)
    K, B, H, W = 70, 2, 4, 8
    E = rng.standard_normal((K, 256)).astype(np.float32)
    z = rng.standard_normal((B, 256, H, W)).astype(np.float32)
    z[0, 5, 1, 3] = np.nan
    z[1, 200, 3, 7] = np.inf
    z[1, 17, 0, 0] = -np.inf
    ref = oracle.forward(z, E, want_zq=False)
    idx_t = torch_cpu_step(torch.from_numpy(z), torch.from_numpy(E), None, 0.25, indices_only=True)[1].numpy()
    bad = [0 * H * W + 1 * W + 3, 1 * H * W + 3 * W + 7, 1 * H * W + 0]
    for r in bad:
        assert ref["idx"][r] == idx_t[r], (r, ref["idx"][r], idx_t[r])
    assert ref["idx"][bad[0]] == 0                                   # all-NaN row: first code
    assert ref["idx"][bad[1]] == int(np.argmax(E[:, 200] > 0))       # +inf: first code whose element there is positive
    assert ref["idx"][bad[2]] == int(np.argmax(E[:, 17] < 0))        # -inf: first code whose element there is negative
    ok = np.ones(B * H * W, bool)
    ok[bad] = False
    cls = classify_index_mismatches(z, E, np.where(ok, ref["idx"], 0), np.where(ok, idx_t, 0), pair_dist=oracle.pair_dist)
    assert cls["real"] == 0
    # a NaN inside the codebook makes that code every row's argmin
    E2 = E.copy()
    E2[33, 7] = np.nan
    assert (oracle.forward(z, E2, want_zq=False)["idx"][ok] == 33).all()


# Broadcast-difference recipe (V_VQDiffusion.sample, v_vq_diffusion.py:114-123)
from cases import DIFFSQ_CASES, make_diffsq_inputs  # noqa: E402


def _classify_diffsq(x, E, idx_a, idx_b):
    """Mismatching rows: exact tie in float64 distance terms or inside the rounding band of a 256-term fp32 sum."""
    rows = x.reshape(-1, x.shape[-1]).astype(np.float64)
    bad = np.nonzero(idx_a != idx_b)[0]
    real = 0
    for r in bad:
        da = ((rows[r] - E[idx_a[r]].astype(np.float64)) ** 2).sum()
        db = ((rows[r] - E[idx_b[r]].astype(np.float64)) ** 2).sum()
        if abs(da - db) > 70 * 2.0 ** -24 * max(da, db):
            real += 1
    return dict(mismatch=int(bad.size), real=real)


@pytest.mark.parametrize("name", sorted(DIFFSQ_CASES))
def test_oracle_matches_reference_diffsq(name, oracle):
    spec = DIFFSQ_CASES[name]
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    x, E = make_diffsq_inputs(spec)
    ref = oracle.nearest_diffsq(x, E)
    cls = _classify_diffsq(x, E, ref["idx"], gold["idx"].reshape(-1).astype(np.int64))
    assert cls["real"] == 0, cls
    assert cls["mismatch"] <= int(gold["ref_tie_rows"]) + int(gold["ref_ne_fp64"]) + 2, cls


def test_diffsq_known_answers(oracle):
    E = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0], [1.0, 1.0]], np.float32)
    x = np.array([[0.1, 0.2], [0.9, 0.2], [0.6, 0.9], [1.0, 1.0]], np.float32)
    out = oracle.nearest_diffsq(x, E)
    assert out["idx"].tolist() == [0, 1, 3, 3]                      # duplicate codes 3 and 4: the lower index wins
    assert out["tie_rows"] == 2 and out["dist_min"][3] == 0.0
    xn = x.copy()
    xn[1, 0] = np.nan
    assert oracle.nearest_diffsq(xn, E)["idx"].tolist() == [0, 0, 3, 3]   # all distances NaN -> first code


# ------------------------------------------------------------------ token-stream formats

def test_token_onehot_oracle_matches_reference():
    """oracle index_to_log_onehot_np == the reference's index_to_log_onehot (both copies), bit for bit."""
    from cases import TOKEN_ONEHOT_CASES, make_token_indices
    from oracle.vq_oracle import index_to_log_onehot_np
    for name, spec in TOKEN_ONEHOT_CASES.items():
        gold = load(name)["out"]
        got = index_to_log_onehot_np(make_token_indices(spec), spec["num_classes"])
        assert got.dtype == np.float32 and got.shape == gold.shape, name
        assert np.array_equal(got, gold), name
    with pytest.raises(RuntimeError):
        index_to_log_onehot_np(np.array([[0, 5]]), 5)


def test_token_blend_oracle_matches_reference():
    """oracle blend_with_sos_np == vqTransformer.py:117-141 evaluated on the stored draws."""
    from cases import TOKEN_BLEND_CASES, make_token_indices
    from oracle.vq_oracle import blend_with_sos_np
    for name, spec in TOKEN_BLEND_CASES.items():
        gold = load(name)
        got = blend_with_sos_np(make_token_indices(spec), gold["mask"], gold["random_indices"], spec["sos_token"])
        assert got.dtype == np.int64 and np.array_equal(got, gold["new_indices"]), name
        assert (got[:, 0] == spec["sos_token"]).all()


def test_token_oracles_against_torch_cpu_ops():
    """Random shapes: the numpy restatements against the reference's op sequence evaluated with torch on CPU."""
    import torch
    from oracle.vq_oracle import blend_with_sos_np, index_to_log_onehot_np
    rng = np.random.default_rng(9)
    for _ in range(12):
        nd = int(rng.integers(1, 4))
        shape = tuple(int(v) for v in rng.integers(1, 9, size=nd))
        C = int(rng.integers(1, 40))
        x = rng.integers(0, C, size=shape, dtype=np.int64)
        xt = torch.from_numpy(x)
        onehot = torch.nn.functional.one_hot(xt, C).permute((0, -1) + tuple(range(1, xt.dim())))      # vq_diffusion.py:31-33
        exp = torch.log(onehot.float().clamp(min=1e-30)).numpy()                                       # :34
        assert np.array_equal(index_to_log_onehot_np(x, C), exp)
    for _ in range(6):
        B, L, V = int(rng.integers(1, 6)), int(rng.integers(1, 50)), int(rng.integers(2, 2000))
        idx = torch.from_numpy(rng.integers(0, V, size=(B, L), dtype=np.int64))
        mask_f = torch.bernoulli(0.5 * torch.ones(B, L))
        rnd = torch.randint_like(idx, high=V)
        m = mask_f.round().to(dtype=torch.int64)
        exp = torch.cat(((torch.ones(B, 1) * 7).long(), m * idx + (1 - m) * rnd), dim=1)              # vqTransformer.py:117-141
        assert np.array_equal(blend_with_sos_np(idx.numpy(), mask_f.numpy(), rnd.numpy(), 7), exp.numpy())


# ------------------------------------------------------------------------------------------------------------------
# The vectorised argmin stage of the oracle (used by the full-size GPU parity tests) against its scalar definition.
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(CASES))
def test_fast_oracle_stage_is_bit_identical_to_the_scalar_one(name, oracle):
    if not oracle.has_fast_path:
        pytest.skip("host has no AVX2 + FMA: forward(fast=True) runs the scalar stage")
    z, E, _ = make_inputs(CASES[name])
    a, b = oracle.forward(z, E), oracle.forward(z, E, fast=True)
    for key in ("idx", "hist", "zq_nhwc", "dist_min"):
        assert np.array_equal(a[key], b[key], equal_nan=key == "dist_min"), key
    assert a["tie_rows"] == b["tie_rows"]
    assert a["loss"] == b["loss"] or (np.isnan(a["loss"]) and np.isnan(b["loss"]))


def test_fast_oracle_stage_nan_inf_and_ragged_k(oracle):
    if not oracle.has_fast_path:
        pytest.skip("host has no AVX2 + FMA")
    rng = np.random.default_rng(5)
    for K in (1, 15, 16, 17, 33, 250):
        E = rng.standard_normal((K, 256), dtype=np.float32)
        z = rng.standard_normal((3, 256, 2, 5), dtype=np.float32)
        z[0, 3, 1, 1] = np.nan
        z[1, :, 0, 2] = np.inf
        z[2, 7, 1, 4] = -np.inf
        if K > 20:
            E[K // 2, 5] = np.nan
            E[3, 9] = np.inf
        a, b = oracle.forward(z, E), oracle.forward(z, E, fast=True)
        assert np.array_equal(a["idx"], b["idx"]) and a["tie_rows"] == b["tie_rows"], K
        assert np.array_equal(a["dist_min"], b["dist_min"], equal_nan=True), K


def test_log_onehot_to_index_oracle_matches_torch_argmax():
    """log_onehot_to_index = log_x.argmax(1) (vq_diffusion.py:37-38): round trip through index_to_log_onehot, ties, NaN."""
    import torch
    from oracle.vq_oracle import index_to_log_onehot_np, log_onehot_to_index_np
    rng = np.random.default_rng(3)
    x = rng.integers(0, 17, size=(3, 5, 4), dtype=np.int64)
    assert np.array_equal(log_onehot_to_index_np(index_to_log_onehot_np(x, 17)), x)
    v = rng.standard_normal((4, 9, 6)).astype(np.float32)
    v[0, 3, 1] = v[0, 5, 1] = 7.0                       # tie: first wins
    v[1, 4, 2] = np.nan
    v[1, 6, 2] = np.nan                                 # first NaN wins
    v[2, :, 0] = -np.inf
    assert np.array_equal(log_onehot_to_index_np(v), torch.from_numpy(v).argmax(1).numpy())


# ------------------------------------------------------------------------------------------------------------------
# Normalise + cdist recipe (VQGaussianDiffusion3DWrapper.gaussian_to_indices, diffusion_gaussian3d.py:543-570)
from cases import CDIST_CASES, make_cdist_inputs  # noqa: E402


def classify_cdist(x, table, idx_a, idx_b):
    """Disagreeing rows between two index vectors under the normalise + cdist recipe: a real mismatch is one whose two
    float64 distances (on float64-normalised vectors) differ by more than the fp32 rounding band of the matmul-form squared
    distance (magnitudes ~|x|^2 + |y|^2 = 2, a D + 2 term fp32 sum and the normalisation's own roundings)."""
    rows = x.reshape(-1, x.shape[-1]).astype(np.float64)
    t = table.astype(np.float64)
    rows /= np.maximum(np.linalg.norm(rows, axis=1, keepdims=True), 1e-12)
    t /= np.maximum(np.linalg.norm(t, axis=1, keepdims=True), 1e-12)
    bad = np.nonzero(idx_a != idx_b)[0]
    real = 0
    for r in bad:
        da = ((rows[r] - t[idx_a[r]]) ** 2).sum()
        db = ((rows[r] - t[idx_b[r]]) ** 2).sum()
        if abs(da - db) > (x.shape[-1] + 16) * 2.0 ** -23:
            real += 1
    return dict(mismatch=int(bad.size), real=real)


@pytest.mark.parametrize("name", sorted(CDIST_CASES))
def test_oracle_matches_reference_cdist(name, oracle):
    spec = CDIST_CASES[name]
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    x, table = make_cdist_inputs(spec)
    ref = oracle.nearest_cdist(x, table)
    gidx = gold["idx"].reshape(-1).astype(np.int64)
    cls = classify_cdist(x, table, ref["idx"], gidx)
    assert cls["real"] == 0, cls
    assert cls["mismatch"] <= int(gold["ref_tie_rows"]) + int(gold["ref_ne_fp64"]) + 2, cls
    if spec["table"] == "dup":
        assert ref["tie_rows"] == ref["idx"].size and (ref["idx"] < spec["K"] // 3).all()     # lowest of the three copies


def test_cdist_oracle_against_torch_ops(oracle):
    """The restated pieces against torch on CPU: F.normalize, and argmin(cdist) up to rounding-band rows."""
    import torch
    rng = np.random.default_rng(9)
    x = rng.standard_normal((50, 96)).astype(np.float32) * 3
    t = rng.standard_normal((200, 96)).astype(np.float32)
    ref = oracle.nearest_cdist(x, t)
    th = torch.nn.functional.normalize(torch.from_numpy(t), p=2, dim=-1).numpy()
    assert np.abs(ref["table_hat"] - th).max() <= 2e-7
    idx = torch.cdist(torch.nn.functional.normalize(torch.from_numpy(x), p=2, dim=-1), torch.from_numpy(th)).argmin(-1).numpy()
    assert classify_cdist(x, t, ref["idx"], idx)["real"] == 0
    # zero rows: normalise divides by 1e-12 and leaves them zero; every table row is then at distance sqrt(|y|^2) ~ 1 (the
    # winner is whichever normalised row rounds lowest: inside the rounding band of any other)
    z = np.zeros((2, 96), np.float32)
    out = oracle.nearest_cdist(z, t)
    assert np.abs(out["dist_min"] - 1.0).max() <= 1e-6
