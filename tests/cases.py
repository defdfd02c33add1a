"""Seeded synthetic inputs shared by the golden-vector generator, the oracle tests and the GPU parity tests.

Inputs are regenerated from seeds (numpy PCG64) instead of being stored, so tests/golden/ only has to hold the
reference's OUTPUTS.  Distributions follow SURVEY.md section 8(d):
  "init"    E ~ U(-1/K, 1/K) (codebook.py:43-45), z ~ N(0,1)            -> many fp32 near-ties
  "trained" E ~ N(0,1), z = E[randint(K)] + 0.3 N(0,1)                  -> well separated
"""
from __future__ import annotations

import numpy as np

# name -> spec.  D is 256 everywhere the CUDA path runs (latent_dim of every reference config).
CASES = {
    # tiny, full arrays stored
    "small_init":     dict(B=2, H=4, W=4, K=64, D=256, dist="init", seed=101),
    "small_trained":  dict(B=2, H=4, W=4, K=64, D=256, dist="trained", seed=102),
    # ragged: N and K are not multiples of any tile size
    "ragged_init":    dict(B=3, H=5, W=7, K=100, D=256, dist="init", seed=103),
    "ragged_trained": dict(B=3, H=5, W=7, K=333, D=256, dist="trained", seed=104),
    # BASELINE.json configs[0]: VQGAN small on 28x28 -> encoder output 1x1, batch 200, K=1024
    "cfg1_init":      dict(B=200, H=1, W=1, K=1024, D=256, dist="init", seed=105),
    # BASELINE.json configs[1] at 1/8 batch: 16x16 latents, K=1024
    "cfg2s_init":     dict(B=8, H=16, W=16, K=1024, D=256, dist="init", seed=106),
    "cfg2s_trained":  dict(B=8, H=16, W=16, K=1024, D=256, dist="trained", seed=107),
    # large-K tie census (configs[2]/[3] codebooks at reduced N)
    "k8192_init":     dict(B=1, H=32, W=32, K=8192, D=256, dist="init", seed=108),
    "k16384_trained": dict(B=1, H=16, W=32, K=16384, D=256, dist="trained", seed=109),
    # edge cases of the domain
    "dup_rows":       dict(B=2, H=8, W=8, K=96, D=256, dist="dup", seed=110),      # duplicated code rows -> lowest index
    "exact_hit":      dict(B=2, H=8, W=8, K=128, D=256, dist="exact", seed=111),   # z equals a code -> distance 0
    "zero_codebook":  dict(B=1, H=4, W=8, K=40, D=256, dist="zero", seed=112),     # every code ties -> index 0
    # oracle-only known answer (D != 256)
    "k4_d2":          dict(B=1, H=1, W=3, K=4, D=2, dist="kat", seed=0),
}

FULL_ARRAY_LIMIT = 1 << 15      # cases with N*D below this store z_q / grads whole, others store samples
N_SAMPLES = 2048


def make_inputs(spec: dict):
    """-> z (B, D, H, W) fp32 C-contiguous, E (K, D) fp32, g_out_nhwc (B, H, W, D) fp32."""
    B, H, W, K, D = spec["B"], spec["H"], spec["W"], spec["K"], spec["D"]
    rng = np.random.default_rng(spec["seed"])
    dist = spec["dist"]
    N = B * H * W
    if dist == "kat":
        # hand-computed case: codes on the corners of a square, three latents
        E = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]], np.float32)
        zf = np.array([[0.1, 0.2], [0.9, 0.2], [0.6, 0.9]], np.float32)      # -> codes 0, 1, 3
        z = np.ascontiguousarray(zf.reshape(B, H, W, D).transpose(0, 3, 1, 2))
        g = np.array([[1.0, -1.0], [0.5, 0.25], [-2.0, 4.0]], np.float32).reshape(B, H, W, D)
        return z, E, g
    if dist == "init":
        E = rng.uniform(-1.0 / K, 1.0 / K, size=(K, D)).astype(np.float32)
        zf = rng.standard_normal((N, D), dtype=np.float32)
    elif dist == "trained":
        E = rng.standard_normal((K, D), dtype=np.float32)
        zf = E[rng.integers(0, K, size=N)] + np.float32(0.3) * rng.standard_normal((N, D), dtype=np.float32)
    elif dist == "dup":
        base = rng.standard_normal((K // 3, D), dtype=np.float32)
        E = np.concatenate([base, base, base], axis=0)[:K]                      # rows k, k+K/3, k+2K/3 identical
        zf = base[rng.integers(0, K // 3, size=N)] + np.float32(0.1) * rng.standard_normal((N, D), dtype=np.float32)
    elif dist == "exact":
        E = rng.standard_normal((K, D), dtype=np.float32)
        zf = E[rng.integers(0, K, size=N)].copy()
    elif dist == "zero":
        E = np.zeros((K, D), np.float32)
        zf = rng.standard_normal((N, D), dtype=np.float32)
    else:
        raise ValueError(dist)
    z = np.ascontiguousarray(zf.reshape(B, H, W, D).transpose(0, 3, 1, 2)).astype(np.float32)
    g = rng.standard_normal((B, H, W, D), dtype=np.float32)
    return z, np.ascontiguousarray(E, dtype=np.float32), g


def sample_positions(n_elems: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed + 7919)
    return np.sort(rng.choice(n_elems, size=min(N_SAMPLES, n_elems), replace=False)).astype(np.int64)


# ---------------------------------------------------------------------------------------------------------------
# Row-major nearest-code search: GaussianDiffusion2D.gaussian_to_indices,
# diffusion_gaussian2d.py:322-347.  Table = torch.rand(K, gaussian_dim) there (:287), gaussian_dim = 96 in configs/*.yml.
NN_CASES = {
    "nn_g96_clean":   dict(B=4, L=64, K=1024, D=96, noise=0.05, distribute_dim=2, seed=201),   # denoised vectors near table rows
    "nn_g96_noisy":   dict(B=3, L=50, K=2048, D=96, noise=1.0, distribute_dim=2, seed=202),    # far from every row: near ties
    "nn_g96_dim1":    dict(B=2, L=32, K=300, D=96, noise=0.2, distribute_dim=1, seed=203),     # (B, D, L) input, permuted inside
    "nn_d256":        dict(B=2, L=40, K=512, D=256, noise=0.3, distribute_dim=2, seed=204),
}


def make_nn_inputs(spec: dict):
    """-> x (B, L, D) fp32, table (K, D) fp32 (uniform [0, 1) like the reference's buffer)."""
    rng = np.random.default_rng(spec["seed"])
    B, L, K, D = spec["B"], spec["L"], spec["K"], spec["D"]
    table = rng.random((K, D), dtype=np.float32)
    x = table[rng.integers(0, K, size=B * L)] + np.float32(spec["noise"]) * rng.standard_normal((B * L, D), dtype=np.float32)
    return np.ascontiguousarray(x.reshape(B, L, D), dtype=np.float32), table


# Broadcast-difference search at the end of V_VQDiffusion.sample (v_vq_diffusion.py:114-123): sampled embeddings
# (B, L, 256) against the VQVAE codebook (K, 256).
DIFFSQ_CASES = {
    "dsq_trained":  dict(B=2, L=64, K=512, D=256, dist="trained", noise=0.3, seed=301),
    "dsq_init":     dict(B=2, L=48, K=1024, D=256, dist="init", noise=1.0, seed=302),     # codebook at its U(-1/K, 1/K) init
    "dsq_ragged":   dict(B=3, L=37, K=333, D=256, dist="trained", noise=0.6, seed=303),
}


def make_diffsq_inputs(spec: dict):
    """-> x (B, L, D) fp32, codebook (K, D) fp32."""
    rng = np.random.default_rng(spec["seed"])
    B, L, K, D = spec["B"], spec["L"], spec["K"], spec["D"]
    if spec["dist"] == "init":
        E = rng.uniform(-1.0 / K, 1.0 / K, size=(K, D)).astype(np.float32)
        x = rng.standard_normal((B * L, D), dtype=np.float32) * np.float32(spec["noise"])
    else:
        E = rng.standard_normal((K, D), dtype=np.float32)
        x = E[rng.integers(0, K, size=B * L)] + np.float32(spec["noise"]) * rng.standard_normal((B * L, D), dtype=np.float32)
    return np.ascontiguousarray(x.reshape(B, L, D), dtype=np.float32), E


# Normalise + cdist search of VQGaussianDiffusion3DWrapper.gaussian_to_indices (diffusion_gaussian3d.py:543-570).  The wrapper's
# table is positional_encoding(gaussian_dim, vocab_size) (diffusion_gaussian3d.py:48-54, :513): sin / cos of position x
# frequency -- restated here because it is the realistic input (neighbouring rows are close: near-ties), next to random tables.
CDIST_CASES = {
    "cd_pe512":     dict(B=2, L=64, K=1024, D=512, table="pe", noise=0.05, seed=501),     # the 3D wrapper's own shape family
    "cd_pe96":      dict(B=3, L=50, K=1024, D=96, table="pe", noise=0.3, seed=502),       # gaussian_dim: 96 of configs/*.yml
    "cd_rand96":    dict(B=2, L=40, K=2048, D=96, table="rand", noise=1.0, seed=503),     # far from every row
    "cd_ragged":    dict(B=3, L=37, K=333, D=70, table="rand", noise=0.2, seed=504),      # width off every vector path
    "cd_rand256":   dict(B=2, L=32, K=512, D=256, table="rand", noise=0.5, seed=505),
    "cd_dup128":    dict(B=2, L=48, K=300, D=128, table="dup", noise=0.1, seed=506),      # every row three times: lowest index wins
}


def positional_table(dim: int, num_vectors: int) -> np.ndarray:
    """diffusion_gaussian3d.py:48-54 (float64 numpy, then float32)."""
    position = np.arange(num_vectors)[:, np.newaxis]
    div_term = np.exp(np.arange(0, dim, 2) * -(np.log(10000.0) / dim))
    pe = np.zeros((num_vectors, dim))
    pe[:, 0::2] = np.sin(position * div_term)
    pe[:, 1::2] = np.cos(position * div_term)
    return pe.astype(np.float32)


def make_cdist_inputs(spec: dict):
    """-> x (B, L, D) fp32 (unnormalised predictions around table rows), table (K, D) fp32."""
    rng = np.random.default_rng(spec["seed"])
    B, L, K, D = spec["B"], spec["L"], spec["K"], spec["D"]
    if spec["table"] == "pe":
        table = positional_table(D, K)
    elif spec["table"] == "dup":
        base = rng.standard_normal((K // 3, D), dtype=np.float32)
        table = np.concatenate([base, base, base], axis=0)[:K]
    else:
        table = rng.standard_normal((K, D), dtype=np.float32)
    x = table[rng.integers(0, K, size=B * L)] * np.float32(1.7) + np.float32(spec["noise"]) * rng.standard_normal((B * L, D), dtype=np.float32)
    return np.ascontiguousarray(x.reshape(B, L, D), dtype=np.float32), np.ascontiguousarray(table)


# Token-stream formats after the tokeniser.  index_to_log_onehot: the vector path (L % 4 == 0), the
# scalar path, trailing dims beyond one (the VQ-Diffusion tokens are (B, L); the function is rank-generic), the class
# count off the 32-class tile, the [MASK] class (num_classes = K + 1) unused.
TOKEN_ONEHOT_CASES = {
    "tok_onehot_vec":    dict(shape=(3, 64), num_classes=45, seed=401),
    "tok_onehot_scalar": dict(shape=(2, 37), num_classes=33, seed=402),
    "tok_onehot_2d":     dict(shape=(2, 6, 10), num_classes=17, seed=403),
    "tok_onehot_mask":   dict(shape=(2, 256), num_classes=129, seed=404, high=128),       # class 128 = [MASK], never set
}

# VQTransformer.forward's input corruption (vqTransformer.py:117-141): pkeep 0.5 is the reference's default.
TOKEN_BLEND_CASES = {
    "tok_blend_small": dict(shape=(4, 256), num_classes=1024, pkeep=0.5, sos_token=0, seed=411),
    "tok_blend_odd":   dict(shape=(3, 77), num_classes=513, pkeep=0.9, sos_token=512, seed=412),
}


def make_token_indices(spec: dict) -> np.ndarray:
    rng = np.random.default_rng(spec["seed"])
    return rng.integers(0, spec.get("high", spec["num_classes"]), size=spec["shape"], dtype=np.int64)
