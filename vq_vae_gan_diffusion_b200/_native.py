"""ctypes loader for libvq_b200.so (the C-ABI declared in include/vq_b200.h).

The library is built in-tree (``vq_vae_gan_diffusion_b200/lib/libvq_b200.so``) by :func:`build` with
``nvcc -gencode arch=compute_100a,code=sm_100a``.  There is no fallback of any kind: if the library is missing
or a call fails, a :class:`VQNativeError` is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(_LIB_DIR, "libvq_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "vq_b200.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden", "-shared", "-cudart", "static",
]

VQ_STAT_NAMES = ("tie_rows", "rerank_rows", "fallback_rows", "candidates")
VQ_RECIPES = {"expanded": 0, "diffsq": 1, "cdist_normalized": 2}          # VQ_RECIPE_* of include/vq_b200.h


class VQNativeError(RuntimeError):
    pass


def _sources():
    return sorted(os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cu", ".cuh"))) + [HEADER_PATH]


def _source_digest() -> str:
    """Content hash of everything the library is built from (the tree is copied between machines, so file times say
    nothing about whether lib/libvq_b200.so matches the sources)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for path in _sources():
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/vq_api.cu (which includes every kernel) into lib/libvq_b200.so.  Cross-compiles without a GPU.
    A no-op when the library was built from the current sources (content hash in lib/libvq_b200.so.src)."""
    digest = _source_digest()
    stamp = LIB_PATH + ".src"
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB_PATH
    os.makedirs(_LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    tmp = f"{LIB_PATH}.{os.getpid()}.tmp"
    cmd = [nvcc, *NVCC_FLAGS, "-ccbin", ccbin, "-o", tmp, os.path.join(_CSRC, "vq_api.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise VQNativeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    os.replace(tmp, LIB_PATH)
    with open(stamp + ".tmp", "w") as f:
        f.write(digest + "\n")
    os.replace(stamp + ".tmp", stamp)
    return LIB_PATH


_lib = None
_lock = threading.Lock()

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_f32 = ctypes.c_float
_sz = ctypes.c_size_t

_SIGNATURES = {
    "vq_abi_version": (_int, []),
    "vq_last_error": (ctypes.c_char_p, []),
    "vq_last_launch_count": (_int, []),
    "vq_device_check": (_int, []),
    "vq_padded_codes": (_int, [_int]),
    "vq_workspace_bytes": (_int, [_i64, _int, _int, ctypes.POINTER(_sz)]),
    "vq_prepare_codebook": (_int, [_vp, _int, _int, _vp, _vp, _vp, _vp]),
    "vq_argmin": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp, _vp, _int, _vp, _vp, _vp, _sz, _vp]),
    "vq_argmin_narrow": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp, _vp, _int, _vp, _int, _vp, _vp, _sz, _vp]),
    "vq_argmin_rows": (_int, [_vp, _i64, _int, _vp, _vp, _vp, _vp, _int, _int, _vp, _int, _vp, _vp, _sz, _vp]),
    "vq_normalize_rows": (_int, [_vp, _i64, _int, _vp, _vp]),
    "vq_forward": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp, _vp, _int, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "vq_backward": (_int, [_vp, ctypes.POINTER(_i64), _f32, _vp, _vp, _vp, _vp, _i64, _i64, _int, _int, _f32, _i64, _vp, _vp, _vp]),
    "vq_backward_workspace_bytes": (_int, [_int, _int, ctypes.POINTER(_sz)]),
    "vq_backward_ex": (_int, [_vp, ctypes.POINTER(_i64), _f32, _vp, _vp, _vp, _vp, _i64, _i64, _int, _int, _f32, _i64, _f32, _int,
                              _vp, _vp, _vp, _vp, _sz, _vp]),
    "vq_forward_ex": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp, _vp, _int, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "vq_prepare_quant_conv": (_int, [_vp, _vp, _vp, _vp]),
    "vq_forward_qconv": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _f32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "vq_allreduce_multimem": (_int, [_vp, _vp, _int, _int, _i64, _vp, _vp]),
    "vq_pack_stats": (_int, [_vp, _vp, _int, _vp, _vp]),
    "vq_embed_nchw": (_int, [_vp, _vp, _i64, _i64, _int, _int, _vp, _vp]),
    "vq_index_to_log_onehot": (_int, [_vp, _i64, _i64, _int, _f32, _vp, _vp]),
    "vq_log_onehot_to_index": (_int, [_vp, _i64, _i64, _int, _vp, _vp]),
    "vq_mask_replace": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _vp, _vp]),
    "vq_profile_enable": (_int, [_int]),
    "vq_profile_collect": (_int, [ctypes.POINTER(_f32), _int, ctypes.POINTER(_int)]),
    "vq_debug_timeline": (_int, [_vp, _int]),
    "vq_debug_scores": (_int, [_vp, _i64, _i64, _int, _vp, _vp, _vp, _int, _vp, _vp, _sz, _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """The loaded library; raises VQNativeError if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                path = os.environ.get("VQ_B200_LIB", LIB_PATH)      # A/B builds of the same ABI (tuning only)
                if not os.path.exists(path):
                    raise VQNativeError(
                        f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU / PyTorch fallback for the VQ hot path)")
                L = ctypes.CDLL(path)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(L, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = L
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().vq_last_error().decode(errors="replace")
        raise VQNativeError(f"{what} failed (rc={rc}): {msg}")


def profile_enable(on: bool) -> None:
    check(lib().vq_profile_enable(1 if on else 0), "vq_profile_enable")


def profile_collect():
    """Milliseconds of the distance-GEMM kernel for each profiled call since the last collect (host sync)."""
    buf = (_f32 * 512)()
    n = _int(0)
    check(lib().vq_profile_collect(buf, 512, ctypes.byref(n)), "vq_profile_collect")
    return [float(buf[i]) for i in range(n.value)]


def padded_codes(K: int) -> int:
    return int(lib().vq_padded_codes(int(K)))


def workspace_bytes(N: int, K: int, D: int) -> int:
    out = _sz(0)
    check(lib().vq_workspace_bytes(int(N), int(K), int(D), ctypes.byref(out)), "vq_workspace_bytes")
    return int(out.value)


_ws_cache = {}
_bws_cache = {}


def backward_workspace_bytes_cached(K: int, D: int) -> int:
    """Scratch the deterministic backward needs (vq_backward_workspace_bytes)."""
    v = _bws_cache.get((K, D))
    if v is None:
        out = _sz(0)
        check(lib().vq_backward_workspace_bytes(int(K), int(D), ctypes.byref(out)), "vq_backward_workspace_bytes")
        v = _bws_cache[(K, D)] = int(out.value)
    return v


def workspace_bytes_cached(N: int, K: int, D: int) -> int:
    key = (N, K, D)
    v = _ws_cache.get(key)
    if v is None:
        if len(_ws_cache) > 256:
            _ws_cache.clear()
        v = _ws_cache[key] = workspace_bytes(N, K, D)
    return v
