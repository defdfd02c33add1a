"""B200-native vector-quantiser hot path of hongrui16/VQ-VAE-GAN-Diffusion.

Public surface (mirrors the reference's ``network/vqvae/submodule/codebook.py``):

    from vq_vae_gan_diffusion_b200 import CodeBook          # drop-in nn.Module
    vq_vae_gan_diffusion_b200.install()                      # make the reference import OUR CodeBook

Everything numeric runs in ``lib/libvq_b200.so`` (hand-written sm_100a CUDA behind the C-ABI of
``include/vq_b200.h``).  There is no CPU fallback.
"""
from __future__ import annotations

import sys
import types

from . import _native
from ._native import VQNativeError, build
from .codebook import CodeBook, vq_embed_nchw
from .postconv import FoldedPostQuant
from .preconv import FoldedQuantConv, FoldedVQ
from .nearest import CodeTable, gaussian_to_indices, nearest_indices
from .tokens import blend_with_sos, index_to_log_onehot, log_onehot_to_index, mask_and_replace

__all__ = ["CodeBook", "FoldedPostQuant", "FoldedQuantConv", "FoldedVQ", "vq_embed_nchw", "CodeTable", "nearest_indices", "gaussian_to_indices", "index_to_log_onehot", "log_onehot_to_index", "mask_and_replace", "blend_with_sos",
           "VQNativeError", "build", "install"]

REFERENCE_MODULE = "network.vqvae.submodule.codebook"


def install() -> None:
    """Register this package's CodeBook under the reference's module path so that
    ``from network.vqvae.submodule.codebook import CodeBook`` (network/vqvae/vqvae.py:17) resolves to it.
    Call before importing ``network.vqvae.vqvae``."""
    from . import codebook as _cb
    mod = types.ModuleType(REFERENCE_MODULE)
    mod.CodeBook = _cb.CodeBook
    mod.__doc__ = "B200-native replacement installed by vq_vae_gan_diffusion_b200.install()"
    sys.modules[REFERENCE_MODULE] = mod
