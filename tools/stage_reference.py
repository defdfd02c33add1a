#!/usr/bin/env python
"""Stage the UNMODIFIED reference tree under git-ignored ``baseline/_ref/`` so it travels to the GPU box.

``/root/reference`` exists only in the build container; ``gpurun`` ships the working tree (git-ignored paths included, see
.gitignore: ``baseline/_ref/`` is ignored by git, not by gpurun).  What is staged: the Python packages the hot path's
callers live in (``network/``) and the YAML configs that size them (``configs/``) -- byte for byte, nothing edited.  It is
used by

* ``bench.py --impl reference``: times the reference's own ``CodeBook`` (``kind: "reference"``);
* ``tests/test_gpu_dropin_reference.py``: runs the reference's ``VQVAE`` / ``VQTransformer`` / ``VQDiffusion`` with the
  B200 ``CodeBook`` swapped in by ``install()``, against the same models with the unmodified class.

Nothing under ``vq_vae_gan_diffusion_b200/`` reads it.  The staged copy never enters git history.
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("VQ_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
PARTS = ("network", "configs")


def staged_root():
    """Directory to put on sys.path to import the reference (``network.*``): the staged copy, else the source tree,
    else None."""
    if os.path.isfile(os.path.join(DST, "network", "vqvae", "submodule", "codebook.py")):
        return DST
    if os.path.isfile(os.path.join(SRC, "network", "vqvae", "submodule", "codebook.py")):
        return SRC
    return None


def stage(verbose: bool = False) -> str | None:
    """Copy SRC/{network,configs} to baseline/_ref (only when the source tree exists; idempotent)."""
    if not os.path.isdir(os.path.join(SRC, "network")):
        return staged_root()
    for part in PARTS:
        s, d = os.path.join(SRC, part), os.path.join(DST, part)
        for dirpath, dirnames, filenames in os.walk(s):
            dirnames[:] = [x for x in dirnames if x != "__pycache__"]
            rel = os.path.relpath(dirpath, s)
            os.makedirs(os.path.join(d, rel), exist_ok=True)
            for f in filenames:
                if not f.endswith((".py", ".yml", ".yaml")):
                    continue
                a, b = os.path.join(dirpath, f), os.path.join(d, rel, f)
                if not os.path.exists(b) or not filecmp.cmp(a, b, shallow=False):
                    shutil.copyfile(a, b)
                    if verbose:
                        print("staged", os.path.relpath(b, ROOT))
    return DST


if __name__ == "__main__":
    print(stage(verbose="-v" in sys.argv))
