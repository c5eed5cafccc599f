"""Recipe for the reference arm of bench.py: make the UNMODIFIED pure-Python reference importable on the GPU box.

The reference has no packaging (no setup.py / pyproject), so it cannot be pip-installed; its model path is plain Python
over torch / einops / yaml, all present in the image.  This script copies the `doc2tex/` package tree (Python sources only)
from /root/reference to baseline/_ref/, which is git-ignored (never part of this repository's history) but travels to the
GPU box with the gpurun snapshot.  `bench.py --impl reference` then times `doc2tex.modules.build_model.Model` itself
(cpu_baseline.kind = "reference"); without the copy it falls back to the oracle port (kind = "port").

    python oracle/install_reference.py        # run in the build container; __graft_entry__.build() calls it too
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/doc2tex"
DST = os.path.join(ROOT, "baseline", "_ref", "doc2tex")


def install(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"[install_reference] {SRC} not present (GPU box?): keeping whatever baseline/_ref already holds")
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.js", "*.json", "*.md", "*.txt"))
    if verbose:
        n = sum(len(f) for _, _, f in os.walk(DST))
        print(f"[install_reference] copied {n} files to {os.path.relpath(DST, ROOT)}")
    return True


def check() -> None:
    """Import the copy and build the model once (CPU) to prove the copy is sufficient."""
    sys.path.insert(0, os.path.dirname(DST))
    sys.path.insert(0, ROOT)
    import copy
    import torch
    from doc2tex.modules.build_model import Model
    from doc2tex_b200 import synth
    cfg = synth.make_config("TFM")
    m = Model(copy.deepcopy(cfg)).eval()
    m.load_state_dict(synth.make_state_dict(cfg, seed=1111), strict=True)
    with torch.no_grad():
        ctx, grid, pad = m.forward_encoder(synth.make_images(1, 64, 256))
    print(f"[install_reference] reference Model imported from {os.path.dirname(DST)}: ctx {tuple(ctx.shape)} grid {tuple(grid)}")


if __name__ == "__main__":
    if install():
        check()
