"""Stage the reference's own fuser sources under oracle/_ref/ -- TEST / BASELINE INFRASTRUCTURE ONLY.

    python oracle/build_ref.py

The reference (olivesgatech/R3D) is pure Python: there is nothing to compile.  What `bench.py --impl reference` and
the `cpu_baseline` leg need on the GPU box -- where /root/reference does not exist -- is the reference's own
`CMFuser` code, so this recipe copies the handful of files the fuser path imports from /root/reference into
oracle/_ref/ (git-ignored: reference sources never enter this repository's history; NOT gpurun-ignored, so the
directory travels to the GPU box like a built .so).  Nothing is modified; oracle/ref_loader.py applies the two
outside shims of SURVEY.md appendix A (a matplotlib stub, a wrapper around the hard-coded `.to('cuda')`) at import
time.

Only tests/, __graft_entry__.smoke()/build() and bench.py's CPU legs may use oracle/; the product never does.
"""
from __future__ import annotations

import os
import shutil
import sys

REF = os.environ.get("R3D_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

FILES = [
    "model/futr_safuser_tokenfusion.py",
    "model/futr_safuser_tokenfusion_vary.py",
    "model/futr_safuser_batchnormalization.py",
    "model/futr_safuser_depth.py",
    "model/extras/transformerblock.py",
    "model/extras/transformer.py",
    "model/extras/position.py",
]


def build(verbose: bool = True) -> bool:
    """Returns True when oracle/_ref is populated (now or already), False when the reference tree is absent."""
    if not os.path.isdir(REF):
        return os.path.isdir(os.path.join(OUT, "model"))
    for rel in FILES:
        src = os.path.join(REF, rel)
        if not os.path.exists(src):
            raise FileNotFoundError(src)
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    for d in ("model", "model/extras"):
        init = os.path.join(OUT, d, "__init__.py")
        if not os.path.exists(init) and os.path.exists(os.path.join(REF, d, "__init__.py")):
            shutil.copyfile(os.path.join(REF, d, "__init__.py"), init)
    if verbose:
        print(f"oracle/_ref: {len(FILES)} reference files staged from {REF}")
    return True


if __name__ == "__main__":
    ok = build()
    sys.exit(0 if ok else 1)
