"""Import the UNMODIFIED reference fuser classes -- TEST / BASELINE INFRASTRUCTURE ONLY.

Looks for the reference sources in /root/reference (this container) or oracle/_ref (staged by oracle/build_ref.py; the
GPU box).  Two shims are applied from OUTSIDE the reference tree: a `matplotlib` stub
(model/extras/transformer.py:15 imports it, the fuser never uses it) and a wrapper around
`CMFuser.generate_cross_attention_mask` whose result ignores the hard-coded `.to('cuda')`
(model/futr_safuser_tokenfusion.py:77) when the module runs on the CPU.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_MODULES = {
    "tokenfusion": "model.futr_safuser_tokenfusion",
    "vary": "model.futr_safuser_tokenfusion_vary",
    "batchnorm": "model.futr_safuser_batchnormalization",
    "safuser": "model.futr_safuser_depth",
}


def reference_root():
    for root in (os.environ.get("R3D_REFERENCE", "/root/reference"), os.path.join(HERE, "_ref")):
        if os.path.exists(os.path.join(root, "model", "futr_safuser_tokenfusion.py")):
            return root
    return None


class _StayOnDevice:
    def __init__(self, t):
        self.t = t

    def to(self, *a, **k):
        return self.t


def load_futr(variant: str = "tokenfusion"):
    """-> the reference module's FUTR class with its CMFuser shimmed for the CPU, or None."""
    if load(variant) is None:
        return None
    return importlib.import_module(_MODULES[variant]).FUTR


def load(variant: str = "tokenfusion"):
    """-> the reference module's CMFuser class (CPU-runnable), or None when no reference tree is available."""
    root = reference_root()
    if root is None:
        return None
    if "matplotlib" not in sys.modules:
        mp, pp = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        mp.pyplot = pp
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mp, pp
    if root not in sys.path:
        sys.path.insert(0, root)
    mod = importlib.import_module(_MODULES[variant])
    cls = mod.CMFuser
    if not getattr(cls, "_r3d_mask_shim", False) and hasattr(cls, "generate_cross_attention_mask"):
        orig = cls.generate_cross_attention_mask
        cls.generate_cross_attention_mask = staticmethod(lambda sz, _o=orig: _StayOnDevice(_o(sz)))
        cls._r3d_mask_shim = True
    return cls
