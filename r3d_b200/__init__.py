"""r3d_b200 -- B200-native Rank-enhancing Token Fuser + effective rank for R3D.

Drop-in surface: ``CMFuser`` (reference: model/futr_safuser_*.py :: CMFuser) and the
operators in :mod:`r3d_b200.ops`.  All compute runs in hand-written sm_100a CUDA
kernels behind the C ABI of ``include/r3d_b200.h``; there is no CPU fallback.
"""
from ._lib import LIB_PATH, R3DError, launch_count  # noqa: F401
from . import ops  # noqa: F401
from .fuser import (CMFuser, TokenFusionCMFuser, VaryCMFuser, BatchNormCMFuser, SAFuser, Block, VARIANTS)  # noqa: F401
from .ops import erank, channel_score, bottomk, exchange  # noqa: F401
from .embed import RGBEmbed, DepthEmbed, FuserFront  # noqa: F401
from .futr import FUTR  # noqa: F401

__all__ = ["CMFuser", "TokenFusionCMFuser", "VaryCMFuser", "BatchNormCMFuser", "SAFuser", "Block", "ops", "erank",
           "channel_score", "bottomk", "exchange", "RGBEmbed", "DepthEmbed", "FuserFront", "FUTR", "R3DError", "launch_count", "LIB_PATH", "VARIANTS"]
