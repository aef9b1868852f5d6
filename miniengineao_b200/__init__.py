"""miniengineao_b200 -- B200-native multi-scale SSAO pipeline behind the MiniEngineAO
AmbientOcclusion component surface.  The compute path is libmeao.so (hand-written sm_100a CUDA,
C ABI in include/meao.h); this package is the thin Python host that mirrors the reference
component and its parameters.  There is no CPU fallback."""
from .ambient_occlusion import AmbientOcclusion, Camera  # noqa: F401
from ._native import MeaoError  # noqa: F401

__all__ = ["AmbientOcclusion", "Camera", "MeaoError"]
