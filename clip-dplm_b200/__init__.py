"""clip_dplm_b200 -- B200-native fused CLIP / InfoNCE hot path (drop-in for SrikarK-code/clip-dplm).

Public surface
    fused_clip_loss(a, b, logit_scale, ...)      the fused op (autograd-aware, optionally row-sharded)
    graph.GraphedClipStep                        forward + backward of one fixed shape replayed as a single CUDA graph
    modules.*                                    drop-in modules / loss functions with the reference's signatures
    engine.CudaEngine                            stage-level access to the C-ABI (include/clipnce.h)

The CUDA library (clip-dplm_b200/csrc/libclipnce.so) is loaded lazily on first use; a missing library
raises -- there is no CPU or eager fallback on the product path.
"""
from . import _lib  # noqa: F401
from .functional import fused_clip_loss  # noqa: F401

__all__ = ["fused_clip_loss"]
__version__ = "0.1.0"
