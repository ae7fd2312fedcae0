"""dge_b200 — B200-native (sm_100a) differentiable Gaussian-splatting rasterizer for DGE.

Only the hot path of bonapark00/DGE lives here: the CUDA kernels and their
C-ABI (csrc/, include/dge_b200.h), the Python mirror of the reference's
`diff_gaussian_rasterization` binding, and the view-sharded fit step. `install()` makes
`import diff_gaussian_rasterization` resolve to this implementation so DGE's renderer and
GaussianModel run unchanged.
"""
import sys

__all__ = ["install", "diff_gaussian_rasterization"]


def install():
    """Register the drop-in under the reference's module name."""
    from . import diff_gaussian_rasterization as mod
    sys.modules["diff_gaussian_rasterization"] = mod
    return mod
