"""Drop-in for DreamScene's `from diff_gaussian_rasterization import
GaussianRasterizationSettings, GaussianRasterizer` (/root/reference/scene_gaussian.py:11-12).
Everything is implemented by dreamscene_b200 (hand-written sm_100a CUDA behind a C ABI)."""
from dreamscene_b200.rasterizer import (GaussianRasterizationSettings, GaussianRasterizer,
                                        rasterize_gaussians)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians"]
