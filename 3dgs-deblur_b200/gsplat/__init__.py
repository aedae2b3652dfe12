"""Drop-in `gsplat` package: the operator surface of SpectacularAI/3dgs-deblur's gsplat fork
(/root/reference/gsplat/gsplat/__init__.py:1-166) backed by the B200 C-ABI library libb200splat.so.

Put `3dgs-deblur_b200/` on PYTHONPATH and `from gsplat.project_gaussians import project_gaussians`,
`from gsplat.rasterize import rasterize_gaussians`, `from gsplat.sh import spherical_harmonics` (the
imports of nerfstudio/models/splatfacto.py:28-31) resolve here unchanged.
"""
import warnings
from typing import Any

import torch

from .project_gaussians import project_gaussians
from .rasterize import rasterize_gaussians
from .sh import spherical_harmonics
from .utils import (
    bin_and_sort_gaussians,
    compute_cov2d_bounds,
    compute_cumulative_intersects,
    get_tile_bin_edges,
    map_gaussian_to_intersects,
)
from .version import __version__

__all__ = [
    "__version__",
    "project_gaussians",
    "rasterize_gaussians",
    "spherical_harmonics",
    "bin_and_sort_gaussians",
    "compute_cumulative_intersects",
    "compute_cov2d_bounds",
    "get_tile_bin_edges",
    "map_gaussian_to_intersects",
    # deprecated Function.apply() aliases kept by the reference (gsplat/__init__.py:43-166)
    "ProjectGaussians",
    "RasterizeGaussians",
    "BinAndSortGaussians",
    "ComputeCumulativeIntersects",
    "ComputeCov2dBounds",
    "GetTileBinEdges",
    "MapGaussiansToIntersects",
    "SphericalHarmonics",
    "NDRasterizeGaussians",
]


def _deprecated(name: str, new_name: str, fn):
    """Forward-only autograd.Function shim that warns and delegates, like the reference's aliases."""

    def forward(ctx, *args, **kwargs):
        warnings.warn(f"{name} is deprecated, use {new_name} instead", DeprecationWarning)
        return fn(*args, **kwargs)

    def backward(ctx: Any, *grad_outputs: Any) -> Any:
        raise NotImplementedError

    return type(name, (torch.autograd.Function,), {"forward": staticmethod(forward), "backward": staticmethod(backward)})


MapGaussiansToIntersects = _deprecated("MapGaussiansToIntersects", "map_gaussian_to_intersects", map_gaussian_to_intersects)
ComputeCumulativeIntersects = _deprecated("ComputeCumulativeIntersects", "compute_cumulative_intersects", compute_cumulative_intersects)
ComputeCov2dBounds = _deprecated("ComputeCov2dBounds", "compute_cov2d_bounds", compute_cov2d_bounds)
GetTileBinEdges = _deprecated("GetTileBinEdges", "get_tile_bin_edges", get_tile_bin_edges)
BinAndSortGaussians = _deprecated("BinAndSortGaussians", "bin_and_sort_gaussians", bin_and_sort_gaussians)
ProjectGaussians = _deprecated("ProjectGaussians", "project_gaussians", project_gaussians)
RasterizeGaussians = _deprecated("RasterizeGaussians", "rasterize_gaussians", rasterize_gaussians)
NDRasterizeGaussians = _deprecated("NDRasterizeGaussians", "rasterize_gaussians", rasterize_gaussians)
SphericalHarmonics = _deprecated("SphericalHarmonics", "spherical_harmonics", spherical_harmonics)
