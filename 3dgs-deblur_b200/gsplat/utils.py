"""`gsplat.utils` -- tile binning helpers (operator surface of the reference's gsplat/utils.py:1-182)."""
from typing import Tuple

from torch import Tensor

import gsplat.cuda as _C


def map_gaussian_to_intersects(num_points: int, num_intersects: int, xys: Tensor, depths: Tensor, radii: Tensor,
                               cum_tiles_hit: Tensor, tile_bounds: Tuple[int, int, int], block_size: int):
    """(isect_ids, gaussian_ids): key = (tile_id << 32) | depth bits, one entry per (Gaussian, tile)."""
    return _C.map_gaussian_to_intersects(num_points, num_intersects, xys.contiguous(), depths.contiguous(),
                                         radii.contiguous(), cum_tiles_hit.contiguous(), tile_bounds, block_size)


def get_tile_bin_edges(num_intersects: int, isect_ids_sorted: Tensor, tile_bounds: Tuple[int, int, int]) -> Tensor:
    """tile_bins[t] = [first, one-past-last) index of tile t in the sorted list; (0, 0) for empty tiles."""
    return _C.get_tile_bin_edges(num_intersects, isect_ids_sorted.contiguous(), tile_bounds)


def compute_cov2d_bounds(cov2d: Tensor):
    """(conics (N,3), radii (N,1)) from upper-triangular 2D covariances (N,3)."""
    assert cov2d.shape[-1] == 3, (
        f"Expected input cov2d to be of shape (*batch, 3) (upper triangular values), but got {tuple(cov2d.shape)}")
    num_pts = cov2d.shape[0]
    assert num_pts > 0
    return _C.compute_cov2d_bounds(num_pts, cov2d.contiguous())


def compute_cumulative_intersects(num_tiles_hit: Tensor):
    """(num_intersects: int, cum_tiles_hit: int32 inclusive scan).  One host sync, as in the reference."""
    return _C.cumulative_intersects(num_tiles_hit.contiguous())


def bin_and_sort_gaussians(num_points: int, num_intersects: int, xys: Tensor, depths: Tensor, radii: Tensor,
                           cum_tiles_hit: Tensor, tile_bounds: Tuple[int, int, int], block_size: int):
    """(isect_ids_unsorted, gaussian_ids_unsorted, isect_ids_sorted, gaussian_ids_sorted, tile_bins).

    Ordering: ascending (tile, depth bits); ties keep emission order (ascending Gaussian id) -- the
    reference's torch.sort leaves ties unspecified (utils.py:179)."""
    isect_ids, gaussian_ids = map_gaussian_to_intersects(num_points, num_intersects, xys, depths, radii,
                                                         cum_tiles_hit, tile_bounds, block_size)
    num_tiles = int(tile_bounds[0]) * int(tile_bounds[1])
    isect_ids_sorted, gaussian_ids_sorted = _C.sort_intersects(num_tiles, isect_ids, gaussian_ids)
    tile_bins = get_tile_bin_edges(num_intersects, isect_ids_sorted, tile_bounds)
    return isect_ids, gaussian_ids, isect_ids_sorted, gaussian_ids_sorted, tile_bins
