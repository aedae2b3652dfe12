"""`gsplat.rasterize` -- tile binning + blur / rolling-shutter alpha blend (operator surface of the
reference's gsplat/rasterize.py:15-294)."""
from typing import Optional

import torch
from torch import Tensor
from torch.autograd import Function

import gsplat.cuda as _C

from .utils import compute_cumulative_intersects


def rasterize_gaussians(
    xys: Tensor,
    depths: Tensor,
    pix_vels: Tensor,
    radii: Tensor,
    conics: Tensor,
    num_tiles_hit: Tensor,
    colors: Tensor,
    opacity: Tensor,
    img_height: int,
    img_width: int,
    block_width: int,
    background: Optional[Tensor] = None,
    return_alpha: Optional[bool] = False,
    rolling_shutter_time: Optional[float] = 0,
    exposure_time: Optional[float] = 0,
    blur_samples: Optional[int] = 1,
) -> Tensor:
    """out_img (H,W,C) [and out_alpha (H,W) if return_alpha]; differentiable w.r.t. xys, pix_vels,
    conics, colors, opacity and background.  `xys.absgrad` receives the per-pixel-sample absolute
    screen-space gradient after backward (rasterize.py:275)."""
    assert block_width > 1 and block_width <= 16, "block_width must be between 2 and 16"
    if colors.dtype == torch.uint8:
        colors = colors.float() / 255
    if background is not None:
        assert background.shape[0] == colors.shape[-1], (
            f"incorrect shape of background color tensor, expected shape {colors.shape[-1]}")
    else:
        background = torch.ones(colors.shape[-1], dtype=torch.float32, device=colors.device)
    if xys.ndimension() != 2 or xys.size(1) != 2:
        raise ValueError("xys must have dimensions (N, 2)")
    if colors.ndimension() != 2:
        raise ValueError("colors must have dimensions (N, D)")
    return _RasterizeGaussians.apply(
        xys.contiguous(), depths.contiguous(), pix_vels.contiguous(), radii.contiguous(), conics.contiguous(),
        num_tiles_hit.contiguous(), colors.contiguous(), opacity.contiguous(), img_height, img_width, block_width,
        background.contiguous(), return_alpha, rolling_shutter_time, exposure_time, blur_samples,
    )


class _RasterizeGaussians(Function):
    @staticmethod
    def forward(ctx, xys, depths, pix_vels, radii, conics, num_tiles_hit, colors, opacity, img_height, img_width,
                block_width, background, return_alpha=False, rolling_shutter_time=0, exposure_time=0, blur_samples=1):
        num_points = xys.size(0)
        tile_bounds = ((img_width + block_width - 1) // block_width, (img_height + block_width - 1) // block_width, 1)
        block = (block_width, block_width, 1)
        img_size = (img_width, img_height, 1)

        three = colors.shape[-1] == 3
        packed = out_alpha = None
        if three:
            # RGB path: pack once, culled two-level binning (ONE host sync: the culled entry count), blend.  The id
            # lists hold only the (tile, Gaussian) pairs that can colour a pixel, in the reference's order; every
            # output is identical to blending the reference's full lists (tests/test_gpu_parity.py).
            n_samples = int(blur_samples)
            if not (0 < n_samples <= 10):
                raise RuntimeError("unsupported blur size")  # bindings.cu:450-452
            packed = _C.pack_records(xys, pix_vels, conics, colors, opacity)
            num_intersects, gaussian_ids_sorted, tile_bins = _C.bin_cull(
                packed, depths, radii, num_tiles_hit, img_height, img_width, block_width, n_samples,
                rolling_shutter_time, exposure_time)
        else:
            num_intersects, cum_tiles_hit = compute_cumulative_intersects(num_tiles_hit)

        if num_intersects < 1:
            # reference behaviour for an empty render (rasterize.py:136-144): background image, final_Ts = 0
            out_img = torch.ones(img_height, img_width, colors.shape[-1], device=xys.device) * background
            gaussian_ids_sorted = torch.zeros(0, 1, device=xys.device)
            tile_bins = torch.zeros(0, 2, device=xys.device)
            final_Ts = torch.zeros(img_height, img_width, blur_samples, device=xys.device)
            final_idx = torch.zeros(img_height, img_width, blur_samples, device=xys.device)
        elif three:
            out_img, final_Ts, final_idx, out_alpha = _C.blend_forward_packed(
                img_height, img_width, block_width, blur_samples, gaussian_ids_sorted, tile_bins, packed,
                rolling_shutter_time, exposure_time, background, want_alpha=True)
        else:
            gaussian_ids_sorted, tile_bins = _C.bin_tiles(num_intersects, xys, depths, radii, num_tiles_hit,
                                                          tile_bounds, block_width)
            out_img, final_Ts, final_idx = _C.nd_rasterize_forward(
                tile_bounds, block, img_size, blur_samples, gaussian_ids_sorted, tile_bins, xys, pix_vels,
                rolling_shutter_time, exposure_time, conics, colors, opacity, background)

        ctx.img_width = img_width
        ctx.img_height = img_height
        ctx.num_intersects = num_intersects
        ctx.block_width = block_width
        ctx.blur_samples = blur_samples
        ctx.rolling_shutter_time = rolling_shutter_time
        ctx.exposure_time = exposure_time
        ctx.has_packed = packed is not None
        ctx.save_for_backward(gaussian_ids_sorted, tile_bins, xys, pix_vels, conics, colors, opacity, background,
                              final_Ts, final_idx, packed if packed is not None else background)
        ctx.set_materialize_grads(False)  # an unused output's cotangent arrives as None instead of a zero image
        if return_alpha:
            if out_alpha is None:
                final_T_mean = final_Ts.mean(dim=-1) if final_Ts.dim() == 3 else final_Ts
                out_alpha = 1 - final_T_mean
            return out_img, out_alpha
        return out_img

    @staticmethod
    def backward(ctx, v_out_img, v_out_alpha=None):
        (gaussian_ids_sorted, tile_bins, xys, pix_vels, conics, colors, opacity, background, final_Ts,
         final_idx, packed) = ctx.saved_tensors
        if v_out_img is None:  # only alpha was used downstream
            v_out_img = torch.zeros(ctx.img_height, ctx.img_width, colors.shape[-1], dtype=torch.float32, device=xys.device)
        if v_out_alpha is None and not ctx.has_packed:
            v_out_alpha = torch.zeros_like(v_out_img[..., 0])

        if ctx.num_intersects < 1:
            v_xy = torch.zeros_like(xys)
            v_xy_abs = torch.zeros_like(xys)
            v_pix_vels = torch.zeros_like(pix_vels)
            v_conic = torch.zeros_like(conics)
            v_colors = torch.zeros_like(colors)
            v_opacity = torch.zeros_like(opacity)
        elif ctx.has_packed:
            v_xy, v_xy_abs, v_pix_vels, v_conic, v_colors, v_opacity = _C.blend_backward_packed(
                xys.size(0), ctx.img_height, ctx.img_width, ctx.block_width, ctx.blur_samples, gaussian_ids_sorted,
                tile_bins, packed, ctx.rolling_shutter_time, ctx.exposure_time, background, final_Ts, final_idx,
                v_out_img.contiguous(), v_out_alpha)  # None = zero alpha cotangent, no (H,W) buffer is read
            v_opacity = v_opacity.reshape(opacity.shape)
        else:
            v_xy, v_xy_abs, v_pix_vels, v_conic, v_colors, v_opacity = _C.nd_rasterize_backward(
                ctx.img_height, ctx.img_width, ctx.block_width, ctx.blur_samples, gaussian_ids_sorted, tile_bins,
                xys, pix_vels, ctx.rolling_shutter_time, ctx.exposure_time, conics, colors, opacity, background,
                final_Ts, final_idx, v_out_img.contiguous(), v_out_alpha.contiguous())
            v_opacity = v_opacity.reshape(opacity.shape)

        v_background = None
        if background.requires_grad:
            final_T_mean = final_Ts.mean(dim=-1) if final_Ts.dim() == 3 else final_Ts
            v_background = torch.matmul(
                v_out_img.float().reshape(-1, background.shape[0]).t(), final_T_mean.float().reshape(-1, 1)).squeeze()

        xys.absgrad = v_xy_abs  # AbsGS split criterion side channel (rasterize.py:272-275)

        return (v_xy, None, v_pix_vels, None, v_conic, None, v_colors, v_opacity, None, None, None, v_background,
                None, None, None, None)
