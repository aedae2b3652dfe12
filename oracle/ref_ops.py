"""Autograd wrappers around the UNMODIFIED reference CUDA extension (oracle/_ref).  BENCH/TEST INFRASTRUCTURE ONLY.

/root/reference cannot travel to the GPU box and its Python sources may not be copied, so the reference's
operator layer is re-driven here: the same kernel sequence, allocations and host syncs as
gsplat/project_gaussians.py:139-345, gsplat/sh.py:62-104, gsplat/rasterize.py:102-294 and
gsplat/utils.py:106-182 (cumsum + .item(), map, torch.sort, torch.gather, bin edges), written from their
documented behaviour.  Used by `bench.py --impl refgpu` to time "the reference's own gsplat CUDA path"
(BASELINE.md B-gpu) beside libb200splat.  Velocities are constants here: with requires_grad velocities the
reference leaves CUDA for ~60 PyTorch ops (project_gaussians.py:81-112), which is slower still.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import build_ref  # noqa: E402

_EXT = None


def ext():
    global _EXT
    if _EXT is None:
        _EXT = build_ref.load_ref()
    return _EXT


class RefProject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3d, scales, glob_scale, quats, lin, ang, rs, exposure, viewmat, fx, fy, cx, cy, H, W, bw, clip):
        n = means3d.shape[0]
        lin_t, ang_t = tuple(lin.detach().reshape(-1).tolist()), tuple(ang.detach().reshape(-1).tolist())  # D2H sync, as the reference
        cov3d, xys, depths, pix_vels, radii, conics, comp, nth = ext().project_gaussians_forward(
            n, means3d, scales, glob_scale, quats, lin_t, ang_t, rs, exposure, viewmat, fx, fy, cx, cy, H, W, bw, clip)
        ctx.cfg = (n, glob_scale, lin_t, ang_t, rs, exposure, fx, fy, cx, cy, H, W)
        ctx.save_for_backward(means3d, scales, quats, viewmat, cov3d, radii, conics, comp)
        return xys, depths, pix_vels, radii, conics, comp, nth, cov3d

    @staticmethod
    def backward(ctx, v_xys, v_depths, v_pix_vels, v_radii, v_conics, v_comp, v_nth, v_cov3d):
        means3d, scales, quats, viewmat, cov3d, radii, conics, comp = ctx.saved_tensors
        n, gs, lin_t, ang_t, rs, exposure, fx, fy, cx, cy, H, W = ctx.cfg
        _, _, v_mean, v_scale, v_quat = ext().project_gaussians_backward(
            n, means3d, scales, gs, quats, lin_t, ang_t, rs, exposure, viewmat, fx, fy, cx, cy, H, W, cov3d, radii, conics,
            comp, v_xys.contiguous(), v_depths.contiguous(), v_pix_vels.contiguous(), v_conics.contiguous(), v_comp.contiguous())
        return (v_mean, v_scale, None, v_quat) + (None,) * 13


class RefSH(torch.autograd.Function):
    @staticmethod
    def forward(ctx, deg_use, viewdirs, coeffs):
        ctx.deg_use = deg_use
        ctx.save_for_backward(viewdirs)
        return ext().compute_sh_forward("fast", coeffs.shape[0], 3, deg_use, viewdirs, coeffs)

    @staticmethod
    def backward(ctx, v_colors):
        (viewdirs,) = ctx.saved_tensors
        return None, None, ext().compute_sh_backward("fast", v_colors.shape[0], 3, ctx.deg_use, viewdirs, v_colors.contiguous())


class RefRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xys, depths, pix_vels, radii, conics, nth, colors, opacity, H, W, bw, background, rs, exposure, S):
        n = xys.size(0)
        tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
        cum = torch.cumsum(nth, dim=0, dtype=torch.int32)
        m = cum[-1].item()  # host sync (utils.py:124)
        isect, gids = ext().map_gaussian_to_intersects(n, m, xys, depths, radii, cum, tb, bw)
        isect_s, order = torch.sort(isect)
        gids_s = torch.gather(gids, 0, order)
        bins = ext().get_tile_bin_edges(m, isect_s, tb)
        img, Ts, fi = ext().rasterize_forward(tb, (bw, bw, 1), (W, H, 1), S, gids_s, bins, xys, pix_vels, rs, exposure, conics,
                                              colors, opacity, background)
        ctx.cfg = (H, W, bw, S, rs, exposure)
        ctx.save_for_backward(gids_s, bins, xys, pix_vels, conics, colors, opacity, background, Ts, fi)
        return img, 1 - Ts.mean(dim=-1)

    @staticmethod
    def backward(ctx, v_img, v_alpha):
        gids_s, bins, xys, pix_vels, conics, colors, opacity, background, Ts, fi = ctx.saved_tensors
        H, W, bw, S, rs, exposure = ctx.cfg
        if v_alpha is None:
            v_alpha = torch.zeros_like(v_img[..., 0])
        v_xy, v_xy_abs, v_pix, v_conic, v_colors, v_opac = ext().rasterize_backward(
            H, W, bw, S, gids_s, bins, xys, pix_vels, rs, exposure, conics, colors, opacity, background, Ts, fi,
            v_img.contiguous(), v_alpha.contiguous())
        xys.absgrad = v_xy_abs
        return (v_xy, None, v_pix, None, v_conic, None, v_colors, v_opac) + (None,) * 7


def project_torch_path(means3d, scales, quats, lin, ang, rs, exposure, viewmat, fx, fy, cx, cy, H, W, bw):
    """The reference's projection when a camera velocity requires grad -- train.py's default (train.py:64-67): it leaves
    CUDA for ~60 PyTorch ops + autograd (project_gaussians.py:81-112 -> _torch_impl.py:396-467).  /root/reference cannot
    travel to the GPU box, so this drives oracle/torch_oracle.py, the restatement of that function that
    tests/test_oracle_golden.py pins against the reference's own outputs (same op sequence, same boolean-mask outputs)."""
    import torch_oracle as TO

    vm4 = torch.cat([viewmat.reshape(3, 4), torch.tensor([[0.0, 0.0, 0.0, 1.0]], device=viewmat.device)], 0)  # project_gaussians.py:84
    o = TO.project(means3d, scales, 1.0, quats, lin.reshape(-1), ang.reshape(-1), rs, exposure, vm4, fx, fy, cx, cy, H, W, bw)
    return (o["xys"].contiguous(), o["depths"].contiguous(), o["pix_vels"].contiguous(), o["radii"].to(torch.int32).contiguous(),
            o["conics"].contiguous(), o["compensation"].contiguous(), o["num_tiles_hit"].to(torch.int32).contiguous(), None)


def render(model, cam, scene, torch_projection=False):
    """The Splatfacto render block on the reference kernels (mirror of gsplat.dp.render).  torch_projection: velocities
    carry gradients, so the projection runs the reference's PyTorch path instead of its CUDA kernel."""
    p = model.params
    H, W, bw = scene["H"], scene["W"], scene["block_width"]
    quats = p["quats"] / p["quats"].norm(dim=-1, keepdim=True)
    assert (quats.norm(dim=-1) - 1 < 1e-6).all()  # project_gaussians.py:69
    if torch_projection:
        xys, depths, pix_vels, radii, conics, comp, nth, _ = project_torch_path(
            p["means"], torch.exp(p["log_scales"]), quats, cam["lin_vel"], cam["ang_vel"], scene["rolling_shutter_time"],
            scene["exposure_time"], cam["viewmat"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], H, W, bw)
    else:
        xys, depths, pix_vels, radii, conics, comp, nth, _ = RefProject.apply(
            p["means"], torch.exp(p["log_scales"]), 1.0, quats, cam["lin_vel"], cam["ang_vel"], scene["rolling_shutter_time"],
            scene["exposure_time"], cam["viewmat"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], H, W, bw, 0.01)
    colors = torch.cat((p["sh_dc"], p["sh_rest"]), dim=1)
    viewdirs = (p["means"].detach() - cam["cam_pos"]).contiguous()
    rgbs = torch.clamp(RefSH.apply(3, viewdirs, colors.contiguous()) + 0.5, min=0.0)
    opacities = torch.sigmoid(p["opacity_logit"]) * comp[:, None]
    S = scene["blur_samples"] if scene["exposure_time"] > 0 else 1
    rgb, alpha = RefRasterize.apply(xys, depths, pix_vels, radii, conics, nth, rgbs.contiguous(), opacities.contiguous(), H, W, bw,
                                    scene["background"], scene["rolling_shutter_time"], scene["exposure_time"], S)
    return rgb, alpha
