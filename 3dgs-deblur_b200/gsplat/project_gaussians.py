"""`gsplat.project_gaussians` -- EWA projection of 3D Gaussians (operator surface of the reference's
gsplat/project_gaussians.py:14-345).

Unlike the reference there is a single CUDA path: when a camera velocity requires grad the reference
falls back to ~60 PyTorch ops (project_gaussians.py:81-112); here the same fused kernel runs and its
backward additionally produces dL/d(linear_velocity), dL/d(angular_velocity) and an exact
dL/d(viewmat).  In that mode the gradients follow the reference's torch path (clamp-aware, exact
rotation gradient); otherwise they follow its CUDA path (backward.cu:371-572 plus the approximate
viewmat block, project_gaussians.py:272-307).
"""
from typing import Optional, Tuple

import torch
from torch import Tensor
from torch.autograd import Function

import gsplat.cuda as _C
from gsplat import _lib


def project_gaussians(
    means3d: Tensor,
    scales: Tensor,
    glob_scale: float,
    quats: Tensor,
    linear_velocity: Optional[Tensor],
    angular_velocity: Optional[Tensor],
    rolling_shutter_time: float,
    exposure_time: float,
    viewmat: Tensor,
    fx: float,
    fy: float,
    cx: float,
    cy: float,
    img_height: int,
    img_width: int,
    block_width: int,
    clip_thresh: float = 0.01,
) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Returns (xys, depths, pix_vels, radii, conics, compensation, num_tiles_hit, cov3d).

    means3d (N,3); scales (N,3) (already exp'd); quats (N,4) normalised wxyz; linear/angular velocity
    (3,) or (1,3) in camera coordinates or None; viewmat (3,4) or (4,4) world-to-camera, row major."""
    assert block_width > 1 and block_width <= 16, "block_width must be between 2 and 16"
    if _lib.SYNC_CHECKS or not quats.is_cuda:
        assert (quats.norm(dim=-1) - 1 < 1e-6).all(), "quats must be normalized"

    if linear_velocity is None:
        assert angular_velocity is None
        assert rolling_shutter_time == 0
        v_lin = torch.zeros(3, dtype=means3d.dtype, device=means3d.device)
        v_ang = v_lin
    else:
        assert angular_velocity is not None
        v_lin, v_ang = linear_velocity, angular_velocity

    return _ProjectGaussians.apply(
        means3d.contiguous(), scales.contiguous(), glob_scale, quats.contiguous(), v_lin, v_ang,
        rolling_shutter_time, exposure_time, viewmat.contiguous(), fx, fy, cx, cy, img_height, img_width,
        block_width, clip_thresh,
    )


def _vel_dev(v: Tensor, device) -> Tensor:
    v = v.detach()
    if v.device == device and v.dtype == torch.float32 and v.numel() == 3 and v.is_contiguous():
        return v.view(3)  # common case: no copies, no launches
    return v.to(device=device, dtype=torch.float32).reshape(-1)[:3].contiguous()


class _ProjectGaussians(Function):
    @staticmethod
    def forward(ctx, means3d, scales, glob_scale, quats, linear_velocity, angular_velocity, rolling_shutter_time,
                exposure_time, viewmat, fx, fy, cx, cy, img_height, img_width, block_width, clip_thresh=0.01):
        num_points = means3d.shape[-2]
        if num_points < 1 or means3d.shape[-1] != 3:
            raise ValueError(f"Invalid shape for means3d: {means3d.shape}")
        dev = means3d.device
        lin, ang = _vel_dev(linear_velocity, dev), _vel_dev(angular_velocity, dev)
        # the normalisation assert of project_gaussians.py:69 is evaluated inside the kernel and raised at the next
        # host sync of the path (the intersection-count read-back in rasterize_gaussians); see gsplat/_lib.py
        flag = None if _lib.SYNC_CHECKS else _lib.new_quat_flag(dev)
        (cov3d, xys, depths, pix_vels, radii, conics, compensation, num_tiles_hit) = _C.project_gaussians_forward(
            num_points, means3d, scales, glob_scale, quats, None, None, rolling_shutter_time, exposure_time, viewmat,
            fx, fy, cx, cy, img_height, img_width, block_width, clip_thresh, _vel_tensors=(lin, ang), _quat_flag=flag)

        ctx.cfg = (num_points, glob_scale, fx, fy, cx, cy, img_height, img_width, rolling_shutter_time, exposure_time)
        ctx.vel_shapes = (linear_velocity.shape, angular_velocity.shape)
        ctx.vel_devs = (linear_velocity.device, angular_velocity.device)  # host-side camera velocities get host gradients
        ctx.vel_grad = bool(linear_velocity.requires_grad or angular_velocity.requires_grad)
        ctx.save_for_backward(means3d, scales, quats, viewmat, cov3d, radii, conics, compensation, lin, ang)
        ctx.mark_non_differentiable(radii, num_tiles_hit)
        return (xys, depths, pix_vels, radii, conics, compensation, num_tiles_hit, cov3d)

    @staticmethod
    def backward(ctx, v_xys, v_depths, v_pix_vels, v_radii, v_conics, v_compensation, v_num_tiles_hit, v_cov3d):
        means3d, scales, quats, viewmat, cov3d, radii, conics, compensation, lin, ang = ctx.saved_tensors
        num_points, glob_scale, fx, fy, cx, cy, H, W, rs_time, exposure = ctx.cfg
        want_vel = ctx.vel_grad and (ctx.needs_input_grad[4] or ctx.needs_input_grad[5])
        want_vm = bool(ctx.needs_input_grad[8])
        out = _C.project_gaussians_backward(
            num_points, means3d, scales, glob_scale, quats, None, None, rs_time, exposure, viewmat, fx, fy, cx, cy,
            H, W, cov3d, radii, conics, compensation, v_xys, v_depths, v_pix_vels, v_conics, v_compensation,
            _vel_tensors=(lin, ang), _exact=ctx.vel_grad, _want_vel=want_vel, _want_viewmat=want_vm, _want_cov=False)
        v_mean3d, v_scale, v_quat = out[2], out[3], out[4]
        rest = list(out[5:])
        v_lin = v_ang = v_viewmat = None
        if want_vel:
            g_lin, g_ang = rest.pop(0), rest.pop(0)
            v_lin = g_lin.reshape(ctx.vel_shapes[0]).to(ctx.vel_devs[0]) if ctx.needs_input_grad[4] else None
            v_ang = g_ang.reshape(ctx.vel_shapes[1]).to(ctx.vel_devs[1]) if ctx.needs_input_grad[5] else None
        if want_vm:
            g = rest.pop(0)  # (3,4)
            v_viewmat = torch.zeros_like(viewmat)
            v_viewmat[..., :3, :4] = g
        return (v_mean3d, v_scale, None, v_quat, v_lin, v_ang, None, None, v_viewmat,
                None, None, None, None, None, None, None, None)
