"""`gsplat.fused` -- the Splatfacto render block as ONE differentiable operator on raw parameters.

"Next" row f-1 of SURVEY.md section 8: the caller-side glue of nerfstudio/models/splatfacto.py:816-880
(exp(scales), quaternion normalisation, torch.cat of the SH tensors, SH -> clamp(rgb + 0.5, 0),
sigmoid(opacity) * compensation, alpha = 1 - mean(T)) costs ~35 small PyTorch kernels per step around the
rasterizer.  `render_gaussians` takes the raw parameter tensors, runs one fused pre-processing kernel, the culled
binning, the blend, and in backward one fused post-processing kernel.  It needs a caller change (hence "next");
`gsplat.project_gaussians / spherical_harmonics / rasterize_gaussians` stay the drop-in surface.

Outputs are the same as chaining the three drop-in operators the way Splatfacto does (tests/test_gpu_fused.py).
"""
from typing import Dict, Optional

import torch
from torch import Tensor
from torch.autograd import Function

import gsplat.cuda as _C
from gsplat import _lib
from gsplat._lib import check, ptr, stream


def render_gaussians(means: Tensor, log_scales: Tensor, quats: Tensor, opacity_logit: Tensor, sh_dc: Tensor,
                     sh_rest: Tensor, viewmat: Tensor, cam_pos: Tensor, linear_velocity: Optional[Tensor],
                     angular_velocity: Optional[Tensor], fx: float, fy: float, cx: float, cy: float, img_height: int,
                     img_width: int, block_width: int, background: Tensor, rolling_shutter_time: float = 0.0,
                     exposure_time: float = 0.0, blur_samples: int = 1, sh_degree_to_use: int = 3,
                     clip_thresh: float = 0.01, grad_sink: Optional[Dict[str, Tensor]] = None, return_depth: bool = False):
    """Returns (rgb (H,W,3), alpha (H,W), info).  info["radii"] (N,) int32; after backward info["absgrad"] (N,2) holds the
    abs-grad densification statistic (the `xys.absgrad` side channel of the drop-in operator).

    return_depth: also info["depth"] (H,W,1), the caller's eval depth image (splatfacto.py:881-897: a second, static
    rasterize call with the depths as colours over a zero background, divided by alpha; the largest rendered depth where
    alpha is 0) -- one more blend launch over the lists already built instead of a second projection-sized binning
    (lists are rebuilt only when the colour pass had rolling shutter or an even sample count with exposure: then they do
    not provably contain the static lists).  Not differentiable.

    grad_sink (optional): {"means", "log_scales", "quats", "opacity_logit", "sh_dc", "sh_rest"} -> preallocated
    gradient buffers (e.g. slices of a flat DP buffer).  When given, backward OVERWRITES them directly (every row,
    zeros for culled Gaussians) and autograd returns no gradient for those inputs -- no accumulate pass, no memset."""
    if linear_velocity is None:
        linear_velocity = torch.zeros(3, device=means.device)
        angular_velocity = torch.zeros(3, device=means.device)
    info: Dict[str, Tensor] = {}
    rgb, alpha = _FusedRender.apply(means, log_scales, quats, opacity_logit, sh_dc, sh_rest, viewmat, cam_pos,
                                    linear_velocity, angular_velocity, background,
                                    (float(fx), float(fy), float(cx), float(cy), int(img_height), int(img_width),
                                     int(block_width), float(rolling_shutter_time), float(exposure_time),
                                     int(blur_samples), int(sh_degree_to_use), float(clip_thresh)), grad_sink, info)
    lists = info.pop("_lists")
    if return_depth:
        info["depth"] = _depth_pass(lists, alpha, int(img_height), int(img_width), int(block_width), int(blur_samples),
                                    float(rolling_shutter_time), float(exposure_time))
    return rgb, alpha, info


def _depth_pass(lists, alpha, H, W, bw, S, rs, ex):
    packed, depths, radii, nth, ids, bins, total = lists
    dev = depths.device
    with torch.no_grad(), _lib.on_device(dev):
        if total < 1:  # the reference's empty-render branch returns the (zero) background
            depth_im = torch.zeros(H, W, 1, device=dev)
        else:
            rec = _C.set_record_colors(packed.clone(), depths[:, None].expand(-1, 3).contiguous())
            if not (rs == 0 and (ex == 0 or S % 2 == 1)):  # see gsplat.rasterize: the static lists may hold other pairs
                _, ids, bins = _C.bin_cull(rec, depths, radii, nth, H, W, bw, 1, 0.0, 0.0)
            img, _, _ = _C.blend_forward_packed(H, W, bw, 1, ids, bins, rec, 0.0, 0.0, torch.zeros(3, device=dev))
            depth_im = img[..., 0:1]
        a = alpha.detach()[..., None]
        return torch.where(a > 0, depth_im / a, depth_im.max())


class _FusedRender(Function):
    @staticmethod
    def forward(ctx, means, log_scales, quats, opacity_logit, sh_dc, sh_rest, viewmat, cam_pos, lin, ang, background,
                cfg, grad_sink, info):
        fx, fy, cx, cy, H, W, bw, rs, ex, S, deg, clip = cfg
        if not (0 < S <= 10):
            raise RuntimeError("unsupported blur size")
        dev = means.device
        lib = _lib.load()
        n = means.shape[0]
        K = sh_rest.shape[1] + 1
        tens = [t.contiguous() for t in (means, log_scales, quats, opacity_logit, sh_dc, sh_rest, viewmat, cam_pos)]
        means, log_scales, quats, opacity_logit, sh_dc, sh_rest, viewmat, cam_pos = tens
        _lib.require_cuda(*tens)
        if lin.numel() < 3 or ang.numel() < 3:
            raise ValueError("render_gaussians: camera velocities need 3 components each")
        # velocities may live on the host (the reference's camera code builds them there): the kernels take device pointers
        lin_d = lin.detach().to(device=dev, dtype=torch.float32).reshape(-1)[:3].contiguous()
        ang_d = ang.detach().to(device=dev, dtype=torch.float32).reshape(-1)[:3].contiguous()
        with _lib.on_device(dev):
            packed = torch.empty((n * lib.b200_packed_record_bytes(),), dtype=torch.uint8, device=dev)
            depths = torch.empty((n,), dtype=torch.float32, device=dev)
            radii = torch.empty((n,), dtype=torch.int32, device=dev)
            nth = torch.empty((n,), dtype=torch.int32, device=dev)
            check(lib.b200_fused_preprocess_forward(
                n, ptr(means), ptr(log_scales), ptr(quats), ptr(opacity_logit), ptr(sh_dc), ptr(sh_rest), K, deg,
                ptr(viewmat), ptr(cam_pos), ptr(lin_d), ptr(ang_d), rs, ex, fx, fy, cx, cy, H, W, bw, clip, ptr(packed),
                ptr(depths), ptr(radii), ptr(nth), stream()))
            total, ids, bins = _C.bin_cull(packed, depths, radii, nth, H, W, bw, S, rs, ex)
            bg = background.contiguous()
            if total < 1:  # reference behaviour for an empty render (rasterize.py:136-144)
                rgb = torch.ones(H, W, 3, device=dev) * bg
                Ts = torch.zeros(H, W, S, device=dev)
                fi = torch.zeros(H, W, S, dtype=torch.int32, device=dev)
                alpha = 1 - Ts.mean(dim=-1)
            else:
                rgb, Ts, fi, alpha = _C.blend_forward_packed(H, W, bw, S, ids, bins, packed, rs, ex, bg, want_alpha=True)
        ctx.cfg, ctx.K, ctx.total = cfg, K, total
        ctx.grad_sink, ctx.info = grad_sink, info
        ctx.vel_shapes = (lin.shape, ang.shape)
        ctx.vel_devs = (lin.device, ang.device)
        ctx.save_for_backward(means, log_scales, quats, opacity_logit, sh_dc, sh_rest, viewmat, cam_pos, lin_d, ang_d, bg,
                              packed, radii, ids, bins, Ts, fi)
        info["radii"] = radii
        info["_lists"] = (packed, depths, radii, nth, ids, bins, total)  # for render_gaussians' optional depth pass
        ctx.mark_non_differentiable(radii)
        ctx.set_materialize_grads(False)  # an unused alpha (or rgb) arrives as None instead of a zero image
        return rgb, alpha

    @staticmethod
    def backward(ctx, v_rgb, v_alpha):
        (means, log_scales, quats, opacity_logit, sh_dc, sh_rest, viewmat, cam_pos, lin_d, ang_d, bg, packed, radii, ids,
         bins, Ts, fi) = ctx.saved_tensors
        fx, fy, cx, cy, H, W, bw, rs, ex, S, deg, clip = ctx.cfg
        dev = means.device
        lib = _lib.load()
        n, K = means.shape[0], ctx.K
        f32 = dict(dtype=torch.float32, device=dev)
        if v_rgb is None:
            v_rgb = torch.zeros(H, W, 3, **f32)
        sink = ctx.grad_sink or {}
        with _lib.on_device(dev):
            if ctx.total < 1:
                v_xy = torch.zeros(n, 2, **f32)
                v_abs, v_pix, v_conic, v_col, v_op = torch.zeros(n, 2, **f32), torch.zeros(n, 2, **f32), torch.zeros(n, 3, **f32), torch.zeros(n, 3, **f32), torch.zeros(n, 1, **f32)
            else:
                v_xy, v_abs, v_pix, v_conic, v_col, v_op = _C.blend_backward_packed(
                    n, H, W, bw, S, ids, bins, packed, rs, ex, bg, Ts, fi, v_rgb.contiguous(), v_alpha)
            ctx.info["absgrad"] = v_abs
            out = {}
            for name, like in (("means", means), ("log_scales", log_scales), ("quats", quats),
                               ("opacity_logit", opacity_logit), ("sh_dc", sh_dc), ("sh_rest", sh_rest)):
                out[name] = sink[name] if name in sink else torch.empty_like(like)
                assert out[name].is_contiguous() and out[name].numel() == like.numel()
            want_vel = ctx.needs_input_grad[8] or ctx.needs_input_grad[9]
            want_vm = ctx.needs_input_grad[6]
            g_lin = torch.empty(3, **f32) if want_vel else None
            g_ang = torch.empty(3, **f32) if want_vel else None
            g_vm = torch.empty(3, 4, **f32) if want_vm else None
            check(lib.b200_fused_preprocess_backward(
                n, ptr(means), ptr(log_scales), ptr(quats), ptr(opacity_logit), ptr(sh_dc), ptr(sh_rest), K, deg,
                ptr(viewmat), ptr(cam_pos), ptr(lin_d), ptr(ang_d), rs, ex, fx, fy, cx, cy, H, W, bw, clip, ptr(packed),
                ptr(radii), ptr(v_xy), ptr(v_pix), ptr(v_conic), ptr(v_col), ptr(v_op), ptr(out["means"]),
                ptr(out["log_scales"]), ptr(out["quats"]), ptr(out["opacity_logit"]), ptr(out["sh_dc"]), ptr(out["sh_rest"]),
                ptr(g_lin), ptr(g_ang), ptr(g_vm), stream()))
        ret = [None if name in sink else out[name].view_as(like)
               for name, like in (("means", means), ("log_scales", log_scales), ("quats", quats),
                                  ("opacity_logit", opacity_logit), ("sh_dc", sh_dc), ("sh_rest", sh_rest))]
        v_viewmat = None
        if want_vm:
            v_viewmat = torch.zeros_like(viewmat)
            v_viewmat[..., :3, :4] = g_vm
        v_lin = g_lin.reshape(ctx.vel_shapes[0]).to(ctx.vel_devs[0]) if (want_vel and ctx.needs_input_grad[8]) else None
        v_ang = g_ang.reshape(ctx.vel_shapes[1]).to(ctx.vel_devs[1]) if (want_vel and ctx.needs_input_grad[9]) else None
        v_bg = None
        if ctx.needs_input_grad[10]:
            v_bg = torch.matmul(v_rgb.float().reshape(-1, 3).t(), Ts.mean(dim=-1).float().reshape(-1, 1)).squeeze()
        return (*ret, v_viewmat, None, v_lin, v_ang, v_bg, None, None, None)
