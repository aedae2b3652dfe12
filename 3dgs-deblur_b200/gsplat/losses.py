"""Fused photometric loss for the rasterizer path ("next" row f-3).

`l1_loss(pred, target)` == `torch.abs(target - pred).mean()` (the L1 term of nerfstudio/models/splatfacto.py:957) with
the cotangent `sign(pred - target) / numel` produced by the same kernel, so that the backward pass starts at the blend
kernel instead of walking sub / abs / mean through autograd.  CUDA only; there is no CPU fallback."""
import torch
from torch.autograd import Function

from . import _lib
from ._lib import check, ptr, stream

_ws = {}  # (device index, stream handle) -> zero-initialised scratch (the kernel leaves its ticket at zero)


def _workspace(dev):
    key = (dev.index, stream())
    ws = _ws.get(key)
    if ws is None:
        ws = torch.zeros(_lib.load().b200_l1_loss_ws_bytes(), dtype=torch.uint8, device=dev)
        _ws[key] = ws
    return ws


def _target_options(name, pred, target, background, min_rgb_level, mask):
    """Checks of the uint8-target form shared by l1_loss / photometric_loss -> (channels, background, min level, mask)."""
    if pred.dim() != 3 or pred.shape[-1] != 3 or target.dim() != 3 or target.shape[:2] != pred.shape[:2] or target.shape[2] not in (3, 4):
        raise ValueError(f"{name}: a uint8 target must be (H, W, 3) or (H, W, 4) for an (H, W, 3) render, got "
                         f"{tuple(target.shape)} / {tuple(pred.shape)}")
    ch = int(target.shape[2])
    if ch == 4:
        if background is None:
            raise ValueError(f"{name}: an RGBA target is composited over `background` (3 floats on the device)")
        background = background.detach().to(device=pred.device, dtype=torch.float32).reshape(-1).contiguous()
        if background.numel() != 3:
            raise ValueError(f"{name}: background must have 3 components")
    else:
        background = None  # (an RGB image is used as it is, splatfacto.py:919-923)
    if mask is not None:
        if tuple(mask.shape[:2]) != tuple(pred.shape[:2]) or mask.numel() != pred.shape[0] * pred.shape[1]:
            raise ValueError(f"{name}: mask must be (H, W) or (H, W, 1)")
        mask = mask.detach().to(device=pred.device, dtype=torch.float32).reshape(pred.shape[0], pred.shape[1]).contiguous()
    level = float(min_rgb_level)
    if level < 0:
        raise ValueError(f"{name}: min_rgb_level must be >= 0")
    return ch, background, level / 255.0, mask


class _L1Loss(Function):
    @staticmethod
    def forward(ctx, pred, target, gamma=None, background=None, min_rgb_level=0.0, mask=None, want_target=False):
        _lib.require_cuda(pred, target)
        if pred.dtype != torch.float32:
            raise RuntimeError("l1_loss: expected a float32 render")
        if pred.numel() < 1:
            raise ValueError("l1_loss: empty input")
        if gamma is not None and not float(gamma) > 0.0:
            raise ValueError("l1_loss: gamma must be positive")
        dev = pred.device
        pred_c, target_c = pred.contiguous(), target.contiguous()
        ctx.prepared_target = None
        if target.dtype == torch.uint8:
            # the dataset's image as it is stored: conversion, compositing, minimum level and mask ride in the loss kernel
            ch, bg, level, mk = _target_options("l1_loss", pred, target, background, min_rgb_level, mask)
            with _lib.on_device(dev):
                loss = torch.empty((), dtype=torch.float32, device=dev)
                grad = torch.empty_like(pred_c) if ctx.needs_input_grad[0] else None
                tout = torch.empty_like(pred_c) if want_target else None
                check(_lib.load().b200_l1_loss_u8(pred_c.shape[0] * pred_c.shape[1], ch, ptr(pred_c), ptr(target_c), ptr(bg), level,
                                                  float(gamma) if gamma is not None else 0.0, ptr(mk), ptr(loss), ptr(grad), ptr(tout),
                                                  ptr(_workspace(dev)), 1, stream()))
            ctx.prepared_target = tout
        else:
            if background is not None or mask is not None or float(min_rgb_level) > 0:
                raise ValueError("l1_loss: background / min_rgb_level / mask apply to a uint8 target (the dataset's image); a float "
                                 "target is taken as already prepared")
            if pred.shape != target.shape:
                raise ValueError(f"l1_loss: shapes differ: {tuple(pred.shape)} vs {tuple(target.shape)}")
            if target.dtype != torch.float32:
                raise RuntimeError("l1_loss: expected float32 tensors")
            with _lib.on_device(dev):
                loss = torch.empty((), dtype=torch.float32, device=dev)
                grad = torch.empty_like(pred_c) if ctx.needs_input_grad[0] else None
                if gamma is None:
                    check(_lib.load().b200_l1_loss(pred_c.numel(), ptr(pred_c), ptr(target_c), ptr(loss), ptr(grad),
                                                   ptr(_workspace(dev)), 1, stream()))
                else:  # pred is the LINEAR render: gamma correction and its backward ride in the same kernel
                    check(_lib.load().b200_l1_loss_gamma(pred_c.numel(), ptr(pred_c), ptr(target_c), float(gamma), ptr(loss), ptr(grad),
                                                         ptr(_workspace(dev)), 1, stream()))
        ctx.grad = grad
        ctx.target_needs = ctx.needs_input_grad[1]
        return loss

    @staticmethod
    def backward(ctx, v_loss):
        if ctx.target_needs:
            raise RuntimeError("l1_loss: the target image is a constant (no gradient)")
        g = ctx.grad
        ctx.grad = None
        return (g * v_loss if g is not None else None), None, None, None, None, None, None


def l1_loss(pred: torch.Tensor, target: torch.Tensor, gamma: float = None, *, background: torch.Tensor = None,
            min_rgb_level: float = 0.0, mask: torch.Tensor = None) -> torch.Tensor:
    """mean |pred - target| as a 0-d tensor; differentiable w.r.t. `pred` only.

    gamma (extension): `pred` is the LINEAR render and the loss is taken on the caller's gamma-corrected image,
    mean |clamp(pred, max=1) ** (1 / gamma) - target| (splatfacto.py:879-880 followed by :957) -- correction, loss and the
    cotangent w.r.t. the linear image in one kernel instead of clamp / pow forward and their three backward passes.

    target may be the dataset's uint8 image, (H, W, 3) or (H, W, 4): the caller's ground-truth preparation then runs inside
    the loss kernel -- / 255 (get_gt_img, splatfacto.py:900-910), RGBA composited over `background`
    (composite_with_background, :912-923), clamp(min=min_rgb_level / 255) (:952-953), and with `mask` ((H, W) or (H, W, 1))
    both images are multiplied by it before the mean (:957-964)."""
    return _L1Loss.apply(pred, target, gamma, background, min_rgb_level, mask, False)


def l1_loss_and_grad(pred: torch.Tensor, target: torch.Tensor, gamma: float = None, *, background: torch.Tensor = None,
                     min_rgb_level: float = 0.0, mask: torch.Tensor = None, return_target: bool = False):
    """(mean |pred - target|, d loss / d pred) from the one kernel, outside autograd -- for callers that chain the
    cotangent themselves (gsplat.dp.fused_shading_phase).  Same arguments and checks as `l1_loss`; with `return_target`
    (uint8 targets) the prepared float target comes back as a third value."""
    with torch.no_grad():
        ctx = _Ctx()
        loss = _L1Loss.forward(ctx, pred, target, gamma, background, min_rgb_level, mask, return_target)
    if return_target:
        return loss, ctx.grad, (ctx.prepared_target if ctx.prepared_target is not None else target)
    return loss, ctx.grad


l1_loss.accepts_uint8 = True  # gsplat.dp hands the dataset's uint8 image straight to such a loss


class _Ctx:
    needs_input_grad = (True, False, False, False, False, False, False)


# ---- SSIM and the full photometric loss (splatfacto.py:957-975) -------------------------------------------------

_ssim_ws = {}
_window = None


def _window_taps():
    """The 11 float32 taps exactly as pytorch_msssim builds them (_fspecial_gauss_1d(11, 1.5): float32 arange, exp,
    normalise), as a ctypes array the C ABI reads on the host."""
    global _window
    if _window is None:
        import ctypes
        coords = torch.arange(11, dtype=torch.float32) - 11 // 2
        g = torch.exp(-(coords ** 2) / (2 * 1.5 ** 2))
        g = g / g.sum()
        _window = (ctypes.c_float * 11)(*[float(x) for x in g])
    return _window


def _ssim_workspace(dev, H, W, C):
    key = (dev.index, stream(), H, W, C)
    ws = _ssim_ws.get(key)
    if ws is None:
        ws = torch.zeros(_lib.load().b200_ssim_ws_bytes(H, W, C), dtype=torch.uint8, device=dev)
        _ssim_ws[key] = ws
    return ws


def _check_images(name, pred, target):
    _lib.require_cuda(pred, target)
    if pred.shape != target.shape or pred.dim() != 3:
        raise ValueError(f"{name}: expected two (H, W, C) images of equal shape, got {tuple(pred.shape)} / {tuple(target.shape)}")
    if pred.dtype != torch.float32 or target.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected float32 tensors")
    if pred.shape[0] < 11 or pred.shape[1] < 11:
        raise ValueError(f"{name}: SSIM needs an image of at least 11 x 11 pixels")


class _Photometric(Function):
    """(1 - lam) * L1 + lam * (1 - SSIM) in three kernels (L1, SSIM forward, SSIM backward + combine); lam = None
    returns the plain SSIM value instead."""

    @staticmethod
    def forward(ctx, pred, target, lam, background=None, min_rgb_level=0.0):
        name = "photometric_loss" if lam is not None else "ssim"
        u8 = target.dtype == torch.uint8
        if u8:
            _lib.require_cuda(pred, target)
            if pred.dtype != torch.float32:
                raise RuntimeError(f"{name}: expected a float32 render")
            ch, bg, level, _ = _target_options(name, pred, target, background, min_rgb_level, None)
            if pred.shape[0] < 11 or pred.shape[1] < 11:
                raise ValueError(f"{name}: SSIM needs an image of at least 11 x 11 pixels")
        else:
            if background is not None or float(min_rgb_level) > 0:
                raise ValueError(f"{name}: background / min_rgb_level apply to a uint8 target")
            _check_images(name, pred, target)
        pred_c, target_c = pred.contiguous(), target.contiguous()
        H, W, C = pred_c.shape
        dev, lib = pred.device, _lib.load()
        need = ctx.needs_input_grad[0]
        with _lib.on_device(dev):
            f32 = dict(dtype=torch.float32, device=dev)
            out = torch.empty((), **f32)
            ssim_val = torch.empty((), **f32)
            maps = torch.empty(lib.b200_ssim_maps_bytes(H, W, C) // 4, **f32) if need else None
            l1, g1 = None, None
            if u8:
                # the L1 kernel prepares the float target (conversion, compositing, minimum level) on its way; the SSIM term
                # is taken against what it wrote
                l1 = torch.empty((), **f32)
                g1 = torch.empty_like(pred_c) if (need and lam is not None) else None
                prepared = torch.empty_like(pred_c)
                check(lib.b200_l1_loss_u8(H * W, ch, ptr(pred_c), ptr(target_c), ptr(bg), level, 0.0, None, ptr(l1), ptr(g1),
                                          ptr(prepared), ptr(_workspace(dev)), 1, stream()))
                target_c = prepared
                if lam is None:
                    l1 = None
            elif lam is not None:
                l1 = torch.empty((), **f32)
                g1 = torch.empty_like(pred_c) if need else None
                check(lib.b200_l1_loss(pred_c.numel(), ptr(pred_c), ptr(target_c), ptr(l1), ptr(g1), ptr(_workspace(dev)), 1,
                                       stream()))
            check(lib.b200_ssim_forward(H, W, C, ptr(pred_c), ptr(target_c), ptr(maps), ptr(ssim_val),
                                        ptr(out) if lam is not None else None, ptr(l1), float(lam or 0.0), _window_taps(),
                                        ptr(_ssim_workspace(dev, H, W, C)), 1, stream()))
        ctx.save_for_backward(pred_c, target_c)
        ctx.maps, ctx.g1, ctx.lam = maps, g1, lam
        ctx.ssim_value = ssim_val
        return out if lam is not None else ssim_val

    @staticmethod
    def backward(ctx, v):
        if ctx.needs_input_grad[1]:
            raise RuntimeError("photometric_loss / ssim: the target image is a constant (no gradient)")
        pred_c, target_c = ctx.saved_tensors
        H, W, C = pred_c.shape
        lam = ctx.lam
        with _lib.on_device(pred_c.device):
            grad = ctx.g1 if ctx.g1 is not None else torch.empty_like(pred_c)  # the L1 cotangent is combined in place
            v = v.contiguous().float()
            if lam is None:   # d SSIM / d pred
                scale, add_in, add_scale = 1.0, None, 0.0
            else:             # (1 - lam) * sign / n - lam * d SSIM / d pred
                scale, add_in, add_scale = -float(lam), ctx.g1, 1.0 - float(lam)
            check(_lib.load().b200_ssim_backward(H, W, C, ptr(pred_c), ptr(target_c), ptr(ctx.maps), scale, ptr(add_in),
                                                 add_scale, ptr(v), _window_taps(), ptr(grad), stream()))
        ctx.maps = ctx.g1 = None
        return grad, None, None, None, None


def ssim(pred: torch.Tensor, target: torch.Tensor, *, background: torch.Tensor = None, min_rgb_level: float = 0.0) -> torch.Tensor:
    """pytorch_msssim.SSIM(data_range=1, size_average=True, channel=C) of two (H, W, C) images (splatfacto.py:260,958);
    differentiable w.r.t. `pred`.  A uint8 target is prepared as in `l1_loss`."""
    return _Photometric.apply(pred, target, None, background, min_rgb_level)


def photometric_loss(pred: torch.Tensor, target: torch.Tensor, ssim_lambda: float = 0.2, *, background: torch.Tensor = None,
                     min_rgb_level: float = 0.0) -> torch.Tensor:
    """Splatfacto's main loss (splatfacto.py:943-975): (1 - ssim_lambda) * mean|target - pred| + ssim_lambda *
    (1 - SSIM(target, pred)); differentiable w.r.t. `pred`.  `target` is the prepared float image or the dataset's uint8
    image, (H, W, 3) or (H, W, 4): then / 255, compositing over `background` and clamp(min=min_rgb_level / 255) run inside
    the L1 kernel, which hands the prepared target to the SSIM kernels.  (A mask multiplies BOTH images in the reference,
    :957-964: apply it to `pred` and to a float target before the call.)"""
    return _Photometric.apply(pred, target, float(ssim_lambda), background, min_rgb_level)


photometric_loss.accepts_uint8 = True
