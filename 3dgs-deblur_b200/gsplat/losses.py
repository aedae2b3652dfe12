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


class _L1Loss(Function):
    @staticmethod
    def forward(ctx, pred, target, gamma=None):
        _lib.require_cuda(pred, target)
        if pred.shape != target.shape:
            raise ValueError(f"l1_loss: shapes differ: {tuple(pred.shape)} vs {tuple(target.shape)}")
        if pred.dtype != torch.float32 or target.dtype != torch.float32:
            raise RuntimeError("l1_loss: expected float32 tensors")
        if pred.numel() < 1:
            raise ValueError("l1_loss: empty input")
        pred_c, target_c = pred.contiguous(), target.contiguous()
        dev = pred.device
        with _lib.on_device(dev):
            loss = torch.empty((), dtype=torch.float32, device=dev)
            grad = torch.empty_like(pred_c) if ctx.needs_input_grad[0] else None
            if gamma is None:
                check(_lib.load().b200_l1_loss(pred_c.numel(), ptr(pred_c), ptr(target_c), ptr(loss), ptr(grad),
                                               ptr(_workspace(dev)), 1, stream()))
            else:  # pred is the LINEAR render: gamma correction and its backward ride in the same kernel
                if not float(gamma) > 0.0:
                    raise ValueError("l1_loss: gamma must be positive")
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
        return (g * v_loss if g is not None else None), None, None


def l1_loss(pred: torch.Tensor, target: torch.Tensor, gamma: float = None) -> torch.Tensor:
    """mean |pred - target| as a 0-d tensor; differentiable w.r.t. `pred` only.

    gamma (extension): `pred` is the LINEAR render and the loss is taken on the caller's gamma-corrected image,
    mean |clamp(pred, max=1) ** (1 / gamma) - target| (splatfacto.py:879-880 followed by :957) -- correction, loss and the
    cotangent w.r.t. the linear image in one kernel instead of clamp / pow forward and their three backward passes."""
    return _L1Loss.apply(pred, target, gamma)


def l1_loss_and_grad(pred: torch.Tensor, target: torch.Tensor, gamma: float = None):
    """(mean |pred - target|, d loss / d pred) from the one kernel, outside autograd -- for callers that chain the
    cotangent themselves (gsplat.dp.fused_shading_phase).  Same arguments and checks as `l1_loss`."""
    with torch.no_grad():
        ctx = _Ctx()
        loss = _L1Loss.forward(ctx, pred, target, gamma)
    return loss, ctx.grad


class _Ctx:
    needs_input_grad = (True, False, False)


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
    def forward(ctx, pred, target, lam):
        _check_images("photometric_loss" if lam is not None else "ssim", pred, target)
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
            if lam is not None:
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
        return grad, None, None


def ssim(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """pytorch_msssim.SSIM(data_range=1, size_average=True, channel=C) of two (H, W, C) images (splatfacto.py:260,958);
    differentiable w.r.t. `pred`."""
    return _Photometric.apply(pred, target, None)


def photometric_loss(pred: torch.Tensor, target: torch.Tensor, ssim_lambda: float = 0.2) -> torch.Tensor:
    """Splatfacto's main loss (splatfacto.py:957-975): (1 - ssim_lambda) * mean|target - pred| + ssim_lambda *
    (1 - SSIM(target, pred)); differentiable w.r.t. `pred`."""
    return _Photometric.apply(pred, target, float(ssim_lambda))
