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
    def forward(ctx, pred, target):
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
            check(_lib.load().b200_l1_loss(pred_c.numel(), ptr(pred_c), ptr(target_c), ptr(loss), ptr(grad),
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
        return (g * v_loss if g is not None else None), None


def l1_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """mean |pred - target| as a 0-d tensor; differentiable w.r.t. `pred` only."""
    return _L1Loss.apply(pred, target)
