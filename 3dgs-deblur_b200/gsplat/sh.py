"""`gsplat.sh` -- spherical-harmonics colours (operator surface of the reference's gsplat/sh.py:1-104)."""
from typing import Literal

from torch import Tensor
from torch.autograd import Function

import gsplat.cuda as _C

_BASES = {0: 1, 1: 4, 2: 9, 3: 16}
_DEGREE = {1: 0, 4: 1, 9: 2, 16: 3, 25: 4}


def num_sh_bases(degree: int):
    return _BASES.get(degree, 25)


def deg_from_sh(num_bases: int):
    assert num_bases in _DEGREE, "Invalid number of SH bases"
    return _DEGREE[num_bases]


def spherical_harmonics(degrees_to_use: int, viewdirs: Tensor, coeffs: Tensor,
                        method: Literal["poly", "fast"] = "fast") -> Tensor:
    """Colours (N,3) from un-normalised view directions (N,3) and coefficients (N,K,3).

    Differentiable w.r.t. `coeffs` only, like the reference (sh.py:45-46)."""
    assert coeffs.shape[-2] >= num_sh_bases(degrees_to_use)
    assert method in ["poly", "fast"]
    return _SphericalHarmonics.apply(method, degrees_to_use, viewdirs.contiguous(), coeffs.contiguous())


class _SphericalHarmonics(Function):
    @staticmethod
    def forward(ctx, method, degrees_to_use, viewdirs, coeffs):
        ctx.degrees_to_use = degrees_to_use
        ctx.degree = deg_from_sh(coeffs.shape[-2])
        ctx.method = method
        ctx.save_for_backward(viewdirs)
        return _C.compute_sh_forward(method, coeffs.shape[0], ctx.degree, degrees_to_use, viewdirs, coeffs)

    @staticmethod
    def backward(ctx, v_colors):
        (viewdirs,) = ctx.saved_tensors
        v_coeffs = _C.compute_sh_backward(ctx.method, v_colors.shape[0], ctx.degree, ctx.degrees_to_use, viewdirs,
                                          v_colors.contiguous())
        return None, None, None, v_coeffs
