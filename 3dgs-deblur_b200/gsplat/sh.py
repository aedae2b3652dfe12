"""`gsplat.sh` -- view-dependent colours from spherical-harmonics coefficients, behind the operator surface of the
reference's gsplat/sh.py:1-104 (`num_sh_bases`, `deg_from_sh`, `spherical_harmonics`)."""
from torch.autograd import Function

import gsplat.cuda as _C

_METHODS = ("poly", "fast")
_BASES_OF_DEGREE = (1, 4, 9, 16, 25)  # (degree + 1)^2 for degree 0 .. 4, the kernel maximum


def num_sh_bases(degree: int):
    """Number of coefficient triples of a degree-`degree` expansion; anything above 3 counts as the maximum, 4."""
    return _BASES_OF_DEGREE[degree] if 0 <= degree <= 3 else _BASES_OF_DEGREE[4]


def deg_from_sh(num_bases: int):
    assert num_bases in _BASES_OF_DEGREE, "Invalid number of SH bases"
    return _BASES_OF_DEGREE.index(num_bases)


_GRAD_SINK = None  # see coeff_grad_sink


class coeff_grad_sink:
    """Extension (not in the reference): `with coeff_grad_sink(buf):` makes the backward of every `spherical_harmonics`
    call issued inside the block write d loss / d coeffs straight into `buf` (a float32 tensor with coeffs' numel, e.g. the
    SH slice of a trainer's flat gradient buffer) and hand autograd nothing for `coeffs` -- the kernel writes every row
    (zeros for unused bases), so it REPLACES the buffer's contents: for a buffer that is zero before the backward pass
    that equals autograd's accumulate without the N*K*3 read-modify-write pass (58 us for 300k x 16 x 3 floats)."""

    def __init__(self, buf):
        self.buf = buf

    def __enter__(self):
        global _GRAD_SINK
        self.prev, _GRAD_SINK = _GRAD_SINK, self.buf
        return self

    def __exit__(self, *exc):
        global _GRAD_SINK
        _GRAD_SINK = self.prev
        return False


def spherical_harmonics(degrees_to_use, viewdirs, coeffs, method="fast"):
    """(N,3) colours from un-normalised view directions (N,3) and coefficients (N,K,3), K in {1,4,9,16,25}.

    Only the first `num_sh_bases(degrees_to_use)` bases are evaluated (Splatfacto ramps the degree up during training).
    `method` selects the basis formulation: "poly" (monomials) or "fast" (recurrences).  The gradient flows to `coeffs`
    only -- view directions are treated as constants, as in the reference (sh.py:45-46)."""
    assert coeffs.shape[-2] >= num_sh_bases(degrees_to_use)
    assert method in _METHODS
    return _SHColors.apply(method, degrees_to_use, viewdirs.contiguous(), coeffs.contiguous())


class _SHColors(Function):
    @staticmethod
    def forward(ctx, method, degrees_to_use, viewdirs, coeffs):
        degree = deg_from_sh(coeffs.shape[-2])
        ctx.cfg = (method, degree, degrees_to_use)
        ctx.sink = _GRAD_SINK
        if ctx.sink is not None and (ctx.sink.numel() != coeffs.numel() or not ctx.sink.is_contiguous()):
            raise ValueError("coeff_grad_sink: the sink must be a contiguous float32 buffer with coeffs' numel")
        ctx.save_for_backward(viewdirs)
        return _C.compute_sh_forward(method, coeffs.shape[0], degree, degrees_to_use, viewdirs, coeffs)

    @staticmethod
    def backward(ctx, grad_colors):
        method, degree, degrees_to_use = ctx.cfg
        (viewdirs,) = ctx.saved_tensors
        if ctx.sink is not None:  # the kernel writes into the caller's buffer; autograd gets nothing for coeffs
            _C.compute_sh_backward(method, grad_colors.shape[0], degree, degrees_to_use, viewdirs, grad_colors.contiguous(),
                                   out=ctx.sink)
            return None, None, None, None
        grad_coeffs = _C.compute_sh_backward(method, grad_colors.shape[0], degree, degrees_to_use, viewdirs,
                                             grad_colors.contiguous())
        return None, None, None, grad_coeffs  # method, degrees_to_use, viewdirs, coeffs
