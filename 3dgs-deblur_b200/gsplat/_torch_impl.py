"""Small pure-PyTorch helpers callers import from the reference's module of the same name.

Splatfacto imports `quat_to_rotmat` from here (nerfstudio/models/splatfacto.py:31; used for the
covariance of split Gaussians).  The reference's torch re-implementations of the kernels are NOT
mirrored here: this package has no CPU / PyTorch compute path (the parity oracle lives in oracle/).
"""
import torch
import torch.nn.functional as F
from torch import Tensor


def normalized_quat_to_rotmat(quat: Tensor) -> Tensor:
    """(..., 4) unit quaternions (w, x, y, z) -> (..., 3, 3) rotation matrices."""
    if quat.shape[-1] != 4:
        raise AssertionError(quat.shape)
    w, x, y, z = torch.unbind(quat, dim=-1)
    xx, yy, zz = x * x, y * y, z * z
    xy, xz, yz = x * y, x * z, y * z
    wx, wy, wz = w * x, w * y, w * z
    rows = [
        1 - 2 * (yy + zz), 2 * (xy - wz), 2 * (xz + wy),
        2 * (xy + wz), 1 - 2 * (xx + zz), 2 * (yz - wx),
        2 * (xz - wy), 2 * (yz + wx), 1 - 2 * (xx + yy),
    ]
    return torch.stack(rows, dim=-1).reshape(quat.shape[:-1] + (3, 3))


def quat_to_rotmat(quat: Tensor) -> Tensor:
    """Like the reference's `_torch_impl.quat_to_rotmat`: normalises first."""
    if quat.shape[-1] != 4:
        raise AssertionError(quat.shape)
    return normalized_quat_to_rotmat(F.normalize(quat, dim=-1))
