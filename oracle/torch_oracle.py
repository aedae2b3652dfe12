"""Autograd restatement of the reference's *torch* projection path.  TEST INFRASTRUCTURE ONLY.

The reference switches `project_gaussians` to its pure-PyTorch implementation
whenever a camera velocity requires grad (/root/reference/gsplat/gsplat/
project_gaussians.py:81-112 -> _torch_impl.py:396-467), so the gradients a user
of the reference observes for means / scales / quats / linear+angular velocity /
viewmat in that mode are the exact autograd gradients of that function.  This
module restates the differentiable part of it (same formulas, own code) in a
dtype-agnostic way so tests can take fp64 "truth" gradients:

  * view transform + near-plane mask            _torch_impl.py:363-367
  * cov3d = (R S)(R S)^T                        _torch_impl.py:232-239
  * EWA with the 1.3*tan(fov) clamp and +0.3    _torch_impl.py:242-294
  * conic / radius                              _torch_impl.py:309-337
  * pixel mean with 1/(z+1e-6)                  _torch_impl.py:340-347
  * pixel velocity with 1/(z+1e-6)              _torch_impl.py:350-360

Pinned against the reference itself by tests/golden/make_golden.py (run in the
build container, where /root/reference is importable).
"""
import torch


def quat_to_rotmat(q):
    w, x, y, z = q.unbind(-1)
    return torch.stack(
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
         2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
         2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], dim=-1
    ).reshape(q.shape[:-1] + (3, 3))


def project(means, scales, glob_scale, quats, lin_vel, ang_vel, rs_time, exposure, viewmat,
            fx, fy, cx, cy, H, W, block_width, clip_thresh=0.01):
    """Returns dict(xys, depths, pix_vels, radii, conics, compensation, num_tiles_hit, cov3d, mask).

    Differentiable outputs: xys, depths, pix_vels, conics, compensation (cov3d too).
    Masked-out Gaussians get zeros like the reference (_torch_impl.py:443-450).
    """
    dt = means.dtype
    Wm = viewmat[:3, :3]
    tv = viewmat[:3, 3]
    p_view = means @ Wm.T + tv
    close = p_view[:, 2] < clip_thresh
    R = quat_to_rotmat(quats)
    M = R * (glob_scale * scales)[:, None, :]
    cov3d = M @ M.transpose(-1, -2)

    z = p_view[:, 2]
    rz = 1.0 / z
    limx = 1.3 * (0.5 * W / fx)
    limy = 1.3 * (0.5 * H / fy)
    tx = z * torch.clamp(p_view[:, 0] * rz, -limx, limx)
    ty = z * torch.clamp(p_view[:, 1] * rz, -limy, limy)
    rz2 = rz * rz
    O = torch.zeros_like(rz)
    J = torch.stack([fx * rz, O, -fx * tx * rz2, O, fy * rz, -fy * ty * rz2], -1).reshape(-1, 2, 3)
    T = J @ Wm
    cov2d = T @ cov3d @ T.transpose(-1, -2)
    det_orig = cov2d[:, 0, 0] * cov2d[:, 1, 1] - cov2d[:, 0, 1] ** 2
    a = cov2d[:, 0, 0] + 0.3
    b = cov2d[:, 0, 1]
    c = cov2d[:, 1, 1] + 0.3
    det = a * c - b * b
    comp = torch.sqrt(torch.clamp(det_orig / det, min=1e-10))
    det_ok = det != 0
    conic = torch.stack([c / det, -b / det, a / det], -1)
    mid = 0.5 * (a + c)
    disc = torch.sqrt(torch.clamp(mid * mid - det, min=0.1))
    radius = torch.ceil(3.0 * torch.sqrt(torch.maximum(mid + disc, mid - disc)))

    rw = 1.0 / (z + 1e-6)
    xys = torch.stack([p_view[:, 0] * rw * fx + cx, p_view[:, 1] * rw * fy + cy], -1)
    if rs_time > 0 or exposure > 0:
        tot = lin_vel.reshape(1, 3) + torch.cross(ang_vel.reshape(1, 3).expand_as(p_view), p_view, dim=-1)
        pvx = -fx * (tot[:, 0] - tot[:, 2] * p_view[:, 0] * rw) * rw
        pvy = -fy * (tot[:, 1] - tot[:, 2] * p_view[:, 1] * rw) * rw
        pix_vel = torch.stack([pvx, pvy], -1)
        radius = radius + pix_vel.norm(dim=-1) * 0.5 * (exposure + rs_time)
    else:
        pix_vel = torch.zeros_like(xys)

    tbx, tby = (W + block_width - 1) // block_width, (H + block_width - 1) // block_width
    tc = xys.detach() / block_width
    tr = (radius.detach() / block_width)[:, None]
    lo = (tc - tr).to(torch.int32)
    hi = (tc + tr).to(torch.int32) + 1
    lo = torch.stack([lo[:, 0].clamp(0, tbx), lo[:, 1].clamp(0, tby)], -1)
    hi = torch.stack([hi[:, 0].clamp(0, tbx), hi[:, 1].clamp(0, tby)], -1)
    area = (hi[:, 0] - lo[:, 0]) * (hi[:, 1] - lo[:, 1])
    mask = (area > 0) & (~close) & det_ok

    zf = torch.zeros((), dtype=dt)
    return dict(
        xys=torch.where(mask[:, None], xys, zf),
        depths=torch.where(mask, z, zf),
        pix_vels=pix_vel,
        radii=torch.where(mask, radius.detach().to(torch.int32), 0),
        conics=torch.where(mask[:, None], conic, zf),
        compensation=torch.where(mask, comp, zf),
        num_tiles_hit=torch.where(mask, area, 0),
        cov3d=torch.where(mask[:, None, None], cov3d, zf),
        mask=mask,
    )


def project_vjp(inputs, cotangents, **cfg):
    """Gradients of sum(<output, cotangent>) w.r.t. means, scales, quats, lin_vel, ang_vel, viewmat.

    inputs: dict of tensors (any float dtype); cotangents: dict with v_xys, v_depths,
    v_pix_vels, v_conics, v_compensation.  Returns dict of grads (same dtype).
    """
    leaves = {k: inputs[k].detach().clone().requires_grad_(True)
              for k in ("means", "scales", "quats", "lin_vel", "ang_vel", "viewmat")}
    out = project(leaves["means"], leaves["scales"], cfg["glob_scale"], leaves["quats"], leaves["lin_vel"],
                  leaves["ang_vel"], cfg["rs_time"], cfg["exposure"], leaves["viewmat"], cfg["fx"], cfg["fy"],
                  cfg["cx"], cfg["cy"], cfg["H"], cfg["W"], cfg["block_width"], cfg.get("clip_thresh", 0.01))
    loss = ((out["xys"] * cotangents["v_xys"]).sum() + (out["depths"] * cotangents["v_depths"]).sum()
            + (out["pix_vels"] * cotangents["v_pix_vels"]).sum() + (out["conics"] * cotangents["v_conics"]).sum()
            + (out["compensation"] * cotangents["v_compensation"]).sum())
    grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    res = {}
    for (k, leaf), g in zip(leaves.items(), grads):
        res["v_" + k] = torch.zeros_like(leaf) if g is None else g
    return res, out
