"""Generate golden vectors from the REFERENCE implementation (run in the build container only).

    PYTHONPATH=/root/reference/gsplat python tests/golden/make_golden.py

Imports the reference's own pure-torch implementation (gsplat/_torch_impl.py -- the production
projection path when velocities need grad, and the oracle of the reference's own tests) and stores
small input/output fixtures as .npz next to this script.  The GPU box has no /root/reference, so
the tests only read the committed .npz files.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/gsplat")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3dgs-deblur_b200", "gsplat"))

from gsplat import _torch_impl as TI  # noqa: E402  (the reference)
import synthetic  # noqa: E402


def npz(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name), **{k: np.asarray(v) for k, v in arrs.items()})
    print("wrote", name, {k: np.asarray(v).shape for k, v in arrs.items()})


def projection_case(name, means, scales, glob_scale, quats, lv, av, rs, ex, viewmat4, fx, fy, cx, cy, H, W, seed):
    """Forward outputs + autograd VJP of the reference torch path for seeded random cotangents."""
    leaves = [t.clone().requires_grad_(True) for t in (means, scales, quats, lv, av, viewmat4)]
    m, s, q, l, a, vm = leaves
    out = TI.project_gaussians_forward(m, s, glob_scale, q, l, a, rs, ex, vm, (fx, fy, cx, cy), (W, H), 16, 0.01)
    cov3d, cov2d, xys, depths, pix_vel, radii, conic, comp, nth, mask = out
    g = torch.Generator().manual_seed(seed)
    N = means.shape[0]
    ct = dict(v_xys=torch.randn(N, 2, generator=g), v_depths=torch.randn(N, generator=g),
              v_pix_vels=torch.randn(N, 2, generator=g) * 0.01, v_conics=torch.randn(N, 3, generator=g),
              v_compensation=torch.randn(N, generator=g))
    # downstream (the blend) only ever sends gradient to Gaussians that were rasterised
    mf = mask.to(torch.float32)
    ct = {k: v * (mf[:, None] if v.ndim == 2 else mf) for k, v in ct.items()}
    loss = ((xys * ct["v_xys"]).sum() + (depths * ct["v_depths"]).sum() + (pix_vel * ct["v_pix_vels"]).sum()
            + (conic * ct["v_conics"]).sum() + (comp * ct["v_compensation"]).sum())
    grads = torch.autograd.grad(loss, leaves, allow_unused=True)
    grads = [torch.zeros_like(x) if gg is None else gg for x, gg in zip(leaves, grads)]
    npz(name, means=means, scales=scales, glob_scale=glob_scale, quats=quats, lin_vel=lv, ang_vel=av, rs_time=rs,
        exposure=ex, viewmat=viewmat4, fx=fx, fy=fy, cx=cx, cy=cy, H=H, W=W,
        cov3d=cov3d.detach(), cov2d=cov2d.detach(), xys=xys.detach(), depths=depths.detach(),
        pix_vels=pix_vel.detach(), radii=radii, conics=conic.detach(), compensation=comp.detach(),
        num_tiles_hit=nth, mask=mask, **{k: v for k, v in ct.items()},
        g_means=grads[0], g_scales=grads[1], g_quats=grads[2], g_lin_vel=grads[3], g_ang_vel=grads[4],
        g_viewmat=grads[5])


def main():
    # 1. the reference test's setting (gsplat/tests/test_project_gaussians.py:39-137): seed 42, N=100, 512^2,
    #    random rotation, non-zero velocities, rs=0.1, exposure=0.2 -- drawn on CPU in the same order.
    torch.manual_seed(42)
    N = 100
    means = torch.randn((N, 3))
    scales = torch.rand((N, 3)) + 0.2
    quats = torch.randn((N, 4))
    quats /= torch.linalg.norm(quats, dim=-1, keepdim=True)
    H, W = 512, 512
    viewmat = torch.eye(4)
    viewmat[2, 3] = 8.0
    viewmat[:3, :3] = TI.quat_to_rotmat(torch.randn(4))
    lv, av = torch.randn(3), torch.randn(3)
    projection_case("proj_seed42.npz", means, scales, 0.1, quats, lv, av, 0.1, 0.2, viewmat, W / 2, W / 2, W / 2,
                    H / 2, H, W, seed=1)
    # same without motion (the pix-velocity branch off)
    projection_case("proj_seed42_static.npz", means, scales, 0.1, quats, lv * 0, av * 0, 0.0, 0.0, viewmat, W / 2,
                    W / 2, W / 2, H / 2, H, W, seed=2)

    # 2. BASELINE config 1 (10k Gaussians, 256^2, 1 sample, zero velocity) and a 3k slice of config 2's
    #    distribution with motion on -- from the synthetic generator.
    for tag, cfg, n, motion in (("c1", "c1", None, False), ("c2s", "c2", 3000, True)):
        sc = synthetic.make_scene(cfg, n_override=n)
        cam = sc["cameras"][0]
        vm4 = torch.cat([cam["viewmat"], torch.tensor([[0.0, 0.0, 0.0, 1.0]])], 0)
        q = sc["quats"] / sc["quats"].norm(dim=-1, keepdim=True)
        lv = cam["lin_vel"] if motion else torch.zeros(3)
        av = cam["ang_vel"] if motion else torch.zeros(3)
        projection_case(f"proj_{tag}.npz", sc["means"], sc["log_scales"].exp(), 1.0, q, lv, av,
                        sc["rolling_shutter_time"] if motion else 0.0, sc["exposure_time"] if motion else 0.0, vm4,
                        cam["fx"], cam["fy"], cam["cx"], cam["cy"], sc["H"], sc["W"], seed=3)

    # 3. spherical harmonics, both methods, degree 4 coefficients, every degrees_to_use
    g = torch.Generator().manual_seed(5)
    n = 64
    dirs = torch.randn(n, 3, generator=g)  # un-normalised, like splatfacto.py:843
    coeffs = torch.randn(n, 25, 3, generator=g)
    sh = {}
    for method in ("poly", "fast"):
        for deg in range(5):
            k = (deg + 1) ** 2
            d = dirs / dirs.norm(dim=-1, keepdim=True)  # the CUDA kernel normalises in-kernel (sh.cuh:67-72)
            sh[f"{method}_{deg}"] = TI.compute_sh_color(d, coeffs[:, :k], method)
    npz("sh.npz", dirs=dirs, coeffs=coeffs, **sh)

    # 4. intersection mapping + bin edges from the reference's python loops (_torch_impl.py:470-527)
    torch.manual_seed(7)
    n, H, W, bw = 300, 96, 128, 16
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    xys = torch.rand(n, 2) * torch.tensor([W, H])
    depths = torch.rand(n) * 5 + 0.1
    radii = torch.randint(1, 25, (n,), dtype=torch.int32)
    tmin, tmax = TI.get_tile_bbox(xys, radii.float(), tb, bw)
    nth = ((tmax[:, 0] - tmin[:, 0]) * (tmax[:, 1] - tmin[:, 1])).to(torch.int32)
    assert (nth > 0).all()
    cum = torch.cumsum(nth, 0, dtype=torch.int32)
    isect, gids = TI.map_gaussian_to_intersects(n, xys, depths, radii, cum, tb, bw)
    isect_s, order = torch.sort(isect, stable=True)
    gids_s = gids[order]
    assert (isect_s[-1] >> 32) == (isect_s[-2] >> 32)  # the python loop mishandles a 1-entry last tile
    bins = TI.get_tile_bin_edges(int(cum[-1]), isect_s, tb)
    npz("map_bins.npz", xys=xys, depths=depths, radii=radii, num_tiles_hit=nth, cum_tiles_hit=cum, H=H, W=W, bw=bw,
        isect_ids=isect, gaussian_ids=gids, isect_ids_sorted=isect_s, gaussian_ids_sorted=gids_s, tile_bins=bins)

    # 5. cov2d bounds (gsplat/tests/test_cov2d_bounds.py:8-37)
    torch.manual_seed(9)
    L = torch.randn(100, 2, 2)
    cov = L @ L.transpose(-1, -2) + 0.1 * torch.eye(2)
    conic, radius, valid = TI.compute_cov2d_bounds(cov)
    npz("cov2d_bounds.npz", cov2d=torch.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 1, 1]], -1), conics=conic,
        radii=radius, valid=valid)


if __name__ == "__main__":
    main()
