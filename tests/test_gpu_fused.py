"""GPU: gsplat.fused.render_gaussians (raw parameters, one operator) against the chain of drop-in
operators Splatfacto uses (project_gaussians -> spherical_harmonics -> rasterize_gaussians), forward and backward."""
import pytest
import torch

pytestmark = pytest.mark.gpu



def _raw_scene(name, n, H, W, S, rs, ex, seed=0):
    import gsplat.synthetic as synthetic

    sc = synthetic.make_scene(name, device="cuda", n_override=n)
    cam = sc["cameras"][0]
    sc.update(H=H, W=W, blur_samples=S, rolling_shutter_time=rs, exposure_time=ex)
    cam.update(fx=W / 2.0, fy=W / 2.0, cx=W / 2.0, cy=H / 2.0)
    return sc, cam


def _leaves(sc, cam):
    names = ("means", "log_scales", "quats", "opacity_logit", "sh_dc", "sh_rest")
    lv = {k: sc[k].clone().requires_grad_(True) for k in names}
    lv["lin"] = cam["lin_vel"].clone().requires_grad_(True)
    lv["ang"] = cam["ang_vel"].clone().requires_grad_(True)
    lv["viewmat"] = cam["viewmat"].clone().requires_grad_(True)
    lv["bg"] = sc["background"].clone().requires_grad_(True)
    return lv


def _chain(sc, cam, lv):
    from gsplat import project_gaussians, rasterize_gaussians, spherical_harmonics

    H, W = sc["H"], sc["W"]
    q = lv["quats"] / lv["quats"].norm(dim=-1, keepdim=True)
    xys, depths, pv, radii, conics, comp, nth, _ = project_gaussians(
        lv["means"], torch.exp(lv["log_scales"]), 1, q, lv["lin"], lv["ang"], sc["rolling_shutter_time"], sc["exposure_time"],
        lv["viewmat"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], H, W, 16)
    colors = torch.cat((lv["sh_dc"], lv["sh_rest"]), dim=1)
    rgbs = torch.clamp(spherical_harmonics(3, lv["means"].detach() - cam["cam_pos"], colors) + 0.5, min=0.0)
    opac = torch.sigmoid(lv["opacity_logit"]) * comp[:, None]
    S = sc["blur_samples"] if sc["exposure_time"] > 0 else 1
    return rasterize_gaussians(xys, depths, pv, radii, conics, nth, rgbs, opac, H, W, 16, background=lv["bg"],
                               return_alpha=True, rolling_shutter_time=sc["rolling_shutter_time"],
                               exposure_time=sc["exposure_time"], blur_samples=S)


def _fused(sc, cam, lv, sink=None):
    from gsplat.fused import render_gaussians

    S = sc["blur_samples"] if sc["exposure_time"] > 0 else 1
    return render_gaussians(lv["means"], lv["log_scales"], lv["quats"], lv["opacity_logit"], lv["sh_dc"], lv["sh_rest"],
                            lv["viewmat"], cam["cam_pos"], lv["lin"], lv["ang"], cam["fx"], cam["fy"], cam["cx"], cam["cy"],
                            sc["H"], sc["W"], 16, lv["bg"], rolling_shutter_time=sc["rolling_shutter_time"],
                            exposure_time=sc["exposure_time"], blur_samples=S, sh_degree_to_use=3, grad_sink=sink)


def _rel(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("name,n,H,W,S,rs,ex", [("c2", 40000, 256, 320, 5, 0.0, 1 / 60), ("c2", 40000, 192, 256, 3, 1 / 50, 1 / 60),
                                                ("c1", 10000, 256, 256, 1, 0.0, 0.0)])
def test_fused_render_matches_operator_chain(name, n, H, W, S, rs, ex):
    sc, cam = _raw_scene(name, n, H, W, S, rs, ex)
    g = torch.Generator(device="cuda").manual_seed(1)
    v_rgb = torch.randn(H, W, 3, device="cuda", generator=g)
    v_a = torch.randn(H, W, device="cuda", generator=g)
    a = _leaves(sc, cam)
    rgb_a, alpha_a = _chain(sc, cam, a)
    ((rgb_a * v_rgb).sum() + (alpha_a * v_a).sum()).backward()
    b = _leaves(sc, cam)
    rgb_b, alpha_b, info = _fused(sc, cam, b)
    ((rgb_b * v_rgb).sum() + (alpha_b * v_a).sum()).backward()
    # forward: same kernels, same lists; only sigmoid (fast exp) and the order of the SH dot product differ
    # (a 1-ulp change of an opacity can flip one alpha >= 1/255 test: <= 1e-3 of the pixels may move by up to 1e-2)
    for x, y in ((rgb_a, rgb_b), (alpha_a, alpha_b)):
        diff = (x - y).abs()
        assert float((diff > 2e-5).float().mean()) <= 1e-3 and float(diff.max()) < 1e-2
    for k in ("means", "log_scales", "quats", "opacity_logit", "sh_dc", "sh_rest", "lin", "ang", "viewmat", "bg"):
        assert a[k].grad is not None and b[k].grad is not None, k
        assert _rel(b[k].grad, a[k].grad) < 3e-3, (k, _rel(b[k].grad, a[k].grad))
    assert info["absgrad"].shape == (n, 2) and info["radii"].shape == (n,)


def test_fused_grad_sink_writes_in_place_and_zeroes_culled_rows():
    sc, cam = _raw_scene("c2", 20000, 128, 160, 5, 0.0, 1 / 60)
    a = _leaves(sc, cam)
    rgb, alpha, _ = _fused(sc, cam, a)
    rgb.mean().backward()
    b = _leaves(sc, cam)
    sink = {k: torch.full_like(b[k], float("nan")) for k in ("means", "log_scales", "quats", "opacity_logit", "sh_dc", "sh_rest")}
    rgb2, alpha2, info = _fused(sc, cam, b, sink=sink)
    rgb2.mean().backward()
    for k, t in sink.items():
        assert torch.isfinite(t).all(), k  # every row overwritten, no NaN sentinel left
        assert b[k].grad is None  # autograd returned nothing for sunk inputs
        assert _rel(t, a[k].grad) < 1e-4, k
    culled = info["radii"] == 0
    assert culled.any() and (sink["sh_rest"][culled] == 0).all() and (sink["means"][culled] == 0).all()


def test_fused_trainer_matches_unfused_trainer():
    """Three optimizer steps through gsplat.dp with and without the fused path end at the same parameters."""
    import gsplat.synthetic as synthetic
    from gsplat.dp import FlatGaussians, ImageShardedTrainer

    outs = []
    for fused in (False, True):
        sc = synthetic.make_scene("c2", device="cuda", n_override=20000, n_cameras=2)
        sc.update(H=128, W=160)
        cams = []
        for c in sc["cameras"]:
            c.update(fx=80.0, fy=80.0, cx=80.0, cy=64.0, vel0=torch.cat([c["lin_vel"], c["ang_vel"]]))
            c["target"] = c["target"][:128, :160].contiguous()
            cams.append(c)
        model = FlatGaussians(sc, "cuda", n_cameras=2, optimize_velocities=True)
        tr = ImageShardedTrainer(model, sc, lr=1e-3, fused=fused)
        for k in range(3):
            tr.train_step(cams[k % 2], cams[k % 2]["target"], k % 2)
        outs.append(model.flat.detach().clone())
    moved = (outs[0] - outs[1]).abs().max()
    assert float(moved) < 5e-4, float(moved)  # Adam normalises, so tiny gradient differences stay tiny steps


def test_flat_adam_matches_torch_adam_slice_by_slice():
    """gsplat.optim.FlatAdam (b200_adam_step) == torch.optim.Adam(eps=1e-15) over several steps, updating in unaligned
    slices (any offset into the flat buffers), folding a gradient scale, and clearing the gradient behind it."""
    from gsplat.optim import FlatAdam
    g = torch.Generator(device="cuda").manual_seed(11)
    n = 100_003
    flat = torch.randn(n, device="cuda", generator=g)
    ref = flat.clone().requires_grad_(True)
    grad = torch.zeros(n, device="cuda")
    opt_ref = torch.optim.Adam([ref], lr=3e-3, eps=1e-15)
    opt = FlatAdam(flat, grad, lr=3e-3, eps=1e-15)
    cuts = [0, 1, 6, 4099, 50_001, n]
    for step in range(5):
        gr = torch.randn(n, device="cuda", generator=g) * (10.0 ** (step - 2))
        gr[::7] = 0.0
        ref.grad = gr.clone() * 0.5
        opt_ref.step()
        grad.copy_(gr)
        opt.begin_step()
        for a, b in zip(cuts[:-1], cuts[1:]):
            opt.update(a, b, grad_scale=0.5, zero_grad=True)
        assert (grad == 0).all()
        torch.testing.assert_close(flat, ref.detach(), rtol=2e-6, atol=2e-7)
    torch.testing.assert_close(opt.exp_avg_sq, opt_ref.state[ref]["exp_avg_sq"], rtol=1e-5, atol=0)
    with pytest.raises(ValueError):
        opt.update(5, 3)
    with pytest.raises(RuntimeError):
        FlatAdam(torch.zeros(4), torch.zeros(4))


def test_trainer_with_flat_adam_matches_torch_adam():
    """gsplat.dp.ImageShardedTrainer: the b200 optimizer (FlatAdam, fused gradient clearing) and torch's fused Adam end
    at the same parameters after three steps of the drop-in path."""
    import gsplat.synthetic as synthetic
    from gsplat.dp import FlatGaussians, ImageShardedTrainer

    outs = []
    for optimizer in ("torch", "b200"):
        sc = synthetic.make_scene("c2", device="cuda", n_override=20000, n_cameras=2)
        sc.update(H=128, W=160)
        cams = []
        for c in sc["cameras"]:
            c.update(fx=80.0, fy=80.0, cx=80.0, cy=64.0, vel0=torch.cat([c["lin_vel"], c["ang_vel"]]))
            c["target"] = c["target"][:128, :160].contiguous()
            cams.append(c)
        model = FlatGaussians(sc, "cuda", n_cameras=2, optimize_velocities=True)
        tr = ImageShardedTrainer(model, sc, lr=1e-3, optimizer=optimizer)
        for k in range(3):
            tr.train_step(cams[k % 2], cams[k % 2]["target"], k % 2)
        outs.append(model.flat.detach().clone())
    moved = (outs[0] - outs[1]).abs().max()
    assert float(moved) < 5e-4, float(moved)


def _depth_chain(sc, cam, lv, reuse):
    """Splatfacto's eval render (splatfacto.py:860-897): colour pass, then depth as colours in a second, static call."""
    import os

    import gsplat.rasterize as R
    from gsplat import project_gaussians, rasterize_gaussians, spherical_harmonics

    R._last_lists.clear()
    os.environ["B200SPLAT_NO_LIST_REUSE"] = "0" if reuse else "1"
    try:
        with torch.no_grad():
            H, W = sc["H"], sc["W"]
            q = lv["quats"] / lv["quats"].norm(dim=-1, keepdim=True)
            xys, depths, pv, radii, conics, comp, nth, _ = project_gaussians(
                lv["means"], torch.exp(lv["log_scales"]), 1, q, lv["lin"], lv["ang"], sc["rolling_shutter_time"], sc["exposure_time"],
                lv["viewmat"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], H, W, 16)
            rgbs = torch.clamp(spherical_harmonics(3, lv["means"] - cam["cam_pos"], torch.cat((lv["sh_dc"], lv["sh_rest"]), dim=1)) + 0.5, min=0.0)
            opac = torch.sigmoid(lv["opacity_logit"]) * comp[:, None]
            S = sc["blur_samples"] if sc["exposure_time"] > 0 else 1
            rgb, alpha = rasterize_gaussians(xys, depths, pv, radii, conics, nth, rgbs, opac, H, W, 16, background=lv["bg"],
                                             return_alpha=True, rolling_shutter_time=sc["rolling_shutter_time"],
                                             exposure_time=sc["exposure_time"], blur_samples=S)
            cached = R._last_lists.get(xys.device.index)
            depth_im = rasterize_gaussians(xys, depths, pv, radii, conics, nth, depths[:, None].repeat(1, 3), opac, H, W, 16,
                                           background=torch.zeros(3, device="cuda"))[..., 0:1]
            reused = cached is not None and R._last_lists.get(xys.device.index) is cached  # a fresh binning replaces the entry
            alpha = alpha[..., None]
            depth = torch.where(alpha > 0, depth_im / alpha, depth_im.detach().max())
        return rgb, alpha, depth, depth_im, reused, dict(xys=xys, depths=depths, radii=radii, conics=conics, nth=nth, opac=opac)
    finally:
        os.environ.pop("B200SPLAT_NO_LIST_REUSE", None)


@pytest.mark.parametrize("S,rs,ex,expect_reuse", [(5, 0.0, 1 / 60, True), (4, 0.0, 1 / 60, False), (5, 1 / 50, 1 / 60, False),
                                                  (1, 0.0, 0.0, True)])
def test_depth_pass_list_reuse_and_fused_depth(S, rs, ex, expect_reuse):
    """The caller's static depth pass over the colour pass's lists (gsplat.rasterize list reuse, gsplat.fused
    return_depth): bit-identical to freshly built lists, equal to the oracle's static render of the depths."""
    import numpy as np

    from oracle import oracle as O

    n, H, W = 30000, 192, 256
    sc, cam = _raw_scene("c2", n, H, W, S, rs, ex)
    lv = {k: v.detach() for k, v in _leaves(sc, cam).items()}
    rgb_a, alpha_a, depth_a, dim_a, reused_a, proj = _depth_chain(sc, cam, lv, reuse=True)
    rgb_b, alpha_b, depth_b, dim_b, reused_b, _ = _depth_chain(sc, cam, lv, reuse=False)
    assert reused_a == expect_reuse and not reused_b
    assert torch.equal(dim_a, dim_b) and torch.equal(depth_a, depth_b) and torch.equal(rgb_a, rgb_b)
    # oracle: static blend of the depths over the reference's full lists
    c = lambda t: t.detach().cpu().numpy()
    b = O.bin_and_sort(c(proj["xys"]), c(proj["depths"]), c(proj["radii"]), c(proj["nth"]), H, W, 16)
    zero2 = np.zeros((n, 2), np.float32)
    img, _, _ = O.rasterize_forward(H, W, 16, 1, b["gaussian_ids_sorted"], b["tile_bins"], c(proj["xys"]), zero2, 0.0, 0.0,
                                    c(proj["conics"]), np.repeat(c(proj["depths"])[:, None], 3, 1).astype(np.float32), c(proj["opac"]),
                                    np.zeros(3, np.float32))
    diff = (dim_a[..., 0].cpu() - torch.from_numpy(img[..., 0])).abs() / max(1.0, float(dim_a.max()))
    assert float((diff > 2e-5).float().mean()) <= 1e-3 and float(diff.max()) < 1e-2
    # fused operator: same depth image from its own lists
    from gsplat.fused import render_gaussians
    with torch.no_grad():
        rgb_f, alpha_f, info = render_gaussians(lv["means"], lv["log_scales"], lv["quats"], lv["opacity_logit"], lv["sh_dc"], lv["sh_rest"],
                                                lv["viewmat"], cam["cam_pos"], lv["lin"], lv["ang"], cam["fx"], cam["fy"], cam["cx"],
                                                cam["cy"], H, W, 16, lv["bg"], rolling_shutter_time=rs, exposure_time=ex,
                                                blur_samples=S if ex > 0 else 1, sh_degree_to_use=3, return_depth=True)
    assert info["depth"].shape == (H, W, 1) and "_lists" not in info
    covered = alpha_a[..., 0] > 0.05
    dd = (info["depth"][..., 0] - depth_a[..., 0]).abs()[covered] / depth_a[..., 0][covered].clamp_min(1e-3)
    assert covered.any() and float((dd > 1e-3).float().mean()) <= 2e-3
