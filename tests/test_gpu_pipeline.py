"""GPU: the sync-free path -- capacity-mode tile lists (gsplat.rasterize.prepare_lists), the device-side optimizer veto
(gsplat.optim.FlatAdam device state), the SH gradient sink, and gsplat.dp.PipelinedTrainer (two CUDA graphs per camera
signature, exchange / SH update behind the next image's geometry) against the synchronising drop-in path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from util_scene import cu, grad_close, oracle_render, scene_np

if torch.cuda.is_available():
    from gsplat import project_gaussians, rasterize_gaussians, spherical_harmonics
    from gsplat.rasterize import prepare_lists


def _project(d):
    return project_gaussians(cu(d["means"]), cu(d["scales"]), 1.0, cu(d["quats"]), cu(d["lin_vel"]), cu(d["ang_vel"]), d["rs"],
                             d["exposure"], cu(d["viewmat"]), d["fx"], d["fy"], d["cx"], d["cy"], d["H"], d["W"], 16)


@pytest.mark.parametrize("name,n,H,W", [("c2", 30000, 192, 256), ("c2", None, None, None), ("c3_rs10", 100000, 360, 640)])
def test_prepared_lists_reproduce_the_synchronising_path_bit_for_bit(name, n, H, W):
    """prepare_lists (pack + capacity-mode binning, colours patched in later) + rasterize_gaussians(prepared=...) ==
    rasterize_gaussians alone: same image, alpha and gradients to the last bit (same lists in the same order, the padding
    behind them belongs to no tile)."""
    d = scene_np(name, n=n, H=H, W=W)
    xys, depths, pv, radii, conics, comp, nth, _ = _project(d)
    g = torch.Generator(device="cuda").manual_seed(3)
    col = torch.rand(d["N"], 3, device="cuda", generator=g)
    opac = cu(d["opacity"]) * comp[:, None]
    bg = cu(d["background"])
    kw = dict(background=bg, return_alpha=True, rolling_shutter_time=d["rs"], exposure_time=d["exposure"], blur_samples=d["S"])

    def run(prepared):
        leaves = [t.detach().clone().requires_grad_(True) for t in (xys, pv, conics, col, opac)]
        img, alpha = rasterize_gaussians(leaves[0], depths, leaves[1], radii, leaves[2], nth, leaves[3], leaves[4], d["H"],
                                         d["W"], 16, prepared=prepared, **kw)
        (img.square().sum() + alpha.sum()).backward()
        return img.detach(), alpha.detach(), [t.grad for t in leaves], leaves[0].absgrad

    ref = run(None)
    status = torch.zeros(4, dtype=torch.int32, device="cuda")
    # capacity: anything >= the entry count; an odd size exercises the padding
    prep = prepare_lists(xys, depths, pv, radii, conics, nth, opac, d["H"], d["W"], 16, d["rs"], d["exposure"], d["S"],
                         capacity=8 * d["N"] + 12345, status=status)
    out = run(prep)
    st = status.tolist()
    assert st[0] == 0 and 0 < st[1] <= 8 * d["N"] + 12345 and st[2] == st[1] and st[3] == int(nth.sum())
    assert torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])
    # gradients: same lists, same kernel -- only the float atomics' order varies between two launches
    for a, b in zip(out[2] + [out[3]], ref[2] + [ref[3]]):
        torch.testing.assert_close(a, b, rtol=2e-4, atol=1e-6 * float(b.abs().max()))
    # the padded tail belongs to no tile, the ranges cover exactly the real entries
    bins = prep.bins.cpu().numpy()
    assert int((bins[:, 1] - bins[:, 0]).sum()) == st[1] and int(bins[:, 1].max()) <= st[1]


def test_capacity_overflow_is_flagged_and_never_writes_out_of_bounds():
    d = scene_np("c2", n=30000, H=192, W=256)
    xys, depths, pv, radii, conics, comp, nth, _ = _project(d)
    opac = cu(d["opacity"]) * comp[:, None]
    status = torch.zeros(4, dtype=torch.int32, device="cuda")
    big = prepare_lists(xys, depths, pv, radii, conics, nth, opac, d["H"], d["W"], 16, d["rs"], d["exposure"], d["S"],
                        capacity=1 << 20, status=status)
    entries = int(status[1])
    status.zero_()
    guard = torch.full((entries + 4096,), -7, dtype=torch.int32, device="cuda")
    small = prepare_lists(xys, depths, pv, radii, conics, nth, opac, d["H"], d["W"], 16, d["rs"], d["exposure"], d["S"],
                          capacity=entries // 2, status=status)
    assert status.tolist()[:3] == [1, entries, entries]      # overflow raised, true count reported
    assert small.ids.numel() == entries // 2 and int(small.bins.max()) <= entries // 2
    assert int(small.ids.min()) >= 0 and int(small.ids.max()) < d["N"] and (guard == -7).all()
    # blending the truncated lists is memory-safe (the result is discarded by the trainer's veto)
    col = torch.rand(d["N"], 3, device="cuda")
    img = rasterize_gaussians(xys, depths, pv, radii, conics, nth, col, opac, d["H"], d["W"], 16, background=cu(d["background"]),
                              rolling_shutter_time=d["rs"], exposure_time=d["exposure"], blur_samples=d["S"], prepared=small)
    assert torch.isfinite(img).all()
    del big


def test_capacity_mode_empty_render_follows_the_reference_branch():
    """Nothing intersects a tile: the reference returns the background, final_Ts = 0 and hence alpha = 1
    (rasterize.py:136-144); in capacity mode the host never learns the count, the blend kernel takes that branch itself."""
    d = scene_np("c1", n=64)
    vm = d["viewmat"].copy()
    vm[2, 3] = -1e4
    out = project_gaussians(cu(d["means"]), cu(d["scales"]), 1.0, cu(d["quats"]), None, None, 0, 0, cu(vm), d["fx"], d["fy"],
                            d["cx"], d["cy"], d["H"], d["W"], 16)
    xys, depths, pv, radii, conics, comp, nth, _ = out
    assert int(nth.sum()) == 0
    status = torch.zeros(4, dtype=torch.int32, device="cuda")
    opac = cu(d["opacity"])
    prep = prepare_lists(xys, depths, pv, radii, conics, nth, opac, d["H"], d["W"], 16, capacity=4096, status=status)
    bg = cu(d["background"])
    col = torch.rand(64, 3, device="cuda", requires_grad=True)
    img, alpha = rasterize_gaussians(xys, depths, pv, radii, conics, nth, col, opac, d["H"], d["W"], 16, background=bg,
                                     return_alpha=True, prepared=prep)
    assert status.tolist() == [0, 0, 0, 0]
    assert torch.equal(img, bg.expand_as(img).contiguous()) and (alpha == 1).all()
    img.sum().backward()
    assert float(col.grad.abs().sum()) == 0.0


def test_sh_gradient_sink_equals_autograd_accumulation():
    from gsplat.sh import coeff_grad_sink

    g = torch.Generator(device="cuda").manual_seed(2)
    n = 5003
    dirs = torch.randn(n, 3, device="cuda", generator=g)
    coeffs = torch.randn(n, 16, 3, device="cuda", generator=g)
    v = torch.randn(n, 3, device="cuda", generator=g)
    a = coeffs.clone().requires_grad_(True)
    spherical_harmonics(2, dirs, a).backward(v)
    b = coeffs.clone().requires_grad_(True)
    sink = torch.full((n * 48,), float("nan"), device="cuda")
    with coeff_grad_sink(sink):
        col = spherical_harmonics(2, dirs, b)
    col.backward(v)
    assert b.grad is None and torch.equal(sink.view(n, 16, 3), a.grad)  # every row written, unused bases zero
    with pytest.raises(ValueError):
        with coeff_grad_sink(torch.zeros(5, device="cuda")):
            spherical_harmonics(2, dirs, b)


def test_flat_adam_device_state_matches_torch_adam_and_vetoes():
    from gsplat.optim import FlatAdam

    g = torch.Generator(device="cuda").manual_seed(11)
    n = 100_004
    flat = torch.randn(n, device="cuda", generator=g)
    ref = flat.clone().requires_grad_(True)
    grad = torch.zeros(n, device="cuda")
    opt_ref = torch.optim.Adam([ref], lr=3e-3, eps=1e-15)
    opt = FlatAdam(flat, grad, lr=3e-3, eps=1e-15).use_device_state()
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    cuts = [0, 4, 4096, 50_000, n]
    applied = 0
    for step in range(6):
        gr = torch.randn(n, device="cuda", generator=g) * (10.0 ** (step - 2))
        grad.copy_(gr)
        veto = step in (2, 4)
        if veto:
            flag.fill_(1)
        else:
            ref.grad = gr.clone() * 0.5
            opt_ref.step()
            applied += 1
        before = flat.clone()
        opt.prepare(flag)
        for j, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):  # (odd slices: the background variant, same arithmetic)
            opt.update_state(a, b, grad_scale=0.5, zero_grad=True, background_ctas=(j % 2) * (1 + step % 3))
        assert (grad == 0).all() and int(flag[0]) == 0
        if veto:
            assert torch.equal(flat, before)
        torch.testing.assert_close(flat, ref.detach(), rtol=2e-6, atol=2e-7)
    st = opt.state.tolist()
    assert st[2] == applied and st[4] == 6 and st[5] == 2 and st[6:8] == [2, 4]
    with pytest.raises(ValueError):
        opt.update_state(2, 10)


def _trainer_scene(n=20000, n_cameras=3):
    import gsplat.synthetic as synthetic

    sc = synthetic.make_scene("c2", device="cuda", n_override=n, n_cameras=n_cameras)
    sc.update(H=128, W=160, fx=80.0, fy=80.0, cx=80.0, cy=64.0)
    cams = []
    for c in sc["cameras"]:
        c.update(fx=80.0, fy=80.0, cx=80.0, cy=64.0, vel0=torch.cat([c["lin_vel"], c["ang_vel"]]))
        c["target"] = c["target"][:128, :160].contiguous()
        cams.append(c)
    return sc, cams


@pytest.mark.parametrize("n", [20000, 20001])
def test_fused_phases_match_the_dropin_phases(n):
    """One image through the two phases of a pipelined step: on the fused raw-parameter kernels (fused_geometry_phase /
    fused_shading_phase: no autograd graph, gradients written into the flat buffer by one kernel) and on the drop-in
    operators under autograd -- same loss, same gradient buffer, same abs-grad statistic.  The odd Gaussian count puts the
    quaternion rows off the 16-byte grid (scalar load / store path of the fused kernels)."""
    from gsplat import dp
    from gsplat.losses import l1_loss

    outs = []
    for fused in (False, True):
        sc, cams = _trainer_scene(n=n)
        m = dp.FlatGaussians(sc, "cuda", n_cameras=3, optimize_velocities=True, sh_layout="split" if fused else "block")
        with torch.no_grad():
            m.cam_vel.add_(0.01 * torch.randn(3, 6, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)))
        c = cams[1]
        st = dict(cam=torch.cat([c["viewmat"].reshape(-1)[:12], c["lin_vel"], c["ang_vel"], c["cam_pos"]]).float().cuda(),
                  cam_index=torch.tensor([1], device="cuda"))
        status = torch.zeros(4, dtype=torch.int32, device="cuda")
        A, B = (dp.fused_geometry_phase, dp.fused_shading_phase) if fused else (dp.geometry_phase, dp.shading_phase)
        geo = A(m, st, sc, 1 << 21, status)
        loss = B(m, geo, sc, c["target"], l1_loss, 3)
        absgrad = geo["absgrad"] if fused else geo["xys"].absgrad
        grads = {k: m.flat_grad[a:b].clone() for k, (a, b) in m.slices.items()}
        if not fused:
            g = grads.pop("sh").view(m.N, m.K, 3)
            grads["sh_dc"], grads["sh_rest"] = g[:, :1].reshape(-1), g[:, 1:].reshape(-1)
        outs.append((float(loss), grads, absgrad, status.tolist(), geo["radii"].clone()))
    (l0, g0, a0, s0, r0), (l1, g1, a1, s1, r1) = outs
    assert s0 == s1 and s0[0] == 0 and s0[1] > 1000 and torch.equal(r0, r1)
    assert abs(l0 - l1) <= 2e-6 * abs(l0)
    assert set(g0) == set(g1)
    # two fp32 evaluation orders of the same chain (exp / normalise inside the kernel vs torch ops before it): elementwise
    # agreement up to a few ill-conditioned Gaussians, the suite's criterion for that (util_scene.grad_close)
    for k in g0:
        assert float(g0[k].abs().max()) > 0, k
        grad_close(g1[k], g0[k], tol=2e-5, name=k, rtol=2e-3, outliers=1e-4)
    grad_close(a1, a0, tol=2e-5, name="absgrad", rtol=2e-3, outliers=1e-4)


@pytest.mark.parametrize("use_graphs,operators", [(False, "dropin"), (True, "dropin"), (False, "fused"), (True, "fused")])
def test_pipelined_trainer_matches_the_synchronising_trainer(use_graphs, operators):
    """Six optimizer steps (two passes over three cameras): gsplat.dp.PipelinedTrainer (capacity-mode lists, device Adam
    state, phase A of image k+1 queued before the SH update of step k; eager and as CUDA graphs; on the drop-in operators
    with the SH block layout + gradient sink, and on the fused raw-parameter kernels) ends at the same parameters as
    ImageShardedTrainer on the synchronising operators."""
    from gsplat.dp import FlatGaussians, ImageShardedTrainer, PipelinedTrainer

    sc, cams = _trainer_scene()
    m0 = FlatGaussians(sc, "cuda", n_cameras=3, optimize_velocities=True)
    t0 = ImageShardedTrainer(m0, sc, lr=1e-3)
    losses0 = []
    for k in range(6):
        losses0.append(float(t0.train_step(cams[k % 3], cams[k % 3]["target"], k % 3)))
    sc, cams = _trainer_scene()
    m1 = FlatGaussians(sc, "cuda", n_cameras=3, optimize_velocities=True, sh_layout="block" if operators == "dropin" else "split")
    t1 = PipelinedTrainer(m1, sc, lr=1e-3, use_graphs=use_graphs, operators=None if operators == "fused" else operators)
    assert t1.operators == operators
    losses1 = []
    t1.prepare(cams[0], 0)
    for k in range(6):
        nxt = (k + 1) % 3 if k + 1 < 6 else None
        loss = t1.train_step(cams[k % 3]["target"], None if nxt is None else cams[nxt], 0 if nxt is None else nxt)
        losses1.append(float(loss))
    t1.finish()
    info = t1.sync_status()
    assert info["vetoed"] == [] and 0 < info["entries_max"] <= info["capacity"]
    if use_graphs:
        assert any(e["gA"] is not None and e["gB"] is not None for e in t1._graphs.values()), "the step was never captured"
    np.testing.assert_allclose(losses1, losses0, rtol=2e-5)
    for k in ("means", "log_scales", "quats", "opacity_logit"):
        assert float((m0.params[k] - m1.params[k]).abs().max()) < 5e-4, k
    assert float((m0.cam_vel - m1.cam_vel).abs().max()) < 5e-4
    sh0 = torch.cat((m0.params["sh_dc"], m0.params["sh_rest"]), 1)
    sh1 = m1.params["sh"] if operators == "dropin" else torch.cat((m1.params["sh_dc"], m1.params["sh_rest"]), 1)
    assert float((sh0 - sh1).abs().max()) < 5e-4  # Adam normalises: tiny gradient differences stay tiny steps


def test_pipelined_trainer_vetoes_an_overflowing_image_and_recovers():
    """A capacity below one image's entry count: that step's update is skipped ON THE DEVICE (parameters untouched,
    gradients cleared), the host learns about it one step late, reports the step and grows the capacity."""
    from gsplat.dp import FlatGaussians, PipelinedTrainer

    sc, cams = _trainer_scene(n_cameras=2)
    m = FlatGaussians(sc, "cuda", n_cameras=2, optimize_velocities=True, sh_layout="block")
    t = PipelinedTrainer(m, sc, lr=1e-3, use_graphs=False, capacity=65536 * 64)
    t.prepare(cams[0], 0)
    t.train_step(cams[0]["target"])
    t.finish()
    need = t.sync_status()["entries_max"]
    assert need > 1024
    t.capacity = need // 4  # far too small
    before = m.flat.detach().clone()
    t.prepare(cams[1], 1)
    t.train_step(cams[1]["target"])
    t.finish()
    info = t.sync_status()
    assert torch.equal(m.flat.detach(), before) and float(m.flat_grad.abs().max()) == 0.0
    assert info["vetoed"] == [1] and info["capacity"] >= need
    t.prepare(cams[1], 1)  # the caller repeats the image
    t.train_step(cams[1]["target"])
    t.finish()
    assert not torch.equal(m.flat.detach(), before) and t.sync_status()["vetoed"] == [1]
