"""GPU: gsplat.densify.Densifier (csrc/densify.cu: flags + prefix sums + per-field gathers on the flat buffers) against
oracle/densify_oracle.py, which tests/test_densify_cpu.py pins to the reference's own refinement_after."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import densify_oracle as DO


def _model(n, layout, seed=0, n_cameras=2):
    import gsplat.synthetic as synthetic
    from gsplat.dp import FlatGaussians

    sc = synthetic.make_scene("c2", device="cuda", n_override=n, n_cameras=n_cameras, seed_offset=seed)
    g = torch.Generator().manual_seed(seed)
    sc["log_scales"] = torch.log(10.0 ** (torch.rand(n, 3, generator=g) * 2.6 - 3.0)).cuda()
    sc["opacity_logit"] = (3.0 * torch.randn(n, 1, generator=g)).cuda()
    return FlatGaussians(sc, "cuda", n_cameras=n_cameras, optimize_velocities=True, sh_layout=layout), sc, g


@pytest.mark.parametrize("layout", ["block", "split"])
@pytest.mark.parametrize("step", [3500, 2500, 3100, 15100])
def test_densifier_matches_the_oracle(layout, step):
    from gsplat.densify import DensifyConfig, Densifier
    from gsplat.optim import FlatAdam

    n, H, W = 20000, 600, 800
    model, sc, g = _model(n, layout, seed=step)
    adam = FlatAdam(model.flat, model.flat_grad, lr=1e-3, eps=1e-15)
    model.flat_grad.copy_(torch.randn(model.flat.numel(), generator=g).cuda())
    adam.step()  # non-trivial moments
    cfg = DensifyConfig()
    dens = Densifier(model, adam, cfg, num_train_data=20)
    stats = {}
    if step < cfg.stop_split_at:
        for _ in range(3):
            radii = ((torch.rand(n, generator=g) * 60).int() * (torch.rand(n, generator=g) < 0.6).int())
            absgrad = torch.rand(n, 2, generator=g) * 2e-3
            dens.accumulate(absgrad.cuda(), radii.cuda(), H, W, step)
            DO.accumulate(stats, absgrad, radii, H, W)
        for k, t in (("grad_norm", dens.grad_norm), ("vis_counts", dens.vis_counts), ("max_2d", dens.max_2d)):
            torch.testing.assert_close(t.cpu(), stats[k], rtol=1e-6 if k == "grad_norm" else 0, atol=1e-9 if k == "grad_norm" else 0)
    else:
        dens.last_size = (H, W)
    names = [k for k in model.params]
    params = {k: model.params[k].detach().cpu().clone() for k in names}
    moments = {k: (adam.exp_avg[model.slices[k][0]:model.slices[k][1]].view(model.params[k].shape).cpu().clone(),
                   adam.exp_avg_sq[model.slices[k][0]:model.slices[k][1]].view(model.params[k].shape).cpu().clone()) for k in names}
    cam_before = model.cam_vel.detach().cpu().clone()
    z_box = {}

    def draw(k):
        z_box["z"] = torch.randn(k, 3, generator=torch.Generator().manual_seed(5))
        return z_box["z"].cuda()

    info = dens.refine(step, normal_samples=draw)
    new_p, new_m, ref = DO.refine(params, moments, stats, cfg, step, 20, (H, W), z=z_box.get("z"))
    assert info is not None and model.N == new_p["means"].shape[0] == info.get("after", model.N)
    if step in (3500, 2500):
        assert info["splits"] == ref["splits"] > 100 and info["dups"] == ref["dups"] > 100
    assert model.flat.numel() == adam.exp_avg.numel() and adam.flat.data_ptr() == model.flat.data_ptr()
    for k in names:
        a0, a1 = model.slices[k]
        # exp / log of the scales and the child placement run through different libm's: last-bit agreement
        tol = dict(rtol=2e-6, atol=2e-6) if k in ("means", "log_scales") else dict(rtol=0, atol=0)
        torch.testing.assert_close(model.params[k].detach().cpu(), new_p[k], **tol, msg=lambda s: f"{k}: {s}")
        torch.testing.assert_close(adam.exp_avg[a0:a1].view(model.params[k].shape).cpu(), new_m[k][0], rtol=0, atol=0)
        torch.testing.assert_close(adam.exp_avg_sq[a0:a1].view(model.params[k].shape).cpu(), new_m[k][1], rtol=0, atol=0)
    torch.testing.assert_close(model.cam_vel.detach().cpu(), cam_before, rtol=0, atol=0)  # camera rows carried over
    assert dens.grad_norm is None and float(model.flat_grad.abs().max()) == 0.0


@pytest.mark.parametrize("layout", ["block", "split"])
def test_training_continues_across_a_refinement(layout):
    """PipelinedTrainer (CUDA graphs) -> statistics -> refine -> on_resize -> more steps on the new buffers."""
    import gsplat.synthetic as synthetic
    from gsplat.densify import DensifyConfig, Densifier
    from gsplat.dp import FlatGaussians, PipelinedTrainer

    sc = synthetic.make_scene("c2", device="cuda", n_override=20000, n_cameras=2)
    sc.update(H=128, W=160, fx=80.0, fy=80.0, cx=80.0, cy=64.0)
    cams = sc["cameras"]
    for c in cams:
        c["target"] = c["target"][:128, :160].contiguous()
    model = FlatGaussians(sc, "cuda", n_cameras=2, optimize_velocities=True, sh_layout=layout)
    tr = PipelinedTrainer(model, sc, lr=1e-3, use_graphs=True)
    assert tr.operators == ("fused" if layout == "split" else "dropin")   # split: the raw-parameter kernels
    cfg = DensifyConfig(warmup_length=0, refine_every=5, densify_grad_thresh=1e-7, reset_alpha_every=30)
    dens = Densifier(model, tr.adam, cfg, num_train_data=2)
    tr.after_backward = lambda absgrad, radii: dens.accumulate(absgrad, radii, 128, 160, step=tr.steps)
    tr.prepare(cams[0], 0)
    n0 = model.N
    for k in range(8):
        tr.train_step(cams[k % 2]["target"], cams[(k + 1) % 2], (k + 1) % 2)
    tr.finish()
    # every image reached the statistics: a Gaussian seen by one of the two alternating cameras counts its 4 images
    # (+ the 1 the first image gives everybody)
    assert dens.vis_counts is not None and float(dens.vis_counts.max()) >= 4.0
    info = dens.refine(step=8)
    assert info is not None and info["densified"] and info["after"] != n0 and model.N == info["after"]
    tr.on_resize()
    tr.prepare(cams[0], 0)
    losses = []
    for k in range(6):
        losses.append(float(tr.train_step(cams[k % 2]["target"], cams[(k + 1) % 2], (k + 1) % 2)))
    tr.finish()
    assert all(np.isfinite(losses)) and tr.sync_status()["vetoed"] == []
    assert any(e["gB"] is not None for e in tr._graphs.values()), "no graph was captured at the new size"
