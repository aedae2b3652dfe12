"""CPU: oracle/densify_oracle.py (the restatement the GPU test holds gsplat.densify to) against the reference ITSELF --
the real SplatfactoModel.after_train / refinement_after (/root/reference/nerfstudio/nerfstudio/models/splatfacto.py:
408-531) run here on the CPU with real torch.optim.Adam optimizers, on the same parameters, statistics, step and random
draw.  Build container only (needs /root/reference); the stubs are those of tests/test_splatfacto_caller_cpu.py."""
import os
import sys
import types

import pytest
import torch

from test_splatfacto_caller_cpu import REF_NS, ROOT, _StubFinder  # noqa: F401

pytestmark = pytest.mark.skipif(not os.path.isdir(REF_NS), reason="needs /root/reference (build container only)")

from oracle import densify_oracle as DO  # noqa: E402


@pytest.fixture(scope="module")
def sf():
    sys.path.insert(0, os.path.join(ROOT, "3dgs-deblur_b200"))
    sys.path.insert(0, REF_NS)
    finder = _StubFinder()
    sys.meta_path.insert(0, finder)
    try:
        import nerfstudio.models.splatfacto as m
        from nerfstudio.data.scene_box import SceneBox
        yield m, SceneBox
    finally:
        sys.meta_path.remove(finder)
        sys.path.remove(REF_NS)


NAMES = {"means": "means", "scales": "log_scales", "quats": "quats", "opacities": "opacity_logit", "features_dc": "features_dc",
         "features_rest": "features_rest"}


def _setup(sf, n, seed):
    m, SceneBox = sf
    g = torch.Generator().manual_seed(seed)
    pts = (torch.rand(n, 3, generator=g) - 0.5) * 2.0
    cfg = m.SplatfactoModelConfig(sh_degree=3)
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        model = m.SplatfactoModel(cfg, scene_box=SceneBox(aabb=torch.tensor([[-1.0, -1, -1], [1.0, 1, 1]])), num_train_data=20,
                                  seed_points=(pts, torch.rand(n, 3, generator=g) * 255))
    finally:
        torch.Tensor.cuda = real_cuda
    with torch.no_grad():  # a spread of sizes / opacities so every branch (split, dup, three kinds of cull) is taken
        model.gauss_params["scales"].copy_(torch.log(10.0 ** (torch.rand(n, 3, generator=g) * 2.6 - 3.0)))
        model.gauss_params["opacities"].copy_(3.0 * torch.randn(n, 1, generator=g))
        model.gauss_params["features_rest"].copy_(0.1 * torch.randn(n, 15, 3, generator=g))
    groups = model.get_gaussian_param_groups()
    opts = types.SimpleNamespace(optimizers={k: torch.optim.Adam(v, lr=1e-3, eps=1e-15) for k, v in groups.items()})
    for k, v in groups.items():  # one real step so every optimizer carries non-trivial moments
        v[0].grad = torch.randn(v[0].shape, generator=g)
        opts.optimizers[k].step()
    return model, opts, g


def _stats_images(model, g, n, images, H, W):
    """`images` after_train calls through the reference, mirrored on the oracle."""
    stats = {}
    for _ in range(images):
        radii = (torch.rand(n, generator=g) * 60).int() * (torch.rand(n, generator=g) < 0.6).int()
        absgrad = torch.rand(n, 2, generator=g) * 2e-3
        model.radii = radii
        model.xys = types.SimpleNamespace(absgrad=absgrad)
        model.last_size = (H, W)
        model.after_train(model.step)
        DO.accumulate(stats, absgrad, radii, H, W)
    torch.testing.assert_close(stats["grad_norm"], model.xys_grad_norm, rtol=0, atol=0)
    torch.testing.assert_close(stats["vis_counts"], model.vis_counts, rtol=0, atol=0)
    torch.testing.assert_close(stats["max_2d"], model.max_2Dsize, rtol=0, atol=0)
    return stats


# (step -> which branch): densify with every cull criterion; densify before the "too big" culls switch on; the step of an
# opacity reset; cull-only after stop_split_at; a step the schedule skips
@pytest.mark.parametrize("step", [3500, 2500, 3100, 15100, 3000 + 50])
def test_oracle_matches_the_reference_refinement(sf, step):
    n, H, W = 3000, 600, 800
    model, opts, g = _setup(sf, n, seed=step)
    model.step = step
    cfg = model.config
    stats = _stats_images(model, g, n, 3, H, W) if step < cfg.stop_split_at else {}
    params = {NAMES[k]: v.detach().clone() for k, v in model.gauss_params.items()}
    moments = {NAMES[k]: (o.state[o.param_groups[0]["params"][0]]["exp_avg"].clone(), o.state[o.param_groups[0]["params"][0]]["exp_avg_sq"].clone())
               for k, o in opts.optimizers.items()}
    from gsplat.densify import DensifyConfig
    dcfg = DensifyConfig()
    for f in dcfg.__dataclass_fields__:  # the port's defaults ARE the reference's
        assert getattr(dcfg, f) == getattr(cfg, f), f
    torch.manual_seed(77)
    model.refinement_after(opts, step)
    torch.manual_seed(77)
    new_p, new_m, info = DO.refine(params, moments, stats, dcfg, step, model.num_train_data, (H, W))
    assert model.num_points == new_p["means"].shape[0]
    if step in (3500, 2500):
        assert info["splits"] > 50 and info["dups"] > 50 and info["after"] != n
    if step == 15100:
        assert info is not None and not info["densified"] and info["after"] < n
    for k, v in model.gauss_params.items():
        # (children's means go through quat -> rotation matrix: last-bit differences between two spellings of it)
        torch.testing.assert_close(new_p[NAMES[k]], v.detach(), rtol=0, atol=2e-6 if k == "means" else 0, msg=lambda s: f"{k}: {s}")
        o = opts.optimizers[k]
        st = o.state[o.param_groups[0]["params"][0]]
        torch.testing.assert_close(new_m[NAMES[k]][0], st["exp_avg"], rtol=0, atol=0)
        torch.testing.assert_close(new_m[NAMES[k]][1], st["exp_avg_sq"], rtol=0, atol=0)
    assert model.xys_grad_norm is None and model.max_2Dsize is None
