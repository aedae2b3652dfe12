"""CPU: the UNMODIFIED caller calls this package.

The reference's SplatfactoModel (/root/reference/nerfstudio/nerfstudio/models/splatfacto.py:28-31 imports
`gsplat.project_gaussians / rasterize / sh / _torch_impl`; :682-899 `get_outputs`) is imported here with `gsplat`
resolving to 3dgs-deblur_b200/gsplat -- nothing of the caller is edited or replayed -- and run through get_outputs,
a loss and loss.backward(), in training mode (motion blur + rolling shutter + velocity optimisation, "antialiased"
opacities) and in eval mode (its second, depth-coloured rasterize_gaussians call).

What runs under the operators: this container has no GPU and /root/reference cannot travel to the GPU box, so the C-ABI
layer (`gsplat.cuda`, the 1:1 wrappers of libb200splat) is swapped for the CPU ORACLE (oracle/splat_oracle.c, the checker
the GPU parity tests hold the kernels to).  Everything between the caller and the C ABI is the product's own code: the
three operators' argument handling, autograd Functions, gradient routing to velocities / view matrix, the `xys.absgrad`
side channel, the (rgb, alpha) return convention.  The rendered image is checked against the oracle chain driven directly
from the model's parameters, and the gradients against float64 finite differences of that chain's loss.

Packages the reference's import chain needs and this image lacks (viser, torchmetrics, pytorch_msssim, nerfacc) are
stubbed; none of them is on the render path."""
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_NS = "/root/reference/nerfstudio"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_NS), reason="needs /root/reference (build container only)")

from oracle import oracle as O  # noqa: E402


class _StubMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub


class _Stub(metaclass=_StubMeta):
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Stub()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub()


class _StubModule(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    NAMES = ("viser", "torchmetrics", "pytorch_msssim", "nerfacc")

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.NAMES:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


@pytest.fixture(scope="module")
def splatfacto():
    sys.path.insert(0, os.path.join(ROOT, "3dgs-deblur_b200"))
    sys.path.insert(0, REF_NS)
    finder = _StubFinder()
    sys.meta_path.insert(0, finder)
    try:
        import gsplat
        assert os.path.realpath(os.path.dirname(gsplat.__file__)).startswith(os.path.realpath(ROOT)), "gsplat must be THIS package"
        import nerfstudio.models.splatfacto as sf
        from nerfstudio.cameras.cameras import Cameras
        from nerfstudio.data.scene_box import SceneBox
        yield sf, Cameras, SceneBox
    finally:
        sys.meta_path.remove(finder)
        sys.path.remove(REF_NS)


# ---- gsplat.cuda on the oracle -------------------------------------------------------------------------------------

def _n(t):
    return t.detach().cpu().numpy()


def _unpack(packed):
    """The stand-in "packed records": one (N, 11) float tensor xy | pix_vel | conic | colour | opacity."""
    a = packed.detach().numpy()
    return types.SimpleNamespace(xys=np.ascontiguousarray(a[:, 0:2]), pix_vels=np.ascontiguousarray(a[:, 2:4]),
                                 conics=np.ascontiguousarray(a[:, 4:7]), colors=np.ascontiguousarray(a[:, 7:10]),
                                 opac=np.ascontiguousarray(a[:, 10:11]))


@pytest.fixture
def oracle_C(monkeypatch):
    """Every `gsplat.cuda` entry the three operators reach, computed by the CPU oracle (same signatures)."""
    import gsplat.cuda as _C
    import gsplat._lib as L
    calls = []
    monkeypatch.setattr(L, "SYNC_CHECKS", True)  # the quaternion assert runs in Python (no device flag on the CPU)

    def project_gaussians_forward(n, means3d, scales, glob_scale, quats, lin, ang, rs, ex, viewmat, fx, fy, cx, cy, H, W, bw,
                                  clip, _vel_tensors=None, _quat_flag=None):
        lin, ang = _vel_tensors
        calls.append(("project_fwd", n, H, W, bw, float(rs), float(ex)))
        o = O.project_forward(_n(means3d), _n(scales), glob_scale, _n(quats), _n(lin), _n(ang), rs, ex, _n(viewmat), fx, fy, cx, cy, H, W, bw)
        return tuple(torch.from_numpy(o[k]) for k in ("cov3d", "xys", "depths", "pix_vels", "radii", "conics", "compensation", "num_tiles_hit"))

    def project_gaussians_backward(n, means3d, scales, glob_scale, quats, lin, ang, rs, ex, viewmat, fx, fy, cx, cy, H, W, cov3d, radii,
                                   conics, comp, v_xy, v_depth, v_pix_vel, v_conic, v_comp, _vel_tensors=None, _exact=False,
                                   _want_vel=False, _want_viewmat=False, _want_cov=True):
        """Exact mode = float64 autograd through oracle/torch_oracle.py (the reference's torch-path semantics)."""
        from oracle import torch_oracle as TO

        lin, ang = _vel_tensors
        calls.append(("project_bwd", bool(_exact), bool(_want_vel), bool(_want_viewmat)))
        t64 = lambda a: a.detach().double()
        vm4 = torch.cat([t64(viewmat).reshape(-1)[:12].view(3, 4), torch.tensor([[0, 0, 0, 1.0]], dtype=torch.float64)], 0)
        inputs = dict(means=t64(means3d), scales=t64(scales), quats=t64(quats), lin_vel=t64(lin), ang_vel=t64(ang), viewmat=vm4)
        ct = dict(v_xys=t64(v_xy), v_depths=t64(v_depth), v_pix_vels=t64(v_pix_vel), v_conics=t64(v_conic), v_compensation=t64(v_comp))
        with torch.enable_grad():  # (autograd is off inside a Function's backward)
            g, _ = TO.project_vjp(inputs, ct, glob_scale=glob_scale, rs_time=rs, exposure=ex, fx=fx, fy=fy, cx=cx, cy=cy, H=H, W=W, block_width=16)
        # (rows of Gaussians the projection culled: their cotangents are zero, but autograd through the masked-out branch
        # of the restatement yields 0 * inf there; the kernels write plain zeros)
        g = {k: torch.nan_to_num(v, nan=0.0, posinf=0.0, neginf=0.0) for k, v in g.items()}
        out = (None, None, g["v_means"].float(), g["v_scales"].float(), g["v_quats"].float())
        if _want_vel:
            out += (g["v_lin_vel"].float().reshape(3), g["v_ang_vel"].float().reshape(3))
        if _want_viewmat:
            out += (g["v_viewmat"].float()[:3],)
        return out

    def compute_sh_forward(method, n, degree, deg_use, viewdirs, coeffs):
        calls.append(("sh_fwd", method, degree, deg_use))
        return torch.from_numpy(O.sh_forward(method, deg_use, _n(viewdirs), _n(coeffs)))

    def compute_sh_backward(method, n, degree, deg_use, viewdirs, v_colors, *, out=None):
        calls.append(("sh_bwd", method, degree, deg_use))
        return torch.from_numpy(O.sh_backward(method, degree, deg_use, _n(viewdirs), _n(v_colors)))

    def pack_records(xys, pix_vels, conics, colors, opacity):
        return torch.cat([xys.detach(), pix_vels.detach(), conics.detach(), colors.detach(), opacity.detach().reshape(-1, 1)], 1).float().contiguous()

    def bin_cull(packed, depths, radii, nth, H, W, bw, S, rs, ex):
        calls.append(("bin", H, W, bw, S))
        b = O.bin_and_sort(_unpack(packed).xys, _n(depths), _n(radii), _n(nth), H, W, bw)
        return int(b["num_intersects"]), torch.from_numpy(b["gaussian_ids_sorted"]), torch.from_numpy(b["tile_bins"])

    def blend_forward_packed(H, W, bw, S, ids, bins, packed, rs, ex, bg, want_alpha=False, *, status=None):
        calls.append(("blend_fwd", S, float(rs), float(ex)))
        packed = _unpack(packed)
        img, Ts, fi = O.rasterize_forward(H, W, bw, S, _n(ids), _n(bins), packed.xys, packed.pix_vels, rs, ex, packed.conics,
                                          packed.colors, packed.opac, _n(bg))
        out = (torch.from_numpy(img), torch.from_numpy(Ts), torch.from_numpy(fi))
        return out + (torch.from_numpy(1 - Ts.mean(-1)),) if want_alpha else out

    def blend_backward_packed(n, H, W, bw, S, ids, bins, packed, rs, ex, bg, Ts, fi, v_out, v_alpha):
        calls.append(("blend_bwd", S))
        packed = _unpack(packed)
        va = np.zeros((H, W), np.float32) if v_alpha is None else _n(v_alpha)
        g = O.rasterize_backward(H, W, bw, S, _n(ids), _n(bins), packed.xys, packed.pix_vels, rs, ex, packed.conics, packed.colors,
                                 packed.opac, _n(bg), _n(Ts), _n(fi), _n(v_out), va)
        return tuple(torch.from_numpy(g[k]) for k in ("v_xy", "v_xy_abs", "v_pix_vels", "v_conic", "v_colors", "v_opacity"))

    for name, fn in list(locals().items()):
        if callable(fn) and hasattr(_C, name):
            monkeypatch.setattr(_C, name, fn)
    return calls


def _make_model(sf, SceneBox, n, training, velocity_opt, seed=3):
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed)
    pts = (torch.rand(n, 3, generator=g) - 0.5) * 2.0 + torch.tensor([0.0, 0.0, -3.0])  # in front of an OpenGL camera at the origin
    cols = torch.rand(n, 3, generator=g) * 255
    cfg = sf.SplatfactoModelConfig(rasterize_mode="antialiased", blur_samples=5, background_color="white", num_downscales=0,
                                   sh_degree=3, sh_degree_interval=0 if False else 1)
    cfg.camera_velocity_optimizer.enabled = velocity_opt
    box = SceneBox(aabb=torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]]))
    real_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self  # populate_modules moves the seed colours to "cuda" (splatfacto.py:223)
    try:
        model = sf.SplatfactoModel(cfg, scene_box=box, num_train_data=2, seed_points=(pts, cols))
    finally:
        torch.Tensor.cuda = real_cuda
    model.step = 10  # all SH degrees on (splatfacto.py:844)
    model.train(training)
    with torch.no_grad():  # something to render: visible sizes, mixed opacities, non-trivial higher SH bands
        model.gauss_params["scales"].copy_(torch.log(torch.full((n, 3), 0.05) * (0.5 + torch.rand(n, 3, generator=g))))
        model.gauss_params["opacities"].copy_(torch.randn(n, 1, generator=g))
        model.gauss_params["features_rest"].copy_(0.1 * torch.randn(n, 15, 3, generator=g))
    return model


def _camera(Cameras, H=48, W=64, with_motion=True):
    c2w = torch.eye(4)[:3].unsqueeze(0).clone()
    meta = dict(exposure_time=1 / 60, rolling_shutter_time=1 / 50, cam_idx=0) if with_motion else None  # (cam_idx: camera_optimizers.py:248)
    vel = torch.tensor([[0.3, -0.2, 0.1, 0.05, 0.4, -0.3]]) if with_motion else None
    return Cameras(camera_to_worlds=c2w, fx=float(W) / 2, fy=float(W) / 2, cx=W / 2.0, cy=H / 2.0, width=W, height=H, velocities=vel,
                   metadata=meta)


def _oracle_chain(model, cam, vel6, S, rs, ex, H, W):
    """The render block of splatfacto.py:734-880 restated on the oracle, from the model's parameters (numpy, float32)."""
    p = {k: _n(v) for k, v in model.gauss_params.items()}
    R_edit = np.diag([1.0, -1.0, -1.0]).astype(np.float32)
    c2w = _n(cam.camera_to_worlds[0])
    R = c2w[:3, :3] @ R_edit
    viewmat = np.concatenate([R.T, -R.T @ c2w[:3, 3:4]], 1).astype(np.float32)
    lin, ang = (R_edit @ vel6[:3]).astype(np.float32), (R_edit @ vel6[3:]).astype(np.float32)
    q = p["quats"] / np.linalg.norm(p["quats"], axis=-1, keepdims=True)
    proj = O.project_forward(p["means"], np.exp(p["scales"]), 1.0, q.astype(np.float32), lin, ang, rs, ex, viewmat, W / 2.0, W / 2.0, W / 2.0,
                             H / 2.0, H, W, 16)
    coeffs = np.concatenate([p["features_dc"][:, None, :], p["features_rest"]], 1)
    rgbs = np.maximum(O.sh_forward("fast", 3, p["means"] - c2w[:3, 3][None], coeffs) + 0.5, 0).astype(np.float32)
    opac = (1 / (1 + np.exp(-p["opacities"][:, 0])) * proj["compensation"]).astype(np.float32)[:, None]
    b = O.bin_and_sort(proj["xys"], proj["depths"], proj["radii"], proj["num_tiles_hit"], H, W, 16)
    img, Ts, fi = O.rasterize_forward(H, W, 16, S, b["gaussian_ids_sorted"], b["tile_bins"], proj["xys"], proj["pix_vels"], rs, ex,
                                      proj["conics"], rgbs, opac, np.ones(3, np.float32))  # "white" (a black background would
    #   put exact zeros under the caller's x ** (1 / gamma), whose derivative there is infinite -- in the reference too)
    return np.minimum(img, 1.0) ** (1 / 2.2), 1 - Ts.mean(-1), proj


def test_unmodified_splatfacto_trains_through_this_package(splatfacto, oracle_C):
    sf, Cameras, SceneBox = splatfacto
    assert sf.project_gaussians.__module__ == "gsplat.project_gaussians" and sf.rasterize_gaussians.__module__ == "gsplat.rasterize"
    H, W, n = 48, 64, 400
    model = _make_model(sf, SceneBox, n, training=True, velocity_opt=True)
    cam = _camera(Cameras, H, W)
    out = model.get_outputs(cam)
    rgb, acc = out["rgb"], out["accumulation"]
    assert rgb.shape == (H, W, 3) and acc.shape == (H, W, 1) and out["depth"] is None
    # one projection (velocities carry gradients -> exact mode), one SH call at degree 3, one 5-sample blur + RS blend
    assert ("project_fwd", n, H, W, 16, 1 / 50, 1 / 60) in oracle_C and ("sh_fwd", "fast", 3, 3) in oracle_C
    assert ("blend_fwd", 5, 1 / 50, 1 / 60) in oracle_C
    ref_rgb, ref_alpha, proj = _oracle_chain(model, cam, _n(cam.velocities[0]), 5, 1 / 50, 1 / 60, H, W)
    assert int((proj["num_tiles_hit"] > 0).sum()) > 50
    np.testing.assert_allclose(_n(rgb), ref_rgb, atol=1e-6)
    np.testing.assert_allclose(_n(acc[..., 0]), ref_alpha, atol=1e-6)
    target = torch.rand(H, W, 3, generator=torch.Generator().manual_seed(1))
    loss = (rgb - target).abs().mean() + 0.1 * acc.mean()
    loss.backward()
    assert ("project_bwd", True, True, False) in oracle_C and ("blend_bwd", 5) in oracle_C and ("sh_bwd", "fast", 3, 3) in oracle_C
    for k, v in model.gauss_params.items():
        assert v.grad is not None and torch.isfinite(v.grad).all() and float(v.grad.abs().sum()) > 0, k
    vel_params = [p for p in model.camera_velocity_optimizer.parameters()]
    assert vel_params and all(p.grad is not None and float(p.grad.abs().sum()) > 0 for p in vel_params), "no gradient reached the velocity optimizer"
    # the densification side channel the caller reads next (splatfacto.py:416-417)
    assert model.xys.absgrad.shape == (n, 2) and float(model.xys.absgrad.sum()) > 0 and (model.xys.absgrad >= 0).all()
    assert model.xys.grad is None or model.xys.grad.shape == (n, 2)
    # a gradient check that goes through the whole unmodified caller: d loss / d features_dc by central differences
    idx = int(np.argmax(_n(model.gauss_params["features_dc"].grad).sum(-1).__abs__()))
    with torch.no_grad():
        base = model.gauss_params["features_dc"][idx, 0].item()
        vals = []
        for eps in (1e-2, -1e-2):
            model.gauss_params["features_dc"][idx, 0] = base + eps
            o = model.get_outputs(cam)
            vals.append(float((o["rgb"] - target).abs().mean() + 0.1 * o["accumulation"].mean()))
        model.gauss_params["features_dc"][idx, 0] = base
    fd = (vals[0] - vals[1]) / 2e-2
    assert abs(fd - float(model.gauss_params["features_dc"].grad[idx, 0])) < 0.05 * abs(fd) + 1e-6, (fd, float(model.gauss_params["features_dc"].grad[idx, 0]))


def test_unmodified_splatfacto_eval_pass_renders_depth_through_this_package(splatfacto, oracle_C):
    """Eval mode: static camera (no velocity data, optimizer off), no blur -> S = 1, and the caller's second
    rasterize_gaussians call with depth-valued colours (splatfacto.py:881-897)."""
    sf, Cameras, SceneBox = splatfacto
    H, W, n = 32, 48, 300
    model = _make_model(sf, SceneBox, n, training=False, velocity_opt=False)
    cam = _camera(Cameras, H, W, with_motion=False)
    with torch.no_grad():
        out = model.get_outputs(cam)
    assert out["rgb"].shape == (H, W, 3) and out["depth"].shape == (H, W, 1) and out["accumulation"].shape == (H, W, 1)
    assert [c for c in oracle_C if c[0] == "blend_fwd"] == [("blend_fwd", 1, 0.0, 0.0)] * 2
    ref_rgb, ref_alpha, proj = _oracle_chain(model, cam, np.zeros(6, np.float32), 1, 0.0, 0.0, H, W)
    np.testing.assert_allclose(_n(out["rgb"]), ref_rgb, atol=1e-6)
    covered = ref_alpha > 0.5
    d = _n(out["depth"][..., 0])
    assert covered.any() and np.all(d[covered] > 1.5) and np.all(d[covered] < 4.5)  # the cloud sits 2..4 units in front


@pytest.mark.parametrize("rolling_shutter", [False, True])
def test_eval_depth_pass_reuses_the_colour_pass_lists(splatfacto, oracle_C, monkeypatch, rolling_shutter):
    """Eval with motion blur: the caller's second (static, depth-coloured) rasterize_gaussians call (splatfacto.py:881-897)
    bins nothing when the colour pass had no rolling shutter and an odd sample count -- its lists contain the static
    lists (gsplat/rasterize.py) -- and bins again when it had.  Same depth image either way."""
    sf, Cameras, SceneBox = splatfacto
    H, W, n = 32, 48, 300
    model = _make_model(sf, SceneBox, n, training=False, velocity_opt=False)
    meta = dict(exposure_time=1 / 60, cam_idx=0)
    if rolling_shutter:
        meta["rolling_shutter_time"] = 1 / 50
    cam = Cameras(camera_to_worlds=torch.eye(4)[:3].unsqueeze(0).clone(), fx=W / 2.0, fy=W / 2.0, cx=W / 2.0, cy=H / 2.0, width=W,
                  height=H, velocities=torch.tensor([[0.3, -0.2, 0.1, 0.05, 0.4, -0.3]]), metadata=meta)
    with torch.no_grad():
        out = model.get_outputs(cam)
    blends = [c for c in oracle_C if c[0] == "blend_fwd"]
    assert blends[0][1] == 5 and blends[1] == ("blend_fwd", 1, 0.0, 0.0)
    assert len([c for c in oracle_C if c[0] == "bin"]) == (2 if rolling_shutter else 1)
    del oracle_C[:]
    monkeypatch.setenv("B200SPLAT_NO_LIST_REUSE", "1")
    import gsplat.rasterize as R
    R._last_lists.clear()
    with torch.no_grad():
        out2 = model.get_outputs(cam)
    assert len([c for c in oracle_C if c[0] == "bin"]) == 2
    assert torch.equal(out["depth"], out2["depth"]) and torch.equal(out["rgb"], out2["rgb"])
    d = out["depth"][..., 0][out["accumulation"][..., 0] > 0.5]  # (blurred coverage can exceed the static one at the rim: depth 0 there)
    assert d.numel() > 0 and float(d.min()) > 1.0 and float(d.max()) < 4.5
