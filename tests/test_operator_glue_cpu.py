"""CPU: the autograd glue of gsplat.rasterize (argument plumbing, the three branches RGB / empty / N-channel, unused-output
cotangents, the absgrad side channel) with the C-ABI calls replaced by shape-faithful fakes.  The kernels themselves are
tested on the GPU (tests/test_gpu_parity.py); this guards the Python around them where no GPU is available."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dgs-deblur_b200"))


@pytest.fixture
def fake_C(monkeypatch):
    import gsplat.cuda as _C
    import gsplat.rasterize as R
    import gsplat.utils as U
    calls = []

    def pack_records(xys, pix_vels, conics, colors, opacity):
        calls.append("pack")
        return torch.zeros(xys.shape[0] * 64, dtype=torch.uint8)

    def bin_cull(packed, depths, radii, nth, H, W, bw, S, rs, ex):
        calls.append(("bin_cull", H, W, bw, S, rs, ex))
        tiles = ((W + bw - 1) // bw) * ((H + bw - 1) // bw)
        m = int(nth.sum())
        return m, torch.zeros(m, dtype=torch.int32), torch.zeros(tiles, 2, dtype=torch.int32)

    def blend_forward_packed(H, W, bw, S, ids, bins, packed, rs, ex, bg, want_alpha=False):
        calls.append(("fwd", S, want_alpha))
        out = (torch.full((H, W, 3), 0.25), torch.full((H, W, S), 0.5), torch.zeros(H, W, S, dtype=torch.int32))
        return out + (torch.full((H, W), 0.5),) if want_alpha else out

    def blend_backward_packed(n, H, W, bw, S, ids, bins, packed, rs, ex, bg, Ts, fi, v_out, v_alpha):
        calls.append(("bwd", tuple(v_out.shape), None if v_alpha is None else tuple(v_alpha.shape)))
        return (torch.ones(n, 2), torch.full((n, 2), 2.0), torch.ones(n, 2), torch.ones(n, 3), torch.ones(n, 3), torch.ones(n, 1))

    def bin_tiles(m, xys, depths, radii, nth, tb, bw):
        calls.append(("bin_tiles", m, tb))
        return torch.zeros(m, dtype=torch.int32), torch.zeros(tb[0] * tb[1], 2, dtype=torch.int32)

    def nd_rasterize_forward(tb, block, img_size, S, ids, bins, xys, pv, rs, ex, conics, colors, opac, bg):
        calls.append(("nd_fwd", block, img_size, colors.shape[-1]))
        W, H = img_size[0], img_size[1]
        return torch.zeros(H, W, colors.shape[-1]), torch.full((H, W), 0.25), torch.zeros(H, W, dtype=torch.int32)

    def nd_rasterize_backward(H, W, bw, S, ids, bins, xys, pv, rs, ex, conics, colors, opac, bg, Ts, fi, v_out, v_alpha):
        calls.append(("nd_bwd", tuple(v_out.shape), tuple(v_alpha.shape)))
        n, c = xys.shape[0], colors.shape[-1]
        return torch.ones(n, 2), torch.ones(n, 2), torch.zeros(n, 2), torch.ones(n, 3), torch.ones(n, c), torch.ones(n, 1)

    for name, fn in list(locals().items()):
        if callable(fn) and hasattr(_C, name):
            monkeypatch.setattr(_C, name, fn)
    monkeypatch.setattr(R, "compute_cumulative_intersects", lambda nth: (int(nth.sum()), torch.cumsum(nth, 0)))
    return calls


def _inputs(n=6, channels=3, hit=True):
    xys = torch.rand(n, 2, requires_grad=True)
    colors = torch.rand(n, channels, requires_grad=True)
    opacity = torch.rand(n, 1, requires_grad=True)
    nth = torch.full((n,), 2 if hit else 0, dtype=torch.int32)
    return xys, torch.rand(n), torch.rand(n, 2), torch.ones(n, dtype=torch.int32), torch.rand(n, 3), nth, colors, opacity


def test_rgb_branch_plumbing_and_unused_alpha(fake_C):
    from gsplat.rasterize import rasterize_gaussians
    xys, depths, pv, radii, conics, nth, colors, opacity = _inputs()
    bg = torch.rand(3, requires_grad=True)
    img, alpha = rasterize_gaussians(xys, depths, pv, radii, conics, nth, colors, opacity, 20, 36, 16, bg, True, 0.02, 0.016, 5)
    assert img.shape == (20, 36, 3) and alpha.shape == (20, 36) and float(alpha[0, 0]) == 0.5
    assert fake_C[0] == "pack" and fake_C[1] == ("bin_cull", 20, 36, 16, 5, 0.02, 0.016) and fake_C[2] == ("fwd", 5, True)
    img.sum().backward()  # alpha unused: its cotangent is None and no (H, W) zero image is built
    assert fake_C[3] == ("bwd", (20, 36, 3), None)
    assert torch.equal(xys.grad, torch.ones(6, 2)) and torch.equal(xys.absgrad, torch.full((6, 2), 2.0))
    assert colors.grad.shape == (6, 3) and opacity.grad.shape == (6, 1)
    # d img / d background = sum over pixels of mean_s(final T) = 0.5 * H * W per channel
    torch.testing.assert_close(bg.grad, torch.full((3,), 0.5 * 20 * 36))
    # only alpha used: the image cotangent is materialised as zeros, alpha's is passed through
    xys2, *rest = _inputs()
    img2, alpha2 = rasterize_gaussians(xys2, *rest, 20, 36, 16, None, True, 0, 0, 1)
    alpha2.sum().backward()
    assert fake_C[-1] == ("bwd", (20, 36, 3), (20, 36))
    # return_alpha=False returns the image alone
    assert rasterize_gaussians(*_inputs(), 20, 36, 16).shape == (20, 36, 3)


def test_empty_render_matches_the_reference_behaviour(fake_C):
    from gsplat.rasterize import rasterize_gaussians
    xys, depths, pv, radii, conics, nth, colors, opacity = _inputs(hit=False)
    bg = torch.tensor([0.1, 0.2, 0.3])
    img, alpha = rasterize_gaussians(xys, depths, pv, radii, conics, nth, colors, opacity, 8, 12, 16, bg, True, 0, 0.01, 3)
    assert torch.allclose(img, bg.expand(8, 12, 3)) and bool((alpha == 1).all())
    (img.sum() + alpha.sum()).backward()
    assert all(float(t.grad.abs().sum()) == 0 for t in (xys, colors, opacity)) and float(xys.absgrad.abs().sum()) == 0
    assert not any(isinstance(c, tuple) and c[0] in ("fwd", "bwd") for c in fake_C)


def test_n_channel_branch_and_argument_checks(fake_C):
    from gsplat.rasterize import rasterize_gaussians
    xys, depths, pv, radii, conics, nth, colors, opacity = _inputs(channels=5)
    img, alpha = rasterize_gaussians(xys, depths, pv, radii, conics, nth, colors, opacity, 20, 36, 8, return_alpha=True)
    assert img.shape == (20, 36, 5) and float(alpha[0, 0]) == 0.75
    assert ("bin_tiles", 12, (5, 3, 1)) in fake_C and ("nd_fwd", (8, 8, 1), (36, 20, 1), 5) in fake_C
    img.sum().backward()
    assert fake_C[-1] == ("nd_bwd", (20, 36, 5), (20, 36)) and colors.grad.shape == (6, 5)
    with pytest.raises(RuntimeError, match="unsupported blur size"):
        rasterize_gaussians(*_inputs(), 20, 36, 16, blur_samples=11)
    with pytest.raises(AssertionError):
        rasterize_gaussians(*_inputs(), 20, 36, 17)
    u8 = _inputs()
    img_u8 = rasterize_gaussians(*u8[:6], (u8[6].detach() * 255).to(torch.uint8), u8[7], 20, 36, 16)
    assert img_u8.dtype == torch.float32


def test_spherical_harmonics_glue(monkeypatch):
    import gsplat.cuda as _C
    from gsplat.sh import deg_from_sh, num_sh_bases, spherical_harmonics
    seen = []

    def fwd(method, n, degree, degrees_to_use, viewdirs, coeffs):
        seen.append(("fwd", method, n, degree, degrees_to_use))
        return coeffs[:, 0, :] * 2.0

    def bwd(method, n, degree, degrees_to_use, viewdirs, v_colors):
        seen.append(("bwd", method, n, degree, degrees_to_use))
        out = torch.zeros(n, num_sh_bases(degree), 3)
        out[:, 0, :] = 2.0 * v_colors
        return out

    monkeypatch.setattr(_C, "compute_sh_forward", fwd)
    monkeypatch.setattr(_C, "compute_sh_backward", bwd)
    coeffs = torch.rand(7, 16, 3, requires_grad=True)
    dirs = torch.rand(7, 3, requires_grad=True)
    col = spherical_harmonics(2, dirs, coeffs, "poly")
    col.sum().backward()
    assert seen == [("fwd", "poly", 7, 3, 2), ("bwd", "poly", 7, 3, 2)]
    assert dirs.grad is None and torch.equal(coeffs.grad[:, 0, :], torch.full((7, 3), 2.0)) and float(coeffs.grad[:, 1:].abs().sum()) == 0
    assert [num_sh_bases(d) for d in (0, 1, 2, 3, 4, 7)] == [1, 4, 9, 16, 25, 25] and deg_from_sh(25) == 4
    with pytest.raises(AssertionError):
        spherical_harmonics(3, dirs, torch.rand(7, 4, 3))  # fewer bases than the requested degree needs
    with pytest.raises(AssertionError):
        spherical_harmonics(1, dirs, coeffs, "exact")
    with pytest.raises(AssertionError, match="Invalid number of SH bases"):
        spherical_harmonics(0, dirs, torch.rand(7, 5, 3))


def test_project_gaussians_glue(monkeypatch):
    """Velocity handling (None / (3,) / (1,3), constant vs requires_grad -> exact mode), the view-matrix gradient slot,
    output order and the non-differentiable integer outputs, with the two C-ABI calls faked."""
    import gsplat.cuda as _C
    from gsplat import _lib
    from gsplat.project_gaussians import project_gaussians
    seen = {}

    def fwd(n, means, scales, glob_scale, quats, lin, ang, rs, ex, viewmat, fx, fy, cx, cy, H, W, bw, clip, _vel_tensors=None,
            _quat_flag=None):
        seen["fwd"] = dict(n=n, vel=[t.clone() for t in _vel_tensors], rs=rs, ex=ex, flag=_quat_flag is not None, clip=clip)
        z = means.sum() * 0  # keeps the outputs attached to nothing: the Function supplies the graph
        return (torch.zeros(n, 6), torch.ones(n, 2) + z.detach(), torch.ones(n), torch.zeros(n, 2), torch.ones(n, dtype=torch.int32),
                torch.ones(n, 3), torch.ones(n), torch.ones(n, dtype=torch.int32))

    def bwd(n, means, scales, glob_scale, quats, lin, ang, rs, ex, viewmat, fx, fy, cx, cy, H, W, cov3d, radii, conics, comp, v_xy,
            v_depth, v_pix, v_conic, v_comp, _vel_tensors=None, _exact=False, _want_vel=False, _want_viewmat=False, _want_cov=True):
        seen["bwd"] = dict(exact=_exact, want_vel=_want_vel, want_vm=_want_viewmat, want_cov=_want_cov)
        out = (None, None, torch.ones(n, 3), torch.full((n, 3), 2.0), torch.full((n, 4), 3.0))
        if _want_vel:
            out = out + (torch.tensor([1.0, 2.0, 3.0]), torch.tensor([4.0, 5.0, 6.0]))
        if _want_viewmat:
            out = out + (torch.full((3, 4), 7.0),)
        return out

    monkeypatch.setattr(_C, "project_gaussians_forward", fwd)
    monkeypatch.setattr(_C, "project_gaussians_backward", bwd)
    monkeypatch.setattr(_lib, "new_quat_flag", lambda dev: torch.zeros(1, dtype=torch.int32))
    n = 5
    means = torch.rand(n, 3, requires_grad=True)
    scales = torch.rand(n, 3, requires_grad=True)
    quats = torch.nn.functional.normalize(torch.randn(n, 4), dim=-1).requires_grad_(True)
    viewmat = torch.eye(4)[:3].clone()
    args = lambda lin, ang, vm: (means, scales, 1.0, quats, lin, ang, 0.02 if lin is not None else 0, 0.01, vm, 50.0, 50.0,
                                 32.0, 24.0, 48, 64, 16)
    # constant velocities given as (1,3): reference CUDA-path gradients, no velocity / view-matrix outputs requested
    out = project_gaussians(*args(torch.tensor([[0.1, 0.2, 0.3]]), torch.tensor([[0.4, 0.5, 0.6]]), viewmat))
    assert len(out) == 8 and out[0].shape == (n, 2) and out[3].dtype == torch.int32 and out[7].shape == (n, 6)
    assert not out[3].requires_grad and not out[6].requires_grad
    assert torch.equal(seen["fwd"]["vel"][0], torch.tensor([0.1, 0.2, 0.3])) and seen["fwd"]["clip"] == 0.01
    (out[0].sum() + out[4].sum()).backward()
    assert seen["bwd"] == dict(exact=False, want_vel=False, want_vm=False, want_cov=False)
    assert torch.equal(means.grad, torch.ones(n, 3)) and torch.equal(quats.grad, torch.full((n, 4), 3.0))
    # velocities that require grad ((3,) and (1,3) shapes) + a view matrix that requires grad: exact mode, shaped grads
    lin = torch.tensor([0.1, 0.2, 0.3], requires_grad=True)
    ang = torch.tensor([[0.4, 0.5, 0.6]], requires_grad=True)
    vm = torch.eye(4).requires_grad_(True)
    out = project_gaussians(*args(lin, ang, vm))
    out[0].sum().backward()
    assert seen["bwd"] == dict(exact=True, want_vel=True, want_vm=True, want_cov=False)
    assert torch.equal(lin.grad, torch.tensor([1.0, 2.0, 3.0])) and torch.equal(ang.grad, torch.tensor([[4.0, 5.0, 6.0]]))
    assert vm.grad.shape == (4, 4) and float(vm.grad[:3].sum()) == 84.0 and float(vm.grad[3].abs().sum()) == 0
    # no velocities: zeros are passed down, and a rolling-shutter time without a velocity is rejected like the reference
    project_gaussians(means, scales, 1.0, quats, None, None, 0, 0.0, viewmat, 50.0, 50.0, 32.0, 24.0, 48, 64, 16)
    assert all(float(v.abs().sum()) == 0 for v in seen["fwd"]["vel"])
    with pytest.raises(AssertionError):
        project_gaussians(means, scales, 1.0, quats, None, None, 0.02, 0.0, viewmat, 50.0, 50.0, 32.0, 24.0, 48, 64, 16)
    with pytest.raises(AssertionError, match="quats must be normalized"):  # CPU tensors are checked eagerly
        project_gaussians(means, scales, 1.0, quats.detach() * 2, None, None, 0, 0.0, viewmat, 50.0, 50.0, 32.0, 24.0, 48, 64, 16)
    with pytest.raises(AssertionError, match="block_width"):
        project_gaussians(means, scales, 1.0, quats, None, None, 0, 0.0, viewmat, 50.0, 50.0, 32.0, 24.0, 48, 64, 1)


def test_fakes_have_the_real_wrappers_arity():
    """The fakes above stand in for gsplat.cuda wrappers: same number of positional parameters, so a plumbing mistake
    in the glue cannot hide behind a more permissive fake."""
    import inspect
    import gsplat.cuda as _C
    expected = dict(pack_records=5, bin_cull=10, blend_forward_packed=11, blend_backward_packed=15, bin_tiles=7,
                    nd_rasterize_forward=14, nd_rasterize_backward=18, compute_sh_forward=6, compute_sh_backward=6)
    for name, n in expected.items():
        params = [p for p in inspect.signature(getattr(_C, name)).parameters.values()
                  if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
        assert len(params) == n, (name, [p.name for p in params])
    fwd = inspect.signature(_C.project_gaussians_forward).parameters
    bwd = inspect.signature(_C.project_gaussians_backward).parameters
    assert list(fwd)[:18][-1] == "clip_thresh" and "_vel_tensors" in fwd and "_quat_flag" in fwd
    assert all(k in bwd for k in ("_vel_tensors", "_exact", "_want_vel", "_want_viewmat", "_want_cov"))


def test_static_second_pass_reuses_lists_only_when_they_provably_cover_it(fake_C):
    """gsplat.rasterize list reuse (the caller's depth pass, splatfacto.py:881-897): same per-Gaussian tensors, static
    second call, colour pass without rolling shutter and with an odd sample count (or no exposure)."""
    import gsplat.rasterize as R
    from gsplat.rasterize import rasterize_gaussians

    def bins():
        return [c for c in fake_C if isinstance(c, tuple) and c[0] == "bin_cull"]

    def run(S, rs, ex, same=True, touch=False):
        R._last_lists.clear()
        del fake_C[:]
        xys, depths, pv, radii, conics, nth, colors, opacity = _inputs()
        rasterize_gaussians(xys, depths, pv, radii, conics, nth, colors, opacity, 20, 36, 16, None, True, rs, ex, S)
        if touch:
            with torch.no_grad():
                conics.add_(0.0)  # an in-place write bumps the version counter: the lists may be stale
        if not same:
            conics = conics.clone()
        depth_cols = depths[:, None].repeat(1, 3)
        rasterize_gaussians(xys, depths, pv, radii, conics, nth, depth_cols, opacity, 20, 36, 16, background=torch.zeros(3))
        return len(bins())

    assert run(5, 0.0, 0.016) == 1          # odd sample count, no rolling shutter: reused
    assert run(1, 0.0, 0.0) == 1            # static colour pass: reused
    assert run(5, 0.0, 0.0) == 1            # no exposure: every sample sits at offset 0
    assert run(4, 0.0, 0.016) == 2          # even count: no sample at offset 0
    assert run(5, 0.02, 0.016) == 2         # rolling shutter: per-row offsets
    assert run(5, 0.0, 0.016, same=False) == 2   # another tensor
    assert run(5, 0.0, 0.016, touch=True) == 2   # same tensor, written since
    # a second call that is not static never reuses
    R._last_lists.clear()
    del fake_C[:]
    a = _inputs()
    rasterize_gaussians(*a, 20, 36, 16, None, True, 0.0, 0.016, 5)
    rasterize_gaussians(*a, 20, 36, 16, None, True, 0.0, 0.016, 5)
    assert len(bins()) == 2


def test_loss_target_options_are_checked_before_any_kernel():
    from gsplat.losses import _target_options

    pred = torch.zeros(16, 20, 3)
    rgb, rgba = torch.zeros(16, 20, 3, dtype=torch.uint8), torch.zeros(16, 20, 4, dtype=torch.uint8)
    ch, bg, level, mask = _target_options("l1_loss", pred, rgb, torch.ones(3), 12.0, torch.ones(16, 20, 1, dtype=torch.bool))
    assert ch == 3 and bg is None and abs(level - 12.0 / 255.0) < 1e-12 and mask.shape == (16, 20) and mask.dtype == torch.float32
    ch, bg, level, mask = _target_options("l1_loss", pred, rgba, torch.tensor([[0.1, 0.2, 0.3]]), 0.0, None)
    assert ch == 4 and bg.shape == (3,) and level == 0.0 and mask is None
    with pytest.raises(ValueError):
        _target_options("l1_loss", pred, rgba, None, 0.0, None)            # RGBA needs a background
    with pytest.raises(ValueError):
        _target_options("l1_loss", pred, torch.zeros(16, 21, 3, dtype=torch.uint8), None, 0.0, None)
    with pytest.raises(ValueError):
        _target_options("l1_loss", pred, rgb, None, 0.0, torch.ones(16, 21))  # mask of another size
    with pytest.raises(ValueError):
        _target_options("l1_loss", pred, rgb, None, -1.0, None)
    with pytest.raises(ValueError):
        _target_options("l1_loss", pred, rgba, torch.ones(4), 0.0, None)
