"""CPU: the oracle (oracle/) against golden vectors produced by the reference itself.

Fixtures come from tests/golden/make_golden.py, which imports the reference's
gsplat/_torch_impl.py in the build container.  Tolerances are the reference's own
(gsplat/tests/test_project_gaussians.py:129-136: atol=rtol=1e-5 forward,
:319-325: atol 5e-4 backward), scaled by the magnitude of pixel-unit quantities.
"""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle import torch_oracle as TO

CASES = ["proj_seed42.npz", "proj_seed42_static.npz", "proj_c1.npz", "proj_c2s.npz"]


def _close(a, b, atol=1e-5, rtol=1e-5, name=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert (err <= 0).all(), f"{name}: max excess {err.max():.3e} (max abs diff {np.abs(a - b).max():.3e})"


@pytest.mark.parametrize("case", CASES)
def test_c_oracle_projection_forward_matches_reference_torch(golden, case):
    g = golden(case)
    o = O.project_forward(g["means"], g["scales"], float(g["glob_scale"]), g["quats"], g["lin_vel"], g["ang_vel"],
                          float(g["rs_time"]), float(g["exposure"]), g["viewmat"], float(g["fx"]), float(g["fy"]),
                          float(g["cx"]), float(g["cy"]), int(g["H"]), int(g["W"]), 16, 0.01)
    m_ref = g["mask"]
    m = o["num_tiles_hit"] > 0
    # Reference-internal discrepancy (documented in DESIGN.md "quirks"): the CUDA bbox is (int)(c + r + 1)
    # (helpers.cuh:18,20) while the torch twin is int(c + r) + 1 (_torch_impl.py:378); they differ when
    # -1 < c + r < 0, i.e. for Gaussians lying entirely above / left of the image, which only the torch
    # path keeps.  The oracle restates the CUDA kernel, so its mask is a subset with exactly that residue.
    assert not (m & ~m_ref).any()
    extra = m_ref & ~m
    if extra.any():
        hi = (g["xys"][extra] + g["radii"][extra, None] + 1.0) / 16.0  # upper bound of c + r in tile units
        assert (hi.min(axis=1) < 1.0 / 16.0 + 1e-3).all() and ((g["xys"][extra] / 16.0).min(axis=1) < 0).all()
    assert (o["num_tiles_hit"][m] == g["num_tiles_hit"][m]).all()
    assert (o["radii"][m] == g["radii"][m]).all()
    _close(o["cov3d"][m], g["cov3d"][m], name="cov3d")
    _close(o["xys"][m], g["xys"][m], atol=2e-4, name="xys")  # pixels, |xy| up to ~1e3
    _close(o["depths"][m], g["depths"][m], name="depths")
    _close(o["conics"][m], g["conics"][m], atol=1e-5, rtol=1e-4, name="conics")
    _close(o["compensation"][m], g["compensation"][m], name="comp")
    # pix_vels is computed but never asserted by the reference test; CUDA uses 1/z, torch 1/(z+1e-6)
    _close(o["pix_vels"][m], g["pix_vels"][m], atol=1e-3, rtol=1e-4, name="pix_vels")


@pytest.mark.parametrize("case", CASES)
def test_torch_oracle_matches_reference_torch_forward_and_grads(golden, case):
    g = golden(case)
    t = lambda k, dt=torch.float32: torch.from_numpy(np.asarray(g[k])).to(dt)
    cfg = dict(glob_scale=float(g["glob_scale"]), rs_time=float(g["rs_time"]), exposure=float(g["exposure"]),
               fx=float(g["fx"]), fy=float(g["fy"]), cx=float(g["cx"]), cy=float(g["cy"]), H=int(g["H"]),
               W=int(g["W"]), block_width=16)
    for dt, tol in ((torch.float32, 1.0), (torch.float64, 1.0)):
        inputs = {k: t(k, dt) for k in ("means", "scales", "quats", "lin_vel", "ang_vel", "viewmat")}
        cts = {k: t(k, dt) for k in ("v_xys", "v_depths", "v_pix_vels", "v_conics", "v_compensation")}
        grads, out = TO.project_vjp(inputs, cts, **cfg)
        m = g["mask"]
        assert (out["mask"].numpy() == m).all()
        assert (out["num_tiles_hit"].numpy() == g["num_tiles_hit"]).all()
        assert (out["radii"].numpy() == g["radii"]).all()
        _close(out["xys"].detach()[m], g["xys"][m], atol=2e-4, name="xys")
        _close(out["conics"].detach()[m], g["conics"][m], atol=1e-5, rtol=1e-4, name="conics")
        _close(out["pix_vels"].detach()[m], g["pix_vels"][m], atol=1e-3, rtol=1e-4, name="pix_vels")
        for k in ("means", "scales", "quats", "lin_vel", "ang_vel", "viewmat"):
            ref = g["g_" + k]
            scale = max(1.0, float(np.abs(ref).max()))
            _close(grads["v_" + k].numpy() / scale, ref / scale, atol=5e-4, rtol=1e-3, name="g_" + k)


def test_c_oracle_projection_backward_matches_autograd_of_cuda_path(golden):
    """C restatement of backward.cu:371-572 vs fp64 autograd, on Gaussians inside 1.3x FOV
    (outside, the CUDA backward deliberately ignores the clamp: SURVEY 7 quirk c)."""
    g = golden("proj_seed42.npz")
    cfg = dict(glob_scale=float(g["glob_scale"]), rs_time=float(g["rs_time"]), exposure=float(g["exposure"]),
               fx=float(g["fx"]), fy=float(g["fy"]), cx=float(g["cx"]), cy=float(g["cy"]), H=int(g["H"]),
               W=int(g["W"]), block_width=16)
    t64 = lambda k: torch.from_numpy(np.asarray(g[k])).double()
    inputs = {k: t64(k) for k in ("means", "scales", "quats", "lin_vel", "ang_vel", "viewmat")}
    cts = {k: t64(k) for k in ("v_xys", "v_depths", "v_pix_vels", "v_conics", "v_compensation")}
    grads, out = TO.project_vjp(inputs, cts, **cfg)
    fwd = O.project_forward(g["means"], g["scales"], cfg["glob_scale"], g["quats"], g["lin_vel"], g["ang_vel"],
                            cfg["rs_time"], cfg["exposure"], g["viewmat"], cfg["fx"], cfg["fy"], cfg["cx"], cfg["cy"],
                            cfg["H"], cfg["W"], 16)
    b = O.project_backward(g["means"], g["scales"], cfg["glob_scale"], g["quats"], g["lin_vel"], g["ang_vel"],
                           cfg["rs_time"], cfg["exposure"], g["viewmat"], cfg["fx"], cfg["fy"], fwd["cov3d"],
                           fwd["radii"], fwd["conics"], fwd["compensation"], g["v_xys"], g["v_depths"],
                           g["v_pix_vels"], g["v_conics"], g["v_compensation"])
    vm = g["viewmat"]
    pv = g["means"] @ vm[:3, :3].T + vm[:3, 3]
    inside = (np.abs(pv[:, 0] / pv[:, 2]) < 1.3 * 0.5 * cfg["W"] / cfg["fx"]) & \
             (np.abs(pv[:, 1] / pv[:, 2]) < 1.3 * 0.5 * cfg["H"] / cfg["fy"]) & (fwd["radii"] > 0)
    assert inside.sum() > 20
    for k, ref in (("v_mean3d", grads["v_means"]), ("v_scale", grads["v_scales"]), ("v_quat", grads["v_quats"])):
        r = ref.numpy()[inside]
        scale = max(1.0, np.abs(r).max())
        _close(b[k][inside] / scale, r / scale, atol=5e-4, rtol=1e-3, name=k)
    # culled Gaussians get exactly zero
    assert (b["v_mean3d"][fwd["radii"] <= 0] == 0).all()


@pytest.mark.parametrize("method", ["poly", "fast"])
def test_c_oracle_sh_matches_reference(golden, method):
    g = golden("sh.npz")
    for deg in range(5):
        col = O.sh_forward(method, deg, g["dirs"], g["coeffs"])
        k = (deg + 1) ** 2
        ref = g[f"{method}_{deg}"]
        _close(col, ref, atol=2e-5, rtol=1e-5, name=f"sh {method} {deg}")
        # restricted coefficient tensor gives the same colours
        col2 = O.sh_forward(method, deg, g["dirs"], np.ascontiguousarray(g["coeffs"][:, :k]))
        assert np.array_equal(col, col2)
        # backward = basis outer v_colors; check through linearity against the forward
        v = np.random.default_rng(deg).standard_normal((64, 3)).astype(np.float32)
        vc = O.sh_backward(method, 4, deg, g["dirs"], v)
        assert (vc[:, k:] == 0).all()
        lhs = (vc.astype(np.float64) * g["coeffs"]).sum()
        rhs = (col.astype(np.float64) * v).sum()
        assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(rhs))


def test_c_oracle_map_and_bins_match_reference_loops(golden):
    g = golden("map_bins.npz")
    H, W, bw = int(g["H"]), int(g["W"]), int(g["bw"])
    b = O.bin_and_sort(g["xys"], g["depths"], g["radii"], g["num_tiles_hit"], H, W, bw)
    assert np.array_equal(b["cum_tiles_hit"], g["cum_tiles_hit"])
    assert np.array_equal(b["isect_ids"], g["isect_ids"])
    assert np.array_equal(b["gaussian_ids"], g["gaussian_ids"])
    assert np.array_equal(b["isect_ids_sorted"], g["isect_ids_sorted"])
    assert np.array_equal(b["gaussian_ids_sorted"], g["gaussian_ids_sorted"])
    assert np.array_equal(b["tile_bins"], g["tile_bins"])


def test_c_oracle_cov2d_bounds_matches_reference(golden):
    g = golden("cov2d_bounds.npz")
    conics, radii = O.cov2d_bounds(g["cov2d"])
    v = g["valid"]
    _close(conics[v], g["conics"][v], atol=5e-4, name="conics")
    _close(radii[v, 0], g["radii"][v], atol=5e-4, name="radii")


def test_blend_oracle_properties():
    """The blur/RS blend has no reference test or torch twin: check the restatement's invariants."""
    rng = np.random.default_rng(0)
    n, H, W, bw, S = 400, 48, 64, 16, 3
    xys = rng.uniform([0, 0], [W, H], (n, 2)).astype(np.float32)
    depths = rng.uniform(0.5, 5, n).astype(np.float32)
    radii = rng.integers(2, 14, n).astype(np.int32)
    sig = rng.uniform(1.5, 5, n)
    conics = np.stack([1 / sig**2, rng.uniform(-0.02, 0.02, n), 1 / sig**2], -1).astype(np.float32)
    vel = rng.normal(0, 30, (n, 2)).astype(np.float32)
    tb = O.tile_bounds(H, W, bw)
    nth = np.zeros(n, np.int32)
    for i in range(n):  # consistent tile counts from the int radii
        tc, tr = xys[i] / bw, radii[i] / bw
        lo = np.clip((tc - tr).astype(int), 0, tb)
        hi = np.clip((tc + tr + 1).astype(int), 0, tb)
        nth[i] = (hi[0] - lo[0]) * (hi[1] - lo[1])
    b = O.bin_and_sort(xys, depths, radii, nth, H, W, bw)
    cols = rng.uniform(0, 1, (n, 3)).astype(np.float32)
    opac = rng.uniform(0.05, 0.9, (n, 1)).astype(np.float32)
    bg = np.array([0.2, 0.4, 0.6], np.float32)
    args = (b["gaussian_ids_sorted"], b["tile_bins"], xys, vel)
    img, Ts, fi = O.rasterize_forward(H, W, bw, S, *args, 0.02, 0.1, conics, cols, opac, bg)
    assert np.isfinite(img).all() and (Ts > 0).all() and (Ts <= 1).all()
    # linear in (colors, background)
    img2, Ts2, fi2 = O.rasterize_forward(H, W, bw, S, *args, 0.02, 0.1, conics, 2 * cols, opac, 2 * bg)
    assert np.array_equal(fi, fi2) and np.array_equal(Ts, Ts2)
    np.testing.assert_allclose(img2, 2 * img, rtol=1e-5, atol=1e-6)
    # zero velocity: every blur sample identical, equals the single-sample render
    z = np.zeros_like(vel)
    i1, T1, f1 = O.rasterize_forward(H, W, bw, 1, b["gaussian_ids_sorted"], b["tile_bins"], xys, z, 0.0, 0.0, conics, cols, opac, bg)
    i3, T3, f3 = O.rasterize_forward(H, W, bw, 3, b["gaussian_ids_sorted"], b["tile_bins"], xys, z, 0.02, 0.1, conics, cols, opac, bg)
    assert (T3 == T1).all() and (f3 == f1).all()
    np.testing.assert_allclose(i3, i1, rtol=1e-5, atol=1e-6)
    # backward vs central differences of the forward on colour / opacity / xy of a contributing Gaussian
    v_out = rng.normal(0, 1, (H, W, 3)).astype(np.float32)
    v_alpha = rng.normal(0, 1, (H, W)).astype(np.float32)
    gr = O.rasterize_backward(H, W, bw, S, *args, 0.02, 0.1, conics, cols, opac, bg, Ts, fi, v_out, v_alpha)

    def loss(cols_, opac_, xys_, vel_):
        im, T, _ = O.rasterize_forward(H, W, bw, S, b["gaussian_ids_sorted"], b["tile_bins"], xys_, vel_, 0.02, 0.1,
                                       conics, cols_, opac_, bg)
        return float((im.astype(np.float64) * v_out).sum() + ((1 - T.mean(-1).astype(np.float64)) * v_alpha).sum())

    # colour gradient is exact (the render is linear in colour)
    gid = int(np.argmax(np.abs(gr["v_colors"]).sum(1)))
    eps = 1e-2
    cp, cm = cols.copy(), cols.copy()
    cp[gid, 1] += eps
    cm[gid, 1] -= eps
    fd = (loss(cp, opac, xys, vel) - loss(cm, opac, xys, vel)) / (2 * eps)
    assert abs(fd - gr["v_colors"][gid, 1]) <= 2e-2 * max(1.0, abs(fd))
    # abs-grad dominates the signed grad
    assert (gr["v_xy_abs"] + 1e-6 >= np.abs(gr["v_xy"])).all()
