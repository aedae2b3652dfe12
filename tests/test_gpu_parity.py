"""GPU parity: libb200splat (through the gsplat operator surface / C ABI) vs the CPU oracle.

Tolerances (fp32 path; the reference's own tests use atol=rtol=1e-5 forward, 5e-4 backward,
gsplat/tests/test_project_gaussians.py:129-136,:319-325):
  * ordering outputs (isect ids, sorted ids, tile bins, cumulative counts): bit exact;
  * thresholded integers (radii, num_tiles_hit, final_idx): exact on reference-sized / static cases; with motion
    the blur-inflated radius is a continuous value truncated to int and the last-contributor index depends on
    `alpha < 1/255` / `T <= 1e-4`, so 1-ulp differences (FMA contraction on the GPU vs plain fp32 in the oracle;
    the reference CUDA build itself uses fast-math) flip ~1e-4..1e-3 of the entries -- bounded per test;
  * projection floats: 1e-5 relative (+2e-4 px absolute for pixel coordinates);
  * image / final_Ts: atol 2e-5 / 1e-6 (+rel 1e-5) for all but <= 1e-3 of the pixel-samples, where one flipped
    contribution (<= 1/255 of the remaining transmittance) is tolerated up to 1e-2; overall PSNR vs oracle > 80 dB;
  * gradients: elementwise 5e-3 relative + 1e-3 of the tensor's max magnitude (fp32 atomics in arbitrary order vs the
    oracle's fp64 sums), <= 2e-4 outliers, cosine similarity > 1 - 1e-5.
"""
import os
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import gsplat
    import gsplat.cuda as _C
    from gsplat import project_gaussians, rasterize_gaussians, spherical_harmonics

from oracle import oracle as O
from oracle import torch_oracle as TO
from util_scene import close, cu, frac_mismatch, grad_close, oracle_colors, oracle_project, oracle_render, scene_np


def gpu_project(d, lin=None, ang=None, viewmat=None, means=None, scales=None, quats=None):
    return project_gaussians(
        cu(d["means"]) if means is None else means, cu(d["scales"]) if scales is None else scales, 1.0,
        cu(d["quats"]) if quats is None else quats, cu(d["lin_vel"]) if lin is None else lin,
        cu(d["ang_vel"]) if ang is None else ang, d["rs"], d["exposure"],
        cu(d["viewmat"]) if viewmat is None else viewmat, d["fx"], d["fy"], d["cx"], d["cy"], d["H"], d["W"], d["bw"])


# motion on: radius = ceil(3 sigma) + |pix_vel| * 0.5 * (exposure + rs) is truncated to int, so a 1-ulp difference in
# |pix_vel| (FMA contraction vs the oracle's plain fp32) flips (int)radius for ~1e-4 of the Gaussians
PROJ_CASES = [("c1", None, False, 0.0), ("c2", 20000, True, 1e-3), ("c2", None, True, 1e-3), ("c4", 50000, True, 3e-3)]


@pytest.mark.parametrize("name,n,motion,int_tol", PROJ_CASES)
def test_projection_forward_vs_oracle(name, n, motion, int_tol):
    d = scene_np(name, n=n, motion=motion)
    o = oracle_project(d)
    xys, depths, pix_vels, radii, conics, comp, nth, cov3d = gpu_project(d)
    assert frac_mismatch(radii.cpu().numpy(), o["radii"]) <= int_tol
    assert frac_mismatch(nth.cpu().numpy(), o["num_tiles_hit"]) <= int_tol
    m = (nth.cpu().numpy() > 0) & (o["num_tiles_hit"] > 0)
    assert m.sum() > 100
    close(cov3d.cpu().numpy()[m], o["cov3d"][m], 1e-7, 1e-5, "cov3d")
    # Gaussians within a few clip distances of the camera plane have |xy| ~ 1e5 px and a relative depth error of
    # ~1e-7 / 0.01: allow 1e-3 of the rows to exceed the elementwise bound (still within 1e-3 relative)
    close(xys.cpu().numpy()[m], o["xys"][m], 2e-4, 1e-5, "xys", outliers=1e-3)
    close(xys.cpu().numpy()[m], o["xys"][m], 2e-4, 1e-3, "xys (hard bound)")
    close(depths.cpu().numpy()[m], o["depths"][m], 1e-6, 1e-5, "depths")
    close(conics.cpu().numpy()[m], o["conics"][m], 1e-6, 1e-4, "conics", outliers=1e-4)
    close(comp.cpu().numpy()[m], o["compensation"][m], 1e-5, 1e-5, "compensation", outliers=1e-4)
    close(pix_vels.cpu().numpy()[m], o["pix_vels"][m], 1e-3, 1e-4, "pix_vels", outliers=1e-4)
    close(pix_vels.cpu().numpy()[m], o["pix_vels"][m], 1e-3, 1e-3, "pix_vels (hard bound)")
    # culled Gaussians: every output the reference leaves at its zeros init is zero here too
    c = o["num_tiles_hit"] == 0
    assert (xys.cpu().numpy()[c] == o["xys"][c]).all() and (depths.cpu().numpy()[c] == 0).all()
    close(conics.cpu().numpy()[c], o["conics"][c], 1e-6, 1e-4, "conics of culled (written before the bbox cull)")


def test_projection_forward_reference_test_case(golden):
    """The reference's own test inputs (seed 42, N=100, 512^2, rs=0.1, exposure=0.2) at its tolerances."""
    for case in ("proj_seed42.npz", "proj_seed42_static.npz"):
        g = golden(case)
        out = project_gaussians(cu(g["means"]), cu(g["scales"]), float(g["glob_scale"]), cu(g["quats"]),
                                cu(g["lin_vel"]), cu(g["ang_vel"]), float(g["rs_time"]), float(g["exposure"]),
                                cu(g["viewmat"]), float(g["fx"]), float(g["fy"]), float(g["cx"]), float(g["cy"]),
                                int(g["H"]), int(g["W"]), 16, 0.01)
        xys, depths, pix_vels, radii, conics, comp, nth, cov3d = [t.cpu().numpy() for t in out]
        m = g["mask"]
        assert ((nth > 0) == m).all()
        assert (radii[m] == g["radii"][m]).all() and (nth[m] == g["num_tiles_hit"][m]).all()
        close(cov3d[m], g["cov3d"][m], 1e-5, 1e-5, "cov3d")
        close(xys[m], g["xys"][m], 2e-4, 1e-5, "xys")
        close(depths[m], g["depths"][m], 1e-5, 1e-5, "depths")
        close(conics[m], g["conics"][m], 1e-5, 1e-4, "conics")
        close(comp[m], g["compensation"][m], 1e-5, 1e-5, "comp")
        close(pix_vels[m], g["pix_vels"][m], 1e-3, 1e-4, "pix_vels")


def _cotangents(n, mask, seed):
    g = np.random.default_rng(seed)
    mk = mask.astype(np.float32)
    return dict(v_xys=(g.standard_normal((n, 2)) * mk[:, None]).astype(np.float32),
                v_depths=(g.standard_normal(n) * mk).astype(np.float32),
                v_pix_vels=(g.standard_normal((n, 2)) * 0.01 * mk[:, None]).astype(np.float32),
                v_conics=(g.standard_normal((n, 3)) * mk[:, None]).astype(np.float32),
                v_compensation=(g.standard_normal(n) * mk).astype(np.float32))


@pytest.mark.parametrize("name,n", [("c2", 20000), ("c4", 30000)])
def test_projection_backward_cuda_path_vs_c_oracle(name, n):
    """Velocities without grad -> the reference's CUDA backward (backward.cu:371-572) + its approximate
    viewmat gradient (project_gaussians.py:272-307)."""
    d = scene_np(name, n=n)
    o = oracle_project(d)
    ct = _cotangents(d["N"], o["num_tiles_hit"] > 0, 3)
    means = cu(d["means"]).requires_grad_(True)
    scales = cu(d["scales"]).requires_grad_(True)
    quats = cu(d["quats"]).requires_grad_(True)
    vm = cu(d["viewmat"]).requires_grad_(True)
    xys, depths, pix_vels, radii, conics, comp, nth, cov3d = gpu_project(d, viewmat=vm, means=means, scales=scales, quats=quats)
    loss = ((xys * cu(ct["v_xys"])).sum() + (depths * cu(ct["v_depths"])).sum() + (pix_vels * cu(ct["v_pix_vels"])).sum()
            + (conics * cu(ct["v_conics"])).sum() + (comp * cu(ct["v_compensation"])).sum())
    loss.backward()
    b = O.project_backward(d["means"], d["scales"], 1.0, d["quats"], d["lin_vel"], d["ang_vel"], d["rs"], d["exposure"],
                           d["viewmat"], d["fx"], d["fy"], o["cov3d"], o["radii"], o["conics"], o["compensation"],
                           ct["v_xys"], ct["v_depths"], ct["v_pix_vels"], ct["v_conics"], ct["v_compensation"])
    same = (radii.cpu().numpy() > 0) == (o["radii"] > 0)
    grad_close(means.grad.cpu().numpy()[same], b["v_mean3d"][same], 2e-3, "v_mean3d")
    grad_close(scales.grad.cpu().numpy()[same], b["v_scale"][same], 2e-3, "v_scale")
    grad_close(quats.grad.cpu().numpy()[same], b["v_quat"][same], 2e-3, "v_quat")
    # approximate viewmat gradient of the CUDA path: v_cam = v_mean R^T ; t: sum ; R[j,l] = sum v_cam[j] mean[l]
    R = d["viewmat"][:3, :3].astype(np.float64)
    v_cam = b["v_mean3d"].astype(np.float64) @ R.T
    ref_vm = np.concatenate([v_cam.T @ d["means"].astype(np.float64), v_cam.sum(0)[:, None]], 1)
    grad_close(vm.grad, ref_vm, 5e-3, "v_viewmat (approx)")


@pytest.mark.parametrize("name,n", [("c2", 20000), ("c4", 30000)])
def test_projection_backward_exact_path_vs_fp64_autograd(name, n):
    """Velocities with grad -> gradients of the reference's torch path incl. velocity / exact viewmat grads."""
    d = scene_np(name, n=n)
    o = oracle_project(d)
    ct = _cotangents(d["N"], o["num_tiles_hit"] > 0, 4)
    leaves = dict(means=cu(d["means"]), scales=cu(d["scales"]), quats=cu(d["quats"]), lin=cu(d["lin_vel"]),
                  ang=cu(d["ang_vel"]), vm=cu(d["viewmat"]))
    for v in leaves.values():
        v.requires_grad_(True)
    xys, depths, pix_vels, radii, conics, comp, nth, cov3d = gpu_project(
        d, lin=leaves["lin"], ang=leaves["ang"], viewmat=leaves["vm"], means=leaves["means"], scales=leaves["scales"],
        quats=leaves["quats"])
    # only send gradient where both sides rasterise the Gaussian (the torch bbox keeps a few more: see DESIGN quirks)
    vm4 = np.concatenate([d["viewmat"], np.array([[0, 0, 0, 1.0]], np.float32)], 0)
    t64 = lambda a: torch.from_numpy(np.asarray(a)).double()
    inputs = dict(means=t64(d["means"]), scales=t64(d["scales"]), quats=t64(d["quats"]), lin_vel=t64(d["lin_vel"]),
                  ang_vel=t64(d["ang_vel"]), viewmat=t64(vm4))
    cfg = dict(glob_scale=1.0, rs_time=d["rs"], exposure=d["exposure"], fx=d["fx"], fy=d["fy"], cx=d["cx"], cy=d["cy"],
               H=d["H"], W=d["W"], block_width=16)
    with torch.no_grad():
        mask64 = TO.project(inputs["means"], inputs["scales"], 1.0, inputs["quats"], inputs["lin_vel"],
                            inputs["ang_vel"], d["rs"], d["exposure"], inputs["viewmat"], d["fx"], d["fy"], d["cx"],
                            d["cy"], d["H"], d["W"], 16)["mask"].numpy()
    both = mask64 & (radii.cpu().numpy() > 0)
    ct = {k: v * (both[:, None] if v.ndim == 2 else both) for k, v in ct.items()}
    loss = ((xys * cu(ct["v_xys"])).sum() + (depths * cu(ct["v_depths"])).sum() + (pix_vels * cu(ct["v_pix_vels"])).sum()
            + (conics * cu(ct["v_conics"])).sum() + (comp * cu(ct["v_compensation"])).sum())
    loss.backward()
    grads, _ = TO.project_vjp(inputs, {k: t64(v) for k, v in ct.items()}, **cfg)
    grad_close(leaves["means"].grad, grads["v_means"].numpy(), 2e-3, "v_means")
    grad_close(leaves["scales"].grad, grads["v_scales"].numpy(), 2e-3, "v_scales")
    grad_close(leaves["quats"].grad, grads["v_quats"].numpy(), 2e-3, "v_quats")
    grad_close(leaves["lin"].grad, grads["v_lin_vel"].numpy(), 2e-3, "v_lin_vel")
    grad_close(leaves["ang"].grad, grads["v_ang_vel"].numpy(), 2e-3, "v_ang_vel")
    grad_close(leaves["vm"].grad, grads["v_viewmat"].numpy()[:3], 2e-3, "v_viewmat (exact)")


@pytest.mark.parametrize("method", ["poly", "fast"])
@pytest.mark.parametrize("k,deg_use", [(25, 4), (16, 3), (16, 1), (9, 2), (4, 0), (1, 0)])
def test_sh_forward_backward_vs_oracle(method, k, deg_use):
    g = np.random.default_rng(k * 10 + deg_use)
    n = 5000 + 37
    dirs = g.standard_normal((n, 3)).astype(np.float32)
    coeffs = g.standard_normal((n, k, 3)).astype(np.float32)
    v = g.standard_normal((n, 3)).astype(np.float32)
    c = cu(coeffs).requires_grad_(True)
    col = spherical_harmonics(deg_use, cu(dirs), c, method)
    close(col, O.sh_forward(method, deg_use, dirs, coeffs), 2e-6, 1e-5, "sh colors")
    col.backward(cu(v))
    deg = {1: 0, 4: 1, 9: 2, 16: 3, 25: 4}[k]
    close(c.grad, O.sh_backward(method, deg, deg_use, dirs, v), 1e-6, 1e-5, "sh v_coeffs")


def test_sh_golden_reference(golden):
    g = golden("sh.npz")
    for method in ("poly", "fast"):
        for deg in range(5):
            col = spherical_harmonics(deg, cu(g["dirs"]), cu(g["coeffs"]), method)
            close(col, g[f"{method}_{deg}"], 2e-5, 1e-5, f"sh golden {method} {deg}")


@pytest.mark.parametrize("name,n,motion", [("c1", None, False), ("c2", 30000, True), ("c2", None, True)])
def test_binning_bit_exact_vs_oracle(name, n, motion):
    """Keys, stable sort order (ties by Gaussian id), tile ranges and the phantom zero-key slots: exact."""
    d = scene_np(name, n=n, motion=motion)
    o = oracle_project(d)
    b = O.bin_and_sort(o["xys"], o["depths"], o["radii"], o["num_tiles_hit"], d["H"], d["W"], d["bw"])
    tb = ((d["W"] + 15) // 16, (d["H"] + 15) // 16, 1)
    m, cum = gsplat.compute_cumulative_intersects(cu(o["num_tiles_hit"]))
    assert m == b["num_intersects"] and np.array_equal(cum.cpu().numpy(), b["cum_tiles_hit"])
    out = gsplat.bin_and_sort_gaussians(d["N"], m, cu(o["xys"]), cu(o["depths"]), cu(o["radii"]), cum, tb, 16)
    names = ["isect_ids", "gaussian_ids", "isect_ids_sorted", "gaussian_ids_sorted", "tile_bins"]
    for t, k in zip(out, names):
        assert np.array_equal(t.cpu().numpy(), b[k]), k
    if motion:
        assert (b["isect_ids"] == 0).sum() > 0  # the blur-inflated float radius really produces phantom slots


@pytest.mark.parametrize("name,n,motion,H,W,bw", [("c1", None, False, None, None, 16), ("c2", 30000, True, None, None, 16),
                                                   ("c2", None, True, None, None, 16), ("c2", 5000, True, 40, 56, 8),
                                                   ("c1", 3000, False, 9, 7, 16), ("c4", 60000, True, None, None, 16)])
def test_fused_two_level_binning_is_order_identical(name, n, motion, H, W, bw):
    """gsplat.cuda.bin_tiles (depth sort, emit, tile sort) == stable sort of the reference's 64-bit keys, phantom
    slots included: gaussian_ids_sorted and tile_bins bit for bit."""
    d = scene_np(name, n=n, motion=motion, H=H, W=W)
    d["bw"] = bw
    o = oracle_project(d)
    b = O.bin_and_sort(o["xys"], o["depths"], o["radii"], o["num_tiles_hit"], d["H"], d["W"], bw)
    tb = ((d["W"] + bw - 1) // bw, (d["H"] + bw - 1) // bw, 1)
    ids, bins = _C.bin_tiles(b["num_intersects"], cu(o["xys"]), cu(o["depths"]), cu(o["radii"]), cu(o["num_tiles_hit"]), tb, bw)
    assert np.array_equal(bins.cpu().numpy(), b["tile_bins"])
    assert np.array_equal(ids.cpu().numpy(), b["gaussian_ids_sorted"])


@pytest.mark.parametrize("name,n,motion,S,rs,exposure,H,W", [
    ("c2", None, True, 5, 0.0, 1 / 60, None, None), ("c2", 50000, True, 10, 1 / 50, 1 / 60, 256, 320),
    ("c1", None, False, 1, 0.0, 0.0, None, None), ("c4", 80000, True, 5, 1 / 50, 1 / 60, 480, 640)])
def test_culled_binning_changes_no_output(name, n, motion, S, rs, exposure, H, W):
    """rasterize_gaussians bins with the culled two-level sort.  Dropped (tile, Gaussian) pairs can never colour a pixel,
    so against a blend of the reference's FULL lists: image and transmittances are bit-identical, every last
    contributor is the same Gaussian, each culled tile list is an order-preserving sub-list, gradients agree."""
    d = scene_np(name, n=n, motion=motion, S=S, rs=rs, exposure=exposure, H=H, W=W)
    o = oracle_project(d)
    col = cu(oracle_colors(d))
    opac = cu((d["opacity"][:, 0] * o["compensation"])[:, None].astype(np.float32))
    xys, depths, pv, radii, conics, nth = (cu(o[k]) for k in ("xys", "depths", "pix_vels", "radii", "conics", "num_tiles_hit"))
    bg = cu(d["background"])
    tb = ((d["W"] + 15) // 16, (d["H"] + 15) // 16, 1)
    m, cum = gsplat.compute_cumulative_intersects(nth)
    ids_full, bins_full = gsplat.bin_and_sort_gaussians(d["N"], m, xys, depths, radii, cum, tb, 16)[3:5]
    img_f, Ts_f, fi_f = _C.rasterize_forward(tb, (16, 16, 1), (d["W"], d["H"], 1), d["S"], ids_full, bins_full, xys, pv, d["rs"],
                                             d["exposure"], conics, col, opac, bg)
    packed = _C.pack_records(xys, pv, conics, col, opac)
    total, ids_c, bins_c = _C.bin_cull(packed, depths, radii, nth, d["H"], d["W"], 16, d["S"], d["rs"], d["exposure"])
    assert total == m and ids_c.numel() <= m
    img_c, Ts_c, fi_c = _C.blend_forward_packed(d["H"], d["W"], 16, d["S"], ids_c, bins_c, packed, d["rs"], d["exposure"], bg)
    assert torch.equal(img_c, img_f) and torch.equal(Ts_c, Ts_f)
    hit = Ts_f < 1.0  # pixel-samples with at least one contributor
    assert torch.equal(ids_c[fi_c[hit].long()], ids_full[fi_f[hit].long()])
    # sub-list property on a sample of tiles
    bf, bc, idf, idc = bins_full.cpu().numpy(), bins_c.cpu().numpy(), ids_full.cpu().numpy(), ids_c.cpu().numpy()
    for t in np.random.default_rng(0).choice(bf.shape[0], size=min(40, bf.shape[0]), replace=False):
        full, cul = idf[bf[t, 0]:bf[t, 1]], idc[bc[t, 0]:bc[t, 1]]
        it = iter(full.tolist())
        assert all(any(x == y for y in it) for x in cul.tolist()), f"tile {t}: culled list is not a sub-list"
    g = np.random.default_rng(3)
    v_out = cu(g.standard_normal((d["H"], d["W"], 3)).astype(np.float32))
    v_alpha = cu(g.standard_normal((d["H"], d["W"])).astype(np.float32))
    gf = _C.rasterize_backward(d["H"], d["W"], 16, d["S"], ids_full, bins_full, xys, pv, d["rs"], d["exposure"], conics, col, opac,
                               bg, Ts_f, fi_f, v_out, v_alpha)
    gc = _C.blend_backward_packed(d["N"], d["H"], d["W"], 16, d["S"], ids_c, bins_c, packed, d["rs"], d["exposure"], bg, Ts_c, fi_c,
                                  v_out, v_alpha)
    for a, b_, k in zip(gc, gf, ["v_xy", "v_xy_abs", "v_pix_vels", "v_conic", "v_colors", "v_opacity"]):
        grad_close(a, b_.cpu().numpy(), 1e-5, k, rtol=1e-4, outliers=1e-5)  # only the atomic order differs


def test_culled_binning_mask_cap_fallback():
    """The count pass keeps a survival mask per 32-tile chunk for the emit pass; chunks beyond the mask capacity are
    re-tested instead.  Re-run the culled-vs-full identity test in a process whose capacity is 7 chunks (and 0)."""
    import subprocess
    import sys
    for cap in ("7", "0"):
        env = dict(os.environ, B200_CULL_MASK_CAP=cap)
        r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k",
                            "test_culled_binning_changes_no_output and c4"], env=env, capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("cfg", [
    ("c2", None, True, 5, 0.0, 1 / 60, None, None), ("c2", 60000, True, 10, 1 / 50, 1 / 60, 256, 320),
    ("c4", 80000, True, 5, 1 / 50, 1 / 60, 480, 640), ("c3_rs", 100000, True, 1, 1 / 50, 0.0, 360, 640),
    ("c1", None, False, 1, 0.0, 0.0, None, None)])
def test_closed_form_tile_test_keeps_a_superset_of_the_per_sample_test(cfg, tmp_path):
    """The binning's tile test covers the S blur samples with one interval test (binning.cu: may_touch_rect_closed_form).
    It must keep every (tile, Gaussian) pair the sample-by-sample test keeps (B200_CULL_PER_SAMPLE=1, run in a separate
    process because the switch is read once), in the same order, and at most a sliver more."""
    import subprocess
    import sys
    import cull_dump
    out = str(tmp_path / "per_sample.npz")
    env = dict(os.environ, B200_CULL_PER_SAMPLE="1")
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cull_dump.py"), out, repr(cfg)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    ref = np.load(out)
    total, ids, bins = cull_dump.culled_lists(*cfg)
    assert total == int(ref["total"])
    assert ids.size >= ref["ids"].size and ids.size <= 1.05 * ref["ids"].size + 64, (ids.size, ref["ids"].size)
    rb, ri = ref["bins"], ref["ids"]
    for t in range(bins.shape[0]):
        mine, theirs = ids[bins[t, 0]:bins[t, 1]], ri[rb[t, 0]:rb[t, 1]]
        if theirs.size == 0:
            continue
        it = iter(mine.tolist())
        assert all(any(x == y for y in it) for x in theirs.tolist()), f"tile {t}: per-sample list is not a sub-list"


def test_map_and_bins_golden_reference(golden):
    g = golden("map_bins.npz")
    H, W, bw = int(g["H"]), int(g["W"]), int(g["bw"])
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    m, cum = gsplat.compute_cumulative_intersects(cu(g["num_tiles_hit"]))
    out = gsplat.bin_and_sort_gaussians(300, m, cu(g["xys"]), cu(g["depths"]), cu(g["radii"]), cum, tb, bw)
    for t, k in zip(out, ["isect_ids", "gaussian_ids", "isect_ids_sorted", "gaussian_ids_sorted", "tile_bins"]):
        assert np.array_equal(t.cpu().numpy(), g[k]), k
    conics, radii = gsplat.compute_cov2d_bounds(cu(golden("cov2d_bounds.npz")["cov2d"]))
    gg = golden("cov2d_bounds.npz")
    close(conics.cpu().numpy()[gg["valid"]], gg["conics"][gg["valid"]], 5e-4, 1e-5, "conics")
    close(radii.cpu().numpy()[gg["valid"], 0], gg["radii"][gg["valid"]], 5e-4, 0, "radii")


BLEND_CASES = [
    ("c1", None, False, None, None, None),          # BASELINE config 1: 10k, 256^2, S=1, static
    ("c2", 40000, True, 5, 0.0, 1 / 60),            # blur only
    ("c2", 40000, True, 1, 1 / 50, 0.0),            # rolling shutter only
    ("c2", 40000, True, 10, 1 / 50, 1 / 60),        # both, kernel maximum S
    ("c2", 40000, True, 3, 1 / 50, 1 / 60),
]


def _gpu_blend_from_oracle_inputs(d, r, S):
    b = r["bins"]
    tb = ((d["W"] + 15) // 16, (d["H"] + 15) // 16, 1)
    return _C.rasterize_forward(tb, (16, 16, 1), (d["W"], d["H"], 1), S, cu(b["gaussian_ids_sorted"]), cu(b["tile_bins"]),
                                cu(r["proj"]["xys"]), cu(r["proj"]["pix_vels"]), d["rs"], d["exposure"],
                                cu(r["proj"]["conics"]), cu(r["colors"]), cu(r["opac"]), cu(d["background"]))


@pytest.mark.parametrize("name,n,motion,S,rs,exposure", BLEND_CASES)
def test_blend_forward_vs_oracle(name, n, motion, S, rs, exposure):
    d = scene_np(name, n=n, motion=motion, S=S, rs=rs, exposure=exposure, H=256 if n else None, W=320 if n else None)
    r = oracle_render(d)
    img, Ts, fi = _gpu_blend_from_oracle_inputs(d, r, d["S"])
    assert frac_mismatch(fi.cpu().numpy(), r["final_idx"]) <= 2e-4, "final_idx"
    same = fi.cpu().numpy() == r["final_idx"]
    close(Ts.cpu().numpy()[same], r["final_Ts"][same], 1e-6, 2e-5, "final_Ts", outliers=1e-3, outlier_atol=1e-2)
    px_same = same.all(-1)
    close(img.cpu().numpy()[px_same], r["img"][px_same], 2e-5, 1e-5, "out_img", outliers=1e-3, outlier_atol=1e-2)
    mse = float(((img.cpu().numpy().astype(np.float64) - r["img"]) ** 2).mean())
    assert mse < 1e-8, f"PSNR {10 * np.log10(1.0 / max(mse, 1e-30)):.1f} dB (bar: 80 dB)"


@pytest.mark.parametrize("name,n,motion,S,rs,exposure", BLEND_CASES)
def test_blend_backward_vs_oracle(name, n, motion, S, rs, exposure):
    d = scene_np(name, n=n, motion=motion, S=S, rs=rs, exposure=exposure, H=256 if n else None, W=320 if n else None)
    r = oracle_render(d)
    g = np.random.default_rng(11)
    v_out = g.standard_normal((d["H"], d["W"], 3)).astype(np.float32)
    v_alpha = g.standard_normal((d["H"], d["W"])).astype(np.float32)
    b = r["bins"]
    # feed the ORACLE's forward state so only the backward kernel is under test
    ref = O.rasterize_backward(d["H"], d["W"], 16, d["S"], b["gaussian_ids_sorted"], b["tile_bins"], r["proj"]["xys"],
                               r["proj"]["pix_vels"], d["rs"], d["exposure"], r["proj"]["conics"], r["colors"], r["opac"],
                               d["background"], r["final_Ts"], r["final_idx"], v_out, v_alpha)
    out = _C.rasterize_backward(d["H"], d["W"], 16, d["S"], cu(b["gaussian_ids_sorted"]), cu(b["tile_bins"]),
                                cu(r["proj"]["xys"]), cu(r["proj"]["pix_vels"]), d["rs"], d["exposure"],
                                cu(r["proj"]["conics"]), cu(r["colors"]), cu(r["opac"]), cu(d["background"]),
                                cu(r["final_Ts"]), cu(r["final_idx"]), cu(v_out), cu(v_alpha))
    for t, k in zip(out, ["v_xy", "v_xy_abs", "v_pix_vels", "v_conic", "v_colors", "v_opacity"]):
        grad_close(t, ref[k], 1e-3, k)


def test_end_to_end_autograd_vs_oracle_chain():
    """project -> SH -> rasterize through the public operators, loss.backward(), vs the oracle chain."""
    d = scene_np("c2", n=30000, H=192, W=256)
    r = oracle_render(d)
    means = cu(d["means"]).requires_grad_(True)
    scales = cu(d["scales"]).requires_grad_(True)
    quats = cu(d["quats"]).requires_grad_(True)
    sh = cu(d["sh"]).requires_grad_(True)
    opac = cu(d["opacity"]).requires_grad_(True)
    bg = cu(d["background"]).requires_grad_(True)
    xys, depths, pix_vels, radii, conics, comp, nth, cov3d = gpu_project(d, means=means, scales=scales, quats=quats)
    xys.retain_grad()
    rgbs = torch.clamp(spherical_harmonics(3, means.detach() - cu(d["cam_pos"]), sh) + 0.5, min=0.0)
    img, alpha = rasterize_gaussians(xys, depths, pix_vels, radii, conics, nth, rgbs, opac * comp[:, None], d["H"], d["W"],
                                     16, background=bg, return_alpha=True, rolling_shutter_time=d["rs"],
                                     exposure_time=d["exposure"], blur_samples=d["S"])
    close(img, r["img"], 5e-5, 1e-4, "e2e image", outliers=1e-3, outlier_atol=1e-2)
    close(alpha, 1 - r["final_Ts"].mean(-1), 5e-5, 1e-4, "e2e alpha", outliers=1e-3, outlier_atol=1e-2)
    g = np.random.default_rng(5)
    v_out = g.standard_normal(r["img"].shape).astype(np.float32)
    v_alpha = g.standard_normal(r["img"].shape[:2]).astype(np.float32)
    ((img * cu(v_out)).sum() + (alpha * cu(v_alpha)).sum()).backward()
    b = r["bins"]
    # d(alpha)/d(final_Ts) = -1/S is folded into the kernel through v_out_alpha exactly like the reference
    rb = O.rasterize_backward(d["H"], d["W"], 16, d["S"], b["gaussian_ids_sorted"], b["tile_bins"], r["proj"]["xys"],
                              r["proj"]["pix_vels"], d["rs"], d["exposure"], r["proj"]["conics"], r["colors"], r["opac"],
                              d["background"], r["final_Ts"], r["final_idx"], v_out, v_alpha)
    grad_close(xys.grad, rb["v_xy"], 2e-3, "v_xy through autograd")
    grad_close(xys.absgrad, rb["v_xy_abs"], 2e-3, "xys.absgrad side channel")
    # SH coefficient gradient: clamp(rgb + 0.5, 0) gate, then basis outer product
    gate = (r["colors"] > 0).astype(np.float32)
    ref_sh = O.sh_backward("fast", 3, 3, d["means"] - d["cam_pos"][None], rb["v_colors"] * gate)
    grad_close(sh.grad, ref_sh, 2e-3, "v_sh")
    grad_close(opac.grad, rb["v_opacity"][:, 0:1] * r["proj"]["compensation"][:, None], 2e-3, "v_opacity logit side")
    ref_bg = (v_out.reshape(-1, 3).astype(np.float64) * r["final_Ts"].mean(-1).reshape(-1, 1)).sum(0)
    grad_close(bg.grad, ref_bg, 2e-3, "v_background")
    pb = O.project_backward(d["means"], d["scales"], 1.0, d["quats"], d["lin_vel"], d["ang_vel"], d["rs"], d["exposure"],
                            d["viewmat"], d["fx"], d["fy"], r["proj"]["cov3d"], r["proj"]["radii"], r["proj"]["conics"],
                            r["proj"]["compensation"], rb["v_xy"], np.zeros(d["N"], np.float32), rb["v_pix_vels"],
                            rb["v_conic"], (rb["v_opacity"][:, 0] * d["opacity"][:, 0]).astype(np.float32))
    grad_close(means.grad, pb["v_mean3d"], 5e-3, "v_means e2e")
    grad_close(scales.grad, pb["v_scale"], 5e-3, "v_scales e2e")
    grad_close(quats.grad, pb["v_quat"], 5e-3, "v_quats e2e")


def test_full_size_c2_properties():
    """BASELINE config 2 at full size (300k Gaussians, 800x800, S=5): size-independent properties."""
    d = scene_np("c2")
    out = gpu_project(d)
    xys, depths, pix_vels, radii, conics, comp, nth, cov3d = out
    col = cu(oracle_colors(d))
    opac = cu(d["opacity"]) * comp[:, None]
    bg = cu(d["background"])
    kw = dict(background=bg, return_alpha=True, rolling_shutter_time=d["rs"], exposure_time=d["exposure"], blur_samples=d["S"])
    img, alpha = rasterize_gaussians(xys, depths, pix_vels, radii, conics, nth, col, opac, d["H"], d["W"], 16, **kw)
    assert torch.isfinite(img).all() and (alpha >= -1e-6).all() and (alpha <= 1 + 1e-6).all()
    # determinism (stable sort, no atomics in the forward)
    img2, alpha2 = rasterize_gaussians(xys, depths, pix_vels, radii, conics, nth, col, opac, d["H"], d["W"], 16, **kw)
    assert torch.equal(img, img2) and torch.equal(alpha, alpha2)
    # linearity in (colours, background); alpha independent of colour
    kw2 = dict(kw, background=2 * bg)
    img3, alpha3 = rasterize_gaussians(xys, depths, pix_vels, radii, conics, nth, 2 * col, opac, d["H"], d["W"], 16, **kw2)
    assert torch.equal(alpha, alpha3)
    torch.testing.assert_close(img3, 2 * img, rtol=1e-5, atol=1e-6)
    # sortedness: within each tile the staged depths are non-decreasing
    m, cum = gsplat.compute_cumulative_intersects(nth)
    tb = (50, 50, 1)
    isect, gids, isect_s, gids_s, bins = gsplat.bin_and_sort_gaussians(d["N"], m, xys, depths, radii, cum, tb, 16)
    assert (isect_s[1:] >= isect_s[:-1]).all()
    assert int((bins[:, 1] - bins[:, 0]).sum()) == m
    # zero velocity: S samples collapse onto the single-sample render
    z = torch.zeros_like(pix_vels)
    a1 = rasterize_gaussians(xys, depths, z, radii, conics, nth, col, opac, d["H"], d["W"], 16, background=bg)
    a5 = rasterize_gaussians(xys, depths, z, radii, conics, nth, col, opac, d["H"], d["W"], 16, background=bg,
                             exposure_time=d["exposure"], blur_samples=5)
    torch.testing.assert_close(a5, a1, rtol=1e-5, atol=2e-6)


def test_empty_and_error_paths():
    d = scene_np("c1", n=64)
    # everything behind the camera -> no intersections -> reference's empty-render branch (rasterize.py:136-144)
    vm = d["viewmat"].copy()
    vm[2, 3] = -1e4
    out = project_gaussians(cu(d["means"]), cu(d["scales"]), 1.0, cu(d["quats"]), None, None, 0, 0, cu(vm), d["fx"], d["fy"],
                            d["cx"], d["cy"], d["H"], d["W"], 16)
    xys, depths, pix_vels, radii, conics, comp, nth, cov3d = out
    assert int(nth.sum()) == 0 and int(radii.sum()) == 0
    bg = cu(d["background"])
    img, alpha = rasterize_gaussians(xys, depths, pix_vels, radii, conics, nth, cu(oracle_colors(d)), cu(d["opacity"]),
                                     d["H"], d["W"], 16, background=bg, return_alpha=True)
    assert torch.allclose(img, bg.expand_as(img)) and (alpha == 1).all()
    with pytest.raises(AssertionError):
        project_gaussians(cu(d["means"]), cu(d["scales"]), 1.0, cu(d["quats"]), None, None, 0, 0, cu(vm), 1, 1, 0, 0, 8, 8, 17)
    # un-normalised quaternions: the reference asserts inside project_gaussians (project_gaussians.py:69); here the
    # kernel raises a device flag and the same AssertionError surfaces at the next host sync of the path
    bad = project_gaussians(cu(d["means"]), cu(d["scales"]), 1.0, cu(d["quats"]) * 2, None, None, 0, 0, cu(d["viewmat"]),
                            d["fx"], d["fy"], d["cx"], d["cy"], d["H"], d["W"], 16)
    with pytest.raises(AssertionError, match="quats must be normalized"):
        gsplat.compute_cumulative_intersects(bad[6])
    from gsplat import _lib
    _lib.SYNC_CHECKS = True
    try:
        with pytest.raises(AssertionError, match="quats must be normalized"):
            project_gaussians(cu(d["means"]), cu(d["scales"]), 1.0, cu(d["quats"]) * 2, None, None, 0, 0, cu(vm), 1, 1, 0, 0, 8, 8, 16)
    finally:
        _lib.SYNC_CHECKS = False
    # a shorter-than-unit quaternion passes, exactly like the reference's one-sided check
    ok = project_gaussians(cu(d["means"]), cu(d["scales"]), 1.0, cu(d["quats"]) * 0.5, None, None, 0, 0, cu(d["viewmat"]),
                           d["fx"], d["fy"], d["cx"], d["cy"], d["H"], d["W"], 16)
    gsplat.compute_cumulative_intersects(ok[6])
    with pytest.raises(RuntimeError, match="unsupported blur size"):
        rasterize_gaussians(xys, depths, pix_vels, radii + 1, conics, nth + 1, cu(oracle_colors(d)), cu(d["opacity"]),
                            d["H"], d["W"], 16, exposure_time=0.1, blur_samples=11)
    with pytest.raises(ValueError):
        rasterize_gaussians(xys[:, :1], depths, pix_vels, radii, conics, nth, cu(oracle_colors(d)), cu(d["opacity"]), 8, 8, 16)


@pytest.mark.parametrize("bw,H,W", [(16, 100, 75), (8, 40, 56), (5, 33, 17), (2, 9, 7)])
def test_ragged_sizes_and_small_tiles(bw, H, W):
    """Image sizes that are not tile multiples and every supported block_width family."""
    d = scene_np("c1", n=3000, H=H, W=W)
    d["bw"] = bw
    r = oracle_render(d)
    b = r["bins"]
    tb = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    img, Ts, fi = _C.rasterize_forward(tb, (bw, bw, 1), (W, H, 1), 1, cu(b["gaussian_ids_sorted"]), cu(b["tile_bins"]),
                                       cu(r["proj"]["xys"]), cu(r["proj"]["pix_vels"]), 0.0, 0.0, cu(r["proj"]["conics"]),
                                       cu(r["colors"]), cu(r["opac"]), cu(d["background"]))
    assert frac_mismatch(fi.cpu().numpy(), r["final_idx"]) <= 1e-3
    close(img, r["img"], 5e-5, 1e-4, "ragged image", outliers=1e-3, outlier_atol=1e-2)


@pytest.mark.parametrize("S", list(range(1, 11)))
def test_every_blur_sample_count_vs_oracle(S):
    """All ten template instantiations of the blend kernels (MAX_BLUR_SAMPLES = 10, helpers.cuh:222), blur + RS."""
    d = scene_np("c2", n=6000, S=S, rs=1 / 50, exposure=1 / 60 if S > 1 else 0.0, H=64, W=96)
    r = oracle_render(d)
    img, Ts, fi = _gpu_blend_from_oracle_inputs(d, r, S)
    assert frac_mismatch(fi.cpu().numpy(), r["final_idx"]) <= 1e-3
    close(img, r["img"], 5e-5, 1e-4, f"S={S} image", outliers=1e-3, outlier_atol=1e-2)
    g = np.random.default_rng(S)
    v_out = g.standard_normal((d["H"], d["W"], 3)).astype(np.float32)
    v_alpha = g.standard_normal((d["H"], d["W"])).astype(np.float32)
    b = r["bins"]
    ref = O.rasterize_backward(d["H"], d["W"], 16, S, b["gaussian_ids_sorted"], b["tile_bins"], r["proj"]["xys"], r["proj"]["pix_vels"],
                               d["rs"], d["exposure"], r["proj"]["conics"], r["colors"], r["opac"], d["background"], r["final_Ts"],
                               r["final_idx"], v_out, v_alpha)
    out = _C.rasterize_backward(d["H"], d["W"], 16, S, cu(b["gaussian_ids_sorted"]), cu(b["tile_bins"]), cu(r["proj"]["xys"]),
                                cu(r["proj"]["pix_vels"]), d["rs"], d["exposure"], cu(r["proj"]["conics"]), cu(r["colors"]), cu(r["opac"]),
                                cu(d["background"]), cu(r["final_Ts"]), cu(r["final_idx"]), cu(v_out), cu(v_alpha))
    for t, k in zip(out, ["v_xy", "v_xy_abs", "v_pix_vels", "v_conic", "v_colors", "v_opacity"]):
        grad_close(t, ref[k], 1e-3, f"S={S} {k}")


@pytest.mark.parametrize("bw,H,W,S", [(8, 40, 56, 3), (5, 33, 17, 4), (13, 50, 41, 2)])
def test_small_tiles_with_blur_and_rolling_shutter(bw, H, W, S):
    """block_width != 16 (one pixel per lane kernels) with motion: through the public operators vs the oracle chain."""
    d = scene_np("c2", n=4000, H=H, W=W, S=S, rs=1 / 50, exposure=1 / 60)
    d["bw"] = bw
    r = oracle_render(d)
    xys, depths, pix_vels, radii, conics, comp, nth, cov3d = gpu_project(d)
    img, alpha = rasterize_gaussians(xys, depths, pix_vels, radii, conics, nth, cu(r["colors"]), cu(d["opacity"]) * comp[:, None],
                                     H, W, bw, background=cu(d["background"]), return_alpha=True, rolling_shutter_time=d["rs"],
                                     exposure_time=d["exposure"], blur_samples=S)
    close(img, r["img"], 5e-5, 1e-4, "image", outliers=2e-3, outlier_atol=1e-2)
    close(alpha, 1 - r["final_Ts"].mean(-1), 5e-5, 1e-4, "alpha", outliers=2e-3, outlier_atol=1e-2)


@pytest.mark.parametrize("name", ["c3_rs", "c3_rs10", "c4"])
def test_full_size_large_configs_culled_vs_full_lists(name):
    """BASELINE configs 3 and 4 at full size (500k / 1.5M Gaussians, up to 1920x1440, up to 1.2e8 reference
    intersections): the culled path reproduces the blend of the reference's full lists bit for bit."""
    from gsplat import synthetic

    sc = synthetic.make_scene(name, device="cuda")
    cam = sc["cameras"][0]
    N, H, W = sc["N"], sc["H"], sc["W"]
    S = sc["blur_samples"] if sc["exposure_time"] > 0 else 1
    rs, ex = sc["rolling_shutter_time"], sc["exposure_time"]
    q = sc["quats"] / sc["quats"].norm(dim=-1, keepdim=True)
    xys, depths, pv, radii, conics, comp, nth, _ = project_gaussians(sc["means"], sc["log_scales"].exp(), 1, q, cam["lin_vel"],
                                                                  cam["ang_vel"], rs, ex, cam["viewmat"], cam["fx"], cam["fy"],
                                                                  cam["cx"], cam["cy"], H, W, 16)
    col = torch.rand(N, 3, device="cuda")
    opac = torch.sigmoid(sc["opacity_logit"]) * comp[:, None]
    bg = sc["background"]
    tb = ((W + 15) // 16, (H + 15) // 16, 1)
    m, cum = gsplat.compute_cumulative_intersects(nth)
    ids_full, bins_full = _C.bin_tiles(m, xys, depths, radii, nth, tb, 16)
    img_f, Ts_f, fi_f = _C.rasterize_forward(tb, (16, 16, 1), (W, H, 1), S, ids_full, bins_full, xys, pv, rs, ex, conics, col, opac, bg)
    img_c, alpha_c = rasterize_gaussians(xys, depths, pv, radii, conics, nth, col, opac, H, W, 16, background=bg, return_alpha=True,
                                         rolling_shutter_time=rs, exposure_time=ex, blur_samples=S)
    assert torch.equal(img_c, img_f)
    # alpha is written by the blend kernel as 1 - (sum_s T_s) / S; torch's mean() adds the S values in another order
    torch.testing.assert_close(alpha_c, 1 - Ts_f.mean(dim=-1), rtol=0, atol=2.5e-7)
    assert torch.isfinite(img_c).all()


def test_nd_rasterize_vs_oracle():
    d = scene_np("c1", n=4000, H=64, W=80)
    r = oracle_render(d)
    g = np.random.default_rng(2)
    C = 5
    cols = g.uniform(0, 1, (d["N"], C)).astype(np.float32)
    bgc = g.uniform(0, 1, C).astype(np.float32)
    b = r["bins"]
    ref = O.nd_rasterize_forward(d["H"], d["W"], 16, b["gaussian_ids_sorted"], b["tile_bins"], r["proj"]["xys"],
                                 r["proj"]["conics"], cols, r["opac"], bgc)
    tb = ((d["W"] + 15) // 16, (d["H"] + 15) // 16, 1)
    out = _C.nd_rasterize_forward(tb, (16, 16, 1), (d["W"], d["H"], 1), 1, cu(b["gaussian_ids_sorted"]), cu(b["tile_bins"]),
                                  cu(r["proj"]["xys"]), cu(r["proj"]["pix_vels"]), 0.0, 0.0, cu(r["proj"]["conics"]),
                                  cu(cols), cu(r["opac"]), cu(bgc))
    assert frac_mismatch(out[2].cpu().numpy(), ref[2]) <= 1e-3
    close(out[0], ref[0], 2e-2, 1e-2, "nd image (fp16 accumulators)")
    v_out = g.standard_normal((d["H"], d["W"], C)).astype(np.float32)
    v_alpha = g.standard_normal((d["H"], d["W"])).astype(np.float32)
    rb = O.nd_rasterize_backward(d["H"], d["W"], 16, b["gaussian_ids_sorted"], b["tile_bins"], r["proj"]["xys"],
                                 r["proj"]["conics"], cols, r["opac"], bgc, ref[1], ref[2], v_out, v_alpha)
    ob = _C.nd_rasterize_backward(d["H"], d["W"], 16, 1, cu(b["gaussian_ids_sorted"]), cu(b["tile_bins"]),
                                  cu(r["proj"]["xys"]), cu(r["proj"]["pix_vels"]), 0.0, 0.0, cu(r["proj"]["conics"]),
                                  cu(cols), cu(r["opac"]), cu(bgc), cu(ref[1]), cu(ref[2]), cu(v_out), cu(v_alpha))
    for t, k in zip([ob[0], ob[1], ob[3], ob[4], ob[5]], ["v_xy", "v_xy_abs", "v_conic", "v_colors", "v_opacity"]):
        grad_close(t, rb[k], 2e-2, "nd " + k)
    with pytest.raises(RuntimeError, match="blur not supported"):
        _C.nd_rasterize_forward(tb, (16, 16, 1), (d["W"], d["H"], 1), 2, cu(b["gaussian_ids_sorted"]), cu(b["tile_bins"]),
                                cu(r["proj"]["xys"]), cu(r["proj"]["pix_vels"]), 0.0, 0.1, cu(r["proj"]["conics"]),
                                cu(cols), cu(r["opac"]), cu(bgc))


def test_public_rasterize_nd_channels_and_uint8_colors():
    """rasterize_gaussians with C != 3 (N-channel kernels, fp16 accumulators) and with uint8 colours (rasterize.py:63-65)."""
    d = scene_np("c1", n=4000, H=64, W=80)
    r = oracle_render(d)
    g = np.random.default_rng(4)
    xys, depths, pv, radii, conics, nth = (cu(r["proj"][k]) for k in ("xys", "depths", "pix_vels", "radii", "conics", "num_tiles_hit"))
    opac = cu(r["opac"])
    cols5 = g.uniform(0, 1, (d["N"], 5)).astype(np.float32)
    bg5 = g.uniform(0, 1, 5).astype(np.float32)
    c5 = cu(cols5).requires_grad_(True)
    img5, alpha5 = rasterize_gaussians(xys, depths, pv, radii, conics, nth, c5, opac, d["H"], d["W"], 16, background=cu(bg5), return_alpha=True)
    b = r["bins"]
    ref = O.nd_rasterize_forward(d["H"], d["W"], 16, b["gaussian_ids_sorted"], b["tile_bins"], r["proj"]["xys"], r["proj"]["conics"], cols5,
                                 r["opac"], bg5)
    close(img5, ref[0], 2e-2, 1e-2, "5-channel image (fp16 accumulators)")
    close(alpha5, 1 - ref[1], 1e-5, 1e-5, "5-channel alpha", outliers=1e-3, outlier_atol=1e-2)
    img5.sum().backward()
    assert c5.grad is not None and torch.isfinite(c5.grad).all() and float(c5.grad.abs().sum()) > 0
    # uint8 colours are scaled to [0, 1] like the reference
    u8 = (g.uniform(0, 1, (d["N"], 3)) * 255).astype(np.uint8)
    img_u8 = rasterize_gaussians(xys, depths, pv, radii, conics, nth, cu(u8), opac, d["H"], d["W"], 16, background=cu(d["background"]))
    img_f = rasterize_gaussians(xys, depths, pv, radii, conics, nth, cu(u8.astype(np.float32) / 255), opac, d["H"], d["W"], 16,
                                background=cu(d["background"]))
    torch.testing.assert_close(img_u8, img_f, rtol=1e-6, atol=1e-6)  # torch divides by 255 as a reciprocal multiply
    # default background is ones (rasterize.py:71-74)
    img_def = rasterize_gaussians(xys, depths, pv, radii, conics, nth, cu(r["colors"]), opac, d["H"], d["W"], 16)
    img_one = rasterize_gaussians(xys, depths, pv, radii, conics, nth, cu(r["colors"]), opac, d["H"], d["W"], 16, background=torch.ones(3, device="cuda"))
    assert torch.equal(img_def, img_one)


def test_alpha_channel_from_the_blend_kernel_and_unused_output_cotangents():
    """return_alpha: the blend kernel writes 1 - mean_s(final_Ts) itself (rasterize.py:161-163 builds it from final_Ts
    with two torch passes); an output that the loss does not use arrives in backward as None and must behave exactly
    like the zero image the reference materialises (rasterize.py:217-218)."""
    d = scene_np("c2", n=30000, motion=True, S=5, rs=1 / 50, exposure=1 / 60, H=160, W=208)
    r = oracle_render(d)
    xys, depths, pv, radii, conics, nth = (cu(r["proj"][k]) for k in ("xys", "depths", "pix_vels", "radii", "conics", "num_tiles_hit"))
    kw = dict(background=cu(d["background"]), rolling_shutter_time=d["rs"], exposure_time=d["exposure"], blur_samples=d["S"])
    g = np.random.default_rng(9)
    w_img = cu(g.standard_normal((d["H"], d["W"], 3)).astype(np.float32))
    w_alpha = cu(g.standard_normal((d["H"], d["W"])).astype(np.float32))

    def run(mode):
        col, op = cu(r["colors"]).requires_grad_(True), cu(r["opac"]).requires_grad_(True)
        x = xys.clone().requires_grad_(True)
        img, alpha = rasterize_gaussians(x, depths, pv, radii, conics, nth, col, op, d["H"], d["W"], 16, return_alpha=True, **kw)
        loss = {"img": (img * w_img).sum(), "img+0alpha": (img * w_img).sum() + (alpha * 0).sum(),
                "alpha": (alpha * w_alpha).sum(), "0img+alpha": (img * 0).sum() + (alpha * w_alpha).sum()}[mode]
        loss.backward()
        return img.detach(), alpha.detach(), [t.grad.clone() for t in (x, col, op)]

    img, alpha, g_img = run("img")
    _, _, g_img0 = run("img+0alpha")
    _, _, g_a = run("alpha")
    _, _, g_a0 = run("0img+alpha")
    # alpha against the oracle's final_Ts (same tolerance as final_Ts itself)
    close(alpha, 1 - r["final_Ts"].mean(axis=-1), 2e-5, 1e-5, "alpha", outliers=1e-3, outlier_atol=1e-2)
    for a, b, k in zip(g_img, g_img0, ("v_xy", "v_colors", "v_opacity")):
        grad_close(a, b.cpu().numpy(), 1e-5, k + " (alpha unused)", rtol=1e-4, outliers=1e-5)
    for a, b, k in zip(g_a, g_a0, ("v_xy", "v_colors", "v_opacity")):
        grad_close(a, b.cpu().numpy(), 1e-5, k + " (image unused)", rtol=1e-4, outliers=1e-5)


def test_fused_l1_loss_vs_torch():
    """gsplat.losses.l1_loss == torch.abs(gt - pred).mean() (splatfacto.py:957) and its autograd cotangent, for sizes with
    and without a 4-float tail; the value is deterministic (same bits on every call)."""
    from gsplat.losses import l1_loss
    g = torch.Generator(device="cuda").manual_seed(5)
    for shape in ((800, 800, 3), (37, 53, 3), (1, 1, 3), (5,)):
        pred = torch.rand(shape, device="cuda", generator=g).requires_grad_(True)
        target = torch.rand(shape, device="cuda", generator=g)
        with torch.no_grad():
            target.view(-1)[0] = pred.view(-1)[0]  # an exact tie: torch.sign gives 0 there
        ref = (target - pred).abs().mean()
        (g_ref,) = torch.autograd.grad(ref * 3.0, pred)
        out = l1_loss(pred, target)
        (g_out,) = torch.autograd.grad(out * 3.0, pred)
        torch.testing.assert_close(out, ref, rtol=2e-6, atol=0)
        assert torch.equal(g_out, g_ref)
        assert torch.equal(l1_loss(pred, target), out)
    with pytest.raises(ValueError):
        l1_loss(torch.zeros(4, device="cuda"), torch.zeros(5, device="cuda"))
    with pytest.raises(RuntimeError):
        l1_loss(torch.zeros(4), torch.zeros(4))
