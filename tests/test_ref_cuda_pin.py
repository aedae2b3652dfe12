"""Pin the oracle (and the product) to the UNMODIFIED reference CUDA extension.

Two layers:
  * test_reference_cuda_live (gpu): loads oracle/_ref/gsplat_ref_csrc.so (built by oracle/build_ref.py from
    /root/reference in the build container; it travels to the GPU box), drives the reference kernels with the
    reference's own op sequence (cumsum -> map -> torch.sort -> gather -> bin edges -> rasterize), and compares
    them with the C oracle and with libb200splat.  With B200_WRITE_REFCUDA=1 it also writes the reference
    outputs to gpurun_out/refcuda_*.npz -- that is how tests/golden/refcuda_*.npz were produced.
  * test_refcuda_fixture_* : compare the oracle (CPU) and the product (GPU) with the committed fixtures, so
    the pin survives without the extension.
This is the only pin the blur / rolling-shutter blend has: the reference ships no test for it.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import oracle as O
from util_scene import cu, oracle_colors, scene_np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

CASES = {
    # tag: (scene, n, H, W, S, rs, exposure)
    "blur_rs": ("c2", 3000, 96, 128, 3, 1 / 50, 1 / 60),
    "static": ("c1", 3000, 96, 128, 1, 0.0, 0.0),
    "blur10": ("c2", 2000, 64, 80, 10, 0.0, 1 / 30),
}


def _case(tag):
    name, n, H, W, S, rs, ex = CASES[tag]
    d = scene_np(name, n=n, H=H, W=W, S=S, rs=rs, exposure=ex, motion=(rs > 0 or ex > 0))
    g = np.random.default_rng(hash(tag) % 1000)
    d["v_out"] = np.random.default_rng(7).standard_normal((H, W, 3)).astype(np.float32)
    d["v_alpha"] = np.random.default_rng(8).standard_normal((H, W)).astype(np.float32)
    return d


def _run_reference_cuda(ref, d):
    """The reference's own call sequence (project_gaussians.py:163-190, rasterize.py:106-208, utils.py:106-182)."""
    n, H, W, S = d["N"], d["H"], d["W"], d["S"]
    tb = ((W + 15) // 16, (H + 15) // 16, 1)
    means, scales, quats, vm = cu(d["means"]), cu(d["scales"]), cu(d["quats"]), cu(d["viewmat"])
    lin, ang = tuple(float(x) for x in d["lin_vel"]), tuple(float(x) for x in d["ang_vel"])
    cov3d, xys, depths, pix_vels, radii, conics, comp, nth = ref.project_gaussians_forward(
        n, means, scales, 1.0, quats, lin, ang, d["rs"], d["exposure"], vm, d["fx"], d["fy"], d["cx"], d["cy"], H, W, 16, 0.01)
    dirs = cu(d["means"] - d["cam_pos"][None])
    colors = torch.clamp(ref.compute_sh_forward("fast", n, 3, 3, dirs, cu(d["sh"])) + 0.5, min=0.0)
    opac = cu(d["opacity"]) * comp[:, None]
    cum = torch.cumsum(nth, dim=0, dtype=torch.int32)
    m = int(cum[-1].item())
    isect, gids = ref.map_gaussian_to_intersects(n, m, xys, depths, radii, cum, tb, 16)
    isect_s, order = torch.sort(isect, stable=True)  # reference: unstable torch.sort (ties unspecified)
    gids_s = torch.gather(gids, 0, order)
    bins = ref.get_tile_bin_edges(m, isect_s, tb)
    bg = cu(d["background"])
    img, Ts, fi = ref.rasterize_forward(tb, (16, 16, 1), (W, H, 1), S, gids_s, bins, xys, pix_vels, d["rs"], d["exposure"],
                                        conics, colors, opac, bg)
    bw = ref.rasterize_backward(H, W, 16, S, gids_s, bins, xys, pix_vels, d["rs"], d["exposure"], conics, colors, opac, bg,
                                Ts, fi, cu(d["v_out"]), cu(d["v_alpha"]))
    v_xy, v_xy_abs, v_pix, v_conic, v_colors, v_opac = bw
    v_comp = (v_opac[:, 0] * cu(d["opacity"])[:, 0]).contiguous()
    pb = ref.project_gaussians_backward(n, means, scales, 1.0, quats, lin, ang, d["rs"], d["exposure"], vm, d["fx"], d["fy"],
                                        d["cx"], d["cy"], H, W, cov3d, radii, conics, comp, v_xy, torch.zeros_like(depths),
                                        v_pix, v_conic, v_comp)
    v_sh = ref.compute_sh_backward("fast", n, 3, 3, dirs, (v_colors * (colors > 0)).contiguous())
    torch.cuda.synchronize()
    out = dict(cov3d=cov3d, xys=xys, depths=depths, pix_vels=pix_vels, radii=radii, conics=conics, compensation=comp,
               num_tiles_hit=nth, colors=colors, isect_ids=isect, gaussian_ids=gids, isect_ids_sorted=isect_s,
               gaussian_ids_sorted=gids_s, tile_bins=bins, img=img, final_Ts=Ts, final_idx=fi, v_xy=v_xy,
               v_xy_abs=v_xy_abs, v_pix_vels=v_pix, v_conic=v_conic, v_colors=v_colors, v_opacity=v_opac,
               v_mean3d=pb[2], v_scale=pb[3], v_quat=pb[4], v_sh=v_sh)
    return {k: v.detach().cpu().numpy() for k, v in out.items()}


ND_CHANNELS = 5


def _nd_inputs(d):
    g = np.random.default_rng(21)
    return dict(nd_colors=g.uniform(0, 1, (d["N"], ND_CHANNELS)).astype(np.float32),
                nd_background=g.uniform(0, 1, ND_CHANNELS).astype(np.float32),
                nd_v_out=g.standard_normal((d["H"], d["W"], ND_CHANNELS)).astype(np.float32),
                nd_v_alpha=g.standard_normal((d["H"], d["W"])).astype(np.float32))


def _run_reference_cuda_nd(ref, d, refout):
    """The reference's N-channel kernels (forward.cu:185-304, backward.cu:22-141) on the binning state of `refout`:
    C = 5 channels, fp16 accumulators, no blur (bindings.cu:506-682; rasterize.py:165,245 select them for C != 3)."""
    H, W = d["H"], d["W"]
    tb = ((W + 15) // 16, (H + 15) // 16, 1)
    nd = _nd_inputs(d)
    opac = cu((d["opacity"][:, 0] * refout["compensation"])[:, None].astype(np.float32))
    args = (cu(refout["gaussian_ids_sorted"]), cu(refout["tile_bins"]), cu(refout["xys"]), cu(refout["pix_vels"]), 0.0, 0.0,
            cu(refout["conics"]), cu(nd["nd_colors"]), opac, cu(nd["nd_background"]))
    img, Ts, fi = ref.nd_rasterize_forward(tb, (16, 16, 1), (W, H, 1), 1, *args)
    bw = ref.nd_rasterize_backward(H, W, 16, 1, *args, Ts, fi, cu(nd["nd_v_out"]), cu(nd["nd_v_alpha"]))
    torch.cuda.synchronize()
    out = dict(nd_img=img, nd_final_Ts=Ts, nd_final_idx=fi, nd_v_xy=bw[0], nd_v_xy_abs=bw[1], nd_v_conic=bw[3],
               nd_v_colors=bw[4], nd_v_opacity=bw[5])
    return {k: v.detach().cpu().numpy() for k, v in out.items()}


def _nd_check(refout, d, who, fwd, bwd):
    """An implementation of the N-channel blend against the reference kernels' outputs.  The reference accumulates the
    image in fp16 (forward.cu:277-283: `__half2float(__hadd(...))`-style adds), so the image tolerance is fp16's."""
    img, Ts, fi = fwd()
    assert (fi == refout["nd_final_idx"]).mean() > 0.999, (who, "nd final_idx")
    same = fi == refout["nd_final_idx"]
    assert np.abs(Ts[same] - refout["nd_final_Ts"][same]).max() < 2e-5, (who, "nd final_Ts")
    assert np.abs(img[same] - refout["nd_img"][same]).max() < 2e-2, (who, "nd image", np.abs(img[same] - refout["nd_img"][same]).max())
    g = bwd()
    for k in ("v_xy", "v_xy_abs", "v_conic", "v_colors", "v_opacity"):
        r = refout["nd_" + k]
        assert _rel(g[k].reshape(r.shape), r) < 2e-2, (who, "nd " + k, _rel(g[k].reshape(r.shape), r))


def _oracle_nd_checks(refout, d):
    nd = _nd_inputs(d)
    opac = (d["opacity"][:, 0] * refout["compensation"])[:, None].astype(np.float32)
    common = (d["H"], d["W"], 16, refout["gaussian_ids_sorted"], refout["tile_bins"], refout["xys"], refout["conics"],
              nd["nd_colors"], opac, nd["nd_background"])
    fwd = lambda: O.nd_rasterize_forward(*common)
    bwd = lambda: O.nd_rasterize_backward(*common, refout["nd_final_Ts"], refout["nd_final_idx"], nd["nd_v_out"], nd["nd_v_alpha"])
    _nd_check(refout, d, "oracle", fwd, bwd)


def _product_nd_checks(refout, d):
    import gsplat.cuda as _C

    H, W = d["H"], d["W"]
    tb = ((W + 15) // 16, (H + 15) // 16, 1)
    nd = _nd_inputs(d)
    opac = cu((d["opacity"][:, 0] * refout["compensation"])[:, None].astype(np.float32))
    args = (cu(refout["gaussian_ids_sorted"]), cu(refout["tile_bins"]), cu(refout["xys"]), cu(refout["pix_vels"]), 0.0, 0.0,
            cu(refout["conics"]), cu(nd["nd_colors"]), opac, cu(nd["nd_background"]))
    fwd = lambda: [t.cpu().numpy() for t in _C.nd_rasterize_forward(tb, (16, 16, 1), (W, H, 1), 1, *args)]

    def bwd():
        o = _C.nd_rasterize_backward(H, W, 16, 1, *args, cu(refout["nd_final_Ts"]), cu(refout["nd_final_idx"]),
                                     cu(nd["nd_v_out"]), cu(nd["nd_v_alpha"]))
        return {k: o[i].cpu().numpy() for k, i in (("v_xy", 0), ("v_xy_abs", 1), ("v_conic", 3), ("v_colors", 4), ("v_opacity", 5))}

    _nd_check(refout, d, "libb200splat", fwd, bwd)


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def _check_against(refout, d, who, project, colors_fn, blend_fwd, blend_bwd, proj_bwd, tol_img=3e-5, tol_grad=2e-3):
    """Compare an implementation (oracle or product) with reference-CUDA outputs `refout`."""
    p = project()
    m = refout["num_tiles_hit"] > 0
    assert (p["num_tiles_hit"] == refout["num_tiles_hit"]).mean() > 0.999, who
    assert (p["radii"] == refout["radii"]).mean() > 0.999, who
    both = m & (p["num_tiles_hit"] > 0)
    for k, tol in (("xys", 1e-5), ("depths", 1e-5), ("conics", 1e-4), ("compensation", 1e-4), ("pix_vels", 1e-4), ("cov3d", 1e-4)):
        assert _rel(p[k][both], refout[k][both]) < tol, (who, k, _rel(p[k][both], refout[k][both]))
    assert _rel(colors_fn(), refout["colors"]) < 1e-5, who
    # blend on the REFERENCE's binning state so indices are comparable
    img, Ts, fi = blend_fwd(refout)
    assert (fi == refout["final_idx"]).mean() > 0.9995, (who, (fi != refout["final_idx"]).mean())
    same = fi == refout["final_idx"]
    assert np.abs(Ts[same] - refout["final_Ts"][same]).max() < 2e-5, who
    assert np.abs(img[same.all(-1)] - refout["img"][same.all(-1)]).max() < tol_img, who
    g = blend_bwd(refout)
    for k in ("v_xy", "v_xy_abs", "v_pix_vels", "v_conic", "v_colors", "v_opacity"):
        assert _rel(g[k].reshape(refout[k].shape), refout[k]) < tol_grad, (who, k, _rel(g[k].reshape(refout[k].shape), refout[k]))
    pb = proj_bwd(refout)
    for k in ("v_mean3d", "v_scale", "v_quat"):
        assert _rel(pb[k], refout[k]) < tol_grad, (who, k, _rel(pb[k], refout[k]))


def _oracle_checks(refout, d):
    n, H, W, S = d["N"], d["H"], d["W"], d["S"]
    proj = lambda: O.project_forward(d["means"], d["scales"], 1.0, d["quats"], d["lin_vel"], d["ang_vel"], d["rs"],
                                     d["exposure"], d["viewmat"], d["fx"], d["fy"], d["cx"], d["cy"], H, W, 16)
    opac = lambda r: (d["opacity"][:, 0] * r["compensation"])[:, None].astype(np.float32)
    fwd = lambda r: O.rasterize_forward(H, W, 16, S, r["gaussian_ids_sorted"], r["tile_bins"], r["xys"], r["pix_vels"],
                                        d["rs"], d["exposure"], r["conics"], r["colors"], opac(r), d["background"])
    bwd = lambda r: O.rasterize_backward(H, W, 16, S, r["gaussian_ids_sorted"], r["tile_bins"], r["xys"], r["pix_vels"],
                                         d["rs"], d["exposure"], r["conics"], r["colors"], opac(r), d["background"],
                                         r["final_Ts"], r["final_idx"], d["v_out"], d["v_alpha"])
    pbw = lambda r: O.project_backward(d["means"], d["scales"], 1.0, d["quats"], d["lin_vel"], d["ang_vel"], d["rs"],
                                       d["exposure"], d["viewmat"], d["fx"], d["fy"], r["cov3d"], r["radii"], r["conics"],
                                       r["compensation"], r["v_xy"], np.zeros(n, np.float32), r["v_pix_vels"], r["v_conic"],
                                       (r["v_opacity"][:, 0] * d["opacity"][:, 0]).astype(np.float32))
    _check_against(refout, d, "oracle", proj, lambda: oracle_colors(d), fwd, bwd, pbw)
    # binning: oracle on the reference's projection outputs must reproduce keys / order / bins bit for bit
    b = O.bin_and_sort(refout["xys"], refout["depths"], refout["radii"], refout["num_tiles_hit"], H, W, 16)
    for k in ("isect_ids", "gaussian_ids", "isect_ids_sorted", "gaussian_ids_sorted", "tile_bins"):
        assert np.array_equal(b[k], refout[k]), ("oracle binning", k)


def _product_checks(refout, d):
    import gsplat
    import gsplat.cuda as _C

    n, H, W, S = d["N"], d["H"], d["W"], d["S"]
    tb = ((W + 15) // 16, (H + 15) // 16, 1)
    names = ["cov3d", "xys", "depths", "pix_vels", "radii", "conics", "compensation", "num_tiles_hit"]

    def proj():
        out = _C.project_gaussians_forward(n, cu(d["means"]), cu(d["scales"]), 1.0, cu(d["quats"]), tuple(d["lin_vel"]),
                                           tuple(d["ang_vel"]), d["rs"], d["exposure"], cu(d["viewmat"]), d["fx"], d["fy"],
                                           d["cx"], d["cy"], H, W, 16, 0.01)
        return {k: v.cpu().numpy() for k, v in zip(names, out)}

    cols = lambda: torch.clamp(_C.compute_sh_forward("fast", n, 3, 3, cu(d["means"] - d["cam_pos"][None]), cu(d["sh"])) + 0.5,
                               min=0).cpu().numpy()
    opac = lambda r: cu((d["opacity"][:, 0] * r["compensation"])[:, None].astype(np.float32))

    def fwd(r):
        o = _C.rasterize_forward(tb, (16, 16, 1), (W, H, 1), S, cu(r["gaussian_ids_sorted"]), cu(r["tile_bins"]), cu(r["xys"]),
                                 cu(r["pix_vels"]), d["rs"], d["exposure"], cu(r["conics"]), cu(r["colors"]), opac(r),
                                 cu(d["background"]))
        return [t.cpu().numpy() for t in o]

    def bwd(r):
        o = _C.rasterize_backward(H, W, 16, S, cu(r["gaussian_ids_sorted"]), cu(r["tile_bins"]), cu(r["xys"]), cu(r["pix_vels"]),
                                  d["rs"], d["exposure"], cu(r["conics"]), cu(r["colors"]), opac(r), cu(d["background"]),
                                  cu(r["final_Ts"]), cu(r["final_idx"]), cu(d["v_out"]), cu(d["v_alpha"]))
        return {k: t.cpu().numpy() for k, t in zip(["v_xy", "v_xy_abs", "v_pix_vels", "v_conic", "v_colors", "v_opacity"], o)}

    def pbw(r):
        o = _C.project_gaussians_backward(n, cu(d["means"]), cu(d["scales"]), 1.0, cu(d["quats"]), tuple(d["lin_vel"]),
                                          tuple(d["ang_vel"]), d["rs"], d["exposure"], cu(d["viewmat"]), d["fx"], d["fy"],
                                          d["cx"], d["cy"], H, W, cu(r["cov3d"]), cu(r["radii"]), cu(r["conics"]),
                                          cu(r["compensation"]), cu(r["v_xy"]), cu(np.zeros(n, np.float32)), cu(r["v_pix_vels"]),
                                          cu(r["v_conic"]), cu((r["v_opacity"][:, 0] * d["opacity"][:, 0]).astype(np.float32)))
        return dict(v_mean3d=o[2].cpu().numpy(), v_scale=o[3].cpu().numpy(), v_quat=o[4].cpu().numpy())

    _check_against(refout, d, "libb200splat", proj, cols, fwd, bwd, pbw)
    m, cum = gsplat.compute_cumulative_intersects(cu(refout["num_tiles_hit"]))
    out = gsplat.bin_and_sort_gaussians(n, m, cu(refout["xys"]), cu(refout["depths"]), cu(refout["radii"]), cum, tb, 16)
    for t, k in zip(out, ["isect_ids", "gaussian_ids", "isect_ids_sorted", "gaussian_ids_sorted", "tile_bins"]):
        assert np.array_equal(t.cpu().numpy(), refout[k]), ("product binning", k)
    v_sh = _C.compute_sh_backward("fast", n, 3, 3, cu(d["means"] - d["cam_pos"][None]),
                                  cu(refout["v_colors"] * (refout["colors"] > 0)))
    assert _rel(v_sh.cpu().numpy(), refout["v_sh"]) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(CASES))
def test_reference_cuda_live(tag):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref

    if not os.path.exists(build_ref.so_path()):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    ref = build_ref.load_ref()
    d = _case(tag)
    refout = _run_reference_cuda(ref, d)
    if tag == "static":  # the N-channel kernels reject blur / rolling shutter (bindings.cu:541-546)
        refout.update(_run_reference_cuda_nd(ref, d, refout))
    if os.environ.get("B200_WRITE_REFCUDA"):
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.savez_compressed(os.path.join(ROOT, "gpurun_out", f"refcuda_{tag}.npz"), **refout)
    _oracle_checks(refout, d)
    _product_checks(refout, d)
    if tag == "static":
        _oracle_nd_checks(refout, d)
        _product_nd_checks(refout, d)


def _fixture(tag):
    p = os.path.join(GOLD, f"refcuda_{tag}.npz")
    if not os.path.exists(p):
        pytest.skip(f"{p} not generated yet")
    return dict(np.load(p))


@pytest.mark.parametrize("tag", list(CASES))
def test_refcuda_fixture_vs_oracle(tag):
    """CPU: the oracle reproduces the committed reference-CUDA outputs."""
    f = _fixture(tag)
    _oracle_checks(f, _case(tag))
    if "nd_img" in f:
        _oracle_nd_checks(f, _case(tag))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(CASES))
def test_refcuda_fixture_vs_product(tag):
    f = _fixture(tag)
    _product_checks(f, _case(tag))
    if "nd_img" in f:
        _product_nd_checks(f, _case(tag))
