"""Shared builders for parity tests: seeded scenes -> numpy inputs -> oracle outputs."""
import numpy as np
import torch

from gsplat import synthetic
from oracle import oracle as O


def scene_np(name="c2", n=None, cam=0, seed_offset=0, motion=True, H=None, W=None, S=None, rs=None, exposure=None):
    sc = synthetic.make_scene(name, n_override=n, seed_offset=seed_offset, n_cameras=cam + 1)
    c = sc["cameras"][cam]
    q = sc["quats"] / sc["quats"].norm(dim=-1, keepdim=True)
    H = sc["H"] if H is None else H
    W = sc["W"] if W is None else W
    d = dict(
        N=sc["N"], H=H, W=W, bw=16, S=sc["blur_samples"] if S is None else S,
        rs=sc["rolling_shutter_time"] if rs is None else rs,
        exposure=sc["exposure_time"] if exposure is None else exposure,
        means=sc["means"].numpy(), scales=sc["log_scales"].exp().numpy(), quats=q.numpy(),
        opacity=torch.sigmoid(sc["opacity_logit"]).numpy(),
        sh=torch.cat([sc["sh_dc"], sc["sh_rest"]], 1).numpy(),
        viewmat=c["viewmat"].numpy(), cam_pos=c["cam_pos"].numpy(),
        fx=W / 2.0, fy=W / 2.0, cx=W / 2.0, cy=H / 2.0,
        lin_vel=(c["lin_vel"].numpy() if motion else np.zeros(3, np.float32)),
        ang_vel=(c["ang_vel"].numpy() if motion else np.zeros(3, np.float32)),
        background=sc["background"].numpy(),
    )
    return d


def oracle_project(d):
    return O.project_forward(d["means"], d["scales"], 1.0, d["quats"], d["lin_vel"], d["ang_vel"], d["rs"],
                             d["exposure"], d["viewmat"], d["fx"], d["fy"], d["cx"], d["cy"], d["H"], d["W"], d["bw"])


def oracle_colors(d, deg=3):
    dirs = d["means"] - d["cam_pos"][None]
    col = O.sh_forward("fast", deg, dirs, d["sh"])
    return np.maximum(col + 0.5, 0.0).astype(np.float32)


def oracle_render(d, proj=None, colors=None):
    """Full oracle forward: projection -> binning -> blend.  Returns dict with everything."""
    proj = oracle_project(d) if proj is None else proj
    colors = oracle_colors(d) if colors is None else colors
    opac = (d["opacity"][:, 0] * proj["compensation"]).astype(np.float32)[:, None]
    b = O.bin_and_sort(proj["xys"], proj["depths"], proj["radii"], proj["num_tiles_hit"], d["H"], d["W"], d["bw"])
    img, Ts, fi = O.rasterize_forward(d["H"], d["W"], d["bw"], d["S"], b["gaussian_ids_sorted"], b["tile_bins"],
                                      proj["xys"], proj["pix_vels"], d["rs"], d["exposure"], proj["conics"], colors,
                                      opac, d["background"])
    return dict(proj=proj, colors=colors, opac=opac, bins=b, img=img, final_Ts=Ts, final_idx=fi)


def cu(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def frac_mismatch(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float((a != b).mean())
