import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "3dgs-deblur_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from util_scene import scene_np, oracle_render, cu
from gsplat import project_gaussians, rasterize_gaussians, spherical_harmonics, synthetic
from gsplat.fused import render_gaussians
import gsplat.cuda as _C
d = scene_np("c2", n=40000, H=256, W=320, S=3, rs=1 / 50, exposure=1 / 60)
r = oracle_render(d)
sc = synthetic.make_scene("c2", n_override=40000, n_cameras=1)
raw = {k: sc[k].cuda() for k in ("means", "log_scales", "quats", "opacity_logit", "sh_dc", "sh_rest")}
img, alpha, info = render_gaussians(raw["means"], raw["log_scales"], raw["quats"], raw["opacity_logit"], raw["sh_dc"], raw["sh_rest"], cu(d["viewmat"]),
    cu(d["cam_pos"]), cu(d["lin_vel"]), cu(d["ang_vel"]), d["fx"], d["fy"], d["cx"], d["cy"], d["H"], d["W"], 16, cu(d["background"]),
    rolling_shutter_time=d["rs"], exposure_time=d["exposure"], blur_samples=d["S"], sh_degree_to_use=3)
xys, depths, pv, radii, conics, comp, nth, _ = project_gaussians(cu(d["means"]), cu(d["scales"]), 1.0, cu(d["quats"]), cu(d["lin_vel"]), cu(d["ang_vel"]),
    d["rs"], d["exposure"], cu(d["viewmat"]), d["fx"], d["fy"], d["cx"], d["cy"], d["H"], d["W"], 16)
kw = dict(background=cu(d["background"]), return_alpha=True, rolling_shutter_time=d["rs"], exposure_time=d["exposure"], blur_samples=d["S"])
img_c, alpha_c = rasterize_gaussians(xys, depths, pv, radii, conics, nth, cu(r["colors"]), cu(d["opacity"]) * comp[:, None], d["H"], d["W"], 16, **kw)
p = r["proj"]
img_o, alpha_o = rasterize_gaussians(cu(p["xys"]), cu(p["depths"]), cu(p["pix_vels"]), cu(p["radii"]), cu(p["conics"]), cu(p["num_tiles_hit"]), cu(r["colors"]), cu(r["opac"]), d["H"], d["W"], 16, **kw)
ra = 1 - r["final_Ts"].mean(-1)
def st(tag, a, b):
    dd = np.abs(a.detach().cpu().numpy().astype(np.float64) - b); bad = dd > 5e-5 + 1e-4 * np.abs(b)
    print(tag, "bad", round(float(bad.mean()), 5), "max", float(dd.max()))
print("background", d["background"])
st("fused alpha vs oracle", alpha, ra); st("fused img vs oracle", img, r["img"])
st("chain alpha vs oracle", alpha_c, ra); st("chain img vs oracle", img_c, r["img"])
st("blend-on-oracle-proj alpha vs oracle", alpha_o, ra); st("blend-on-oracle-proj img", img_o, r["img"])
m = (nth.cpu().numpy() > 0) & (p["num_tiles_hit"] > 0)
print("xys max abs diff", np.abs(xys.cpu().numpy()[m] - p["xys"][m]).max(), "n near plane (z<0.1):", int(((p["depths"] > 0) & (p["depths"] < 0.1)).sum()))
