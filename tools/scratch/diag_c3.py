import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "3dgs-deblur_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
from util_scene import scene_np, oracle_render, cu
from gsplat import project_gaussians, rasterize_gaussians
import gsplat, gsplat.cuda as _C
name = sys.argv[1]
d = scene_np(name)
r = oracle_render(d)
xys, depths, pv, radii, conics, comp, nth, _ = project_gaussians(cu(d["means"]), cu(d["scales"]), 1.0, cu(d["quats"]), cu(d["lin_vel"]), cu(d["ang_vel"]),
    d["rs"], d["exposure"], cu(d["viewmat"]), d["fx"], d["fy"], d["cx"], d["cy"], d["H"], d["W"], 16)
col = cu(r["colors"]); opac = cu(d["opacity"]) * comp[:, None]; bg = cu(d["background"])
kw = dict(background=bg, return_alpha=True, rolling_shutter_time=d["rs"], exposure_time=d["exposure"], blur_samples=d["S"])
img, alpha = rasterize_gaussians(xys, depths, pv, radii, conics, nth, col, opac, d["H"], d["W"], 16, **kw)
# full lists through the same kernels
H, W = d["H"], d["W"]; tb = ((W + 15) // 16, (H + 15) // 16, 1)
m, cum = gsplat.compute_cumulative_intersects(nth)
ids_full, bins_full = _C.bin_tiles(m, xys, depths, radii, nth, tb, 16)
img_f, Ts_f, fi_f = _C.rasterize_forward(tb, (16, 16, 1), (W, H, 1), d["S"], ids_full, bins_full, xys, pv, d["rs"], d["exposure"], conics, col, opac, bg)
# GPU blend on the ORACLE's projection + lists
b = r["bins"]
img_o, Ts_o, fi_o = _C.rasterize_forward(tb, (16, 16, 1), (W, H, 1), d["S"], cu(b["gaussian_ids_sorted"]), cu(b["tile_bins"]), cu(r["proj"]["xys"]), cu(r["proj"]["pix_vels"]),
                                         d["rs"], d["exposure"], cu(r["proj"]["conics"]), col, cu(r["opac"]), bg)
ref = r["img"]
def stat(tag, x):
    x = x.detach().cpu().numpy().astype(np.float64); dd = np.abs(x - ref)
    bad = dd > 5e-5 + 1e-4 * np.abs(ref)
    rows = bad.any(-1).mean(1)
    print(tag, "bad frac", bad.mean(), "max", dd.max(), "mean signed", (x - ref).mean(), "bad rows top/mid/bottom", rows[:H//4].mean(), rows[H//4:3*H//4].mean(), rows[3*H//4:].mean())
stat("culled ops vs oracle", img)
stat("full lists (gpu proj) vs oracle", img_f)
stat("gpu blend on oracle proj+lists vs oracle", img_o)
print("culled == full:", bool(torch.equal(img, img_f)), float((img - img_f).abs().max()))
for k, g in (("xys", xys), ("pix_vels", pv), ("conics", conics), ("depths", depths)):
    o = r["proj"][k]; gg = g.cpu().numpy(); mk = (nth.cpu().numpy() > 0) & (r["proj"]["num_tiles_hit"] > 0)
    print(k, "max abs", np.abs(gg[mk] - o[mk]).max(), "max rel", (np.abs(gg[mk] - o[mk]) / (np.abs(o[mk]) + 1e-6)).max())
print("radii mismatch", (radii.cpu().numpy() != r["proj"]["radii"]).mean(), "nth mismatch", (nth.cpu().numpy() != r["proj"]["num_tiles_hit"]).mean())
