"""Stand-alone driver for the blend kernels on a BASELINE config (ncu captures / timing experiments).

    python tools/blend_probe.py [--config c2] [--reps 20] [--what fwd,bwd]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dgs-deblur_b200"))
import torch

import gsplat.cuda as _C
from gsplat import synthetic


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--what", default="fwd,bwd")
    ap.add_argument("--full", action="store_true", help="blend the reference's full tile lists instead of the culled ones")
    ap.add_argument("--counters", action="store_true",
                    help="needs a -DB200_BLEND_COUNTERS build (B200SPLAT_LIB=.../libb200splat_cnt.so): print what one forward and "
                         "one backward launch execute (entries tested, warp visits, sample blocks, evaluations, useful evaluations)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    sc = synthetic.make_scene(a.config, device=dev, n_override=a.n)
    cam = sc["cameras"][0]
    N, H, W, S = sc["N"], sc["H"], sc["W"], sc["blur_samples"] if sc["exposure_time"] > 0 else 1
    rs, ex = sc["rolling_shutter_time"], sc["exposure_time"]
    q = sc["quats"] / sc["quats"].norm(dim=-1, keepdim=True)
    cov3d, xys, depths, pix_vels, radii, conics, comp, nth = _C.project_gaussians_forward(
        N, sc["means"], sc["log_scales"].exp(), 1.0, q, None, None, rs, ex, cam["viewmat"], cam["fx"], cam["fy"], cam["cx"],
        cam["cy"], H, W, 16, 0.01, _vel_tensors=(cam["lin_vel"], cam["ang_vel"]))
    coeffs = torch.cat((sc["sh_dc"], sc["sh_rest"]), 1).contiguous()
    colors = torch.clamp(_C.compute_sh_forward("fast", N, 3, 3, (sc["means"] - cam["cam_pos"]).contiguous(), coeffs) + 0.5, min=0)
    opac = (torch.sigmoid(sc["opacity_logit"]) * comp[:, None]).contiguous()
    I, cum = _C.cumulative_intersects(nth)
    tb = ((W + 15) // 16, (H + 15) // 16, 1)
    isect, gids = _C.map_gaussian_to_intersects(N, I, xys, depths, radii, cum, tb, 16)
    isect_s, gids_s = _C.sort_intersects(tb[0] * tb[1], isect, gids)
    bins = _C.get_tile_bin_edges(I, isect_s, tb)
    bg = sc["background"]
    packed = _C.pack_records(xys, pix_vels, conics, colors, opac)
    if not a.full:
        _, gids_s, bins = _C.bin_cull(packed, depths, radii, nth, H, W, 16, S, rs, ex)
    fwd = lambda: _C.blend_forward_packed(H, W, 16, S, gids_s, bins, packed, rs, ex, bg)
    img, Ts, fi = fwd()
    v_out = torch.sign(img - cam["target"]) / img.numel()
    v_alpha = torch.zeros(H, W, device=dev)
    bwd = lambda: _C.blend_backward_packed(N, H, W, 16, S, gids_s, bins, packed, rs, ex, bg, Ts, fi, v_out, v_alpha)
    ln = (bins[:, 1] - bins[:, 0]).float()
    print(f"config {a.config}: N={N} I={I} list entries={gids_s.numel()} ({'full' if a.full else 'culled'}) visible={int((nth > 0).sum())} "
          f"tiles={bins.shape[0]} list len mean/max {ln.mean():.0f}/{ln.max():.0f}")
    if a.counters:
        import ctypes
        import json

        from gsplat import _lib
        lib = _lib.load()
        buf = (ctypes.c_ulonglong * 16)()
        lib.b200_blend_counters(None, 1)
        fwd()
        bwd()
        lib.b200_blend_counters(buf, 0)
        names = ["entries_tested_by_a_warp", "warp_visits", "sample_blocks", "pixel_sample_evaluations",
                 "evaluations_passing_the_alpha_test", "visits_that_blend_or_reduce"]
        out = {"config": a.config, "lists": "full" if a.full else "culled", "list_entries": int(gids_s.numel()),
               "forward": dict(zip(names, [int(x) for x in buf[0:6]])), "backward": dict(zip(names, [int(x) for x in buf[8:14]]))}
        for k in ("forward", "backward"):
            d = out[k]
            d["useful_fraction_of_evaluations"] = round(d["evaluations_passing_the_alpha_test"] / max(1, d["pixel_sample_evaluations"]), 4)
        print(json.dumps(out))
        return
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        if name not in a.what.split(","):
            continue
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / a.reps * 1000:.1f} us per call (memsets + kernel)")


if __name__ == "__main__":
    main()
