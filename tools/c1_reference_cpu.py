"""BASELINE config 1 as BASELINE.json states it: "project_gaussians + rasterize fwd on 10k random Gaussians, 256x256,
1 pose sample, CPU PyTorch reference (no GPU)" -- the reference's OWN torch implementation (gsplat/_torch_impl.py:
project_gaussians_forward :396-467, compute_sh_color, map_gaussian_to_intersects :470-503, get_tile_bin_edges :506-527,
rasterize_forward :530-597, a per-pixel Python loop), imported from /root/reference, timed on this container's cores.
Runs only where /root/reference exists (the build container); the result is committed under profiles/.

    python tools/c1_reference_cpu.py [--rows 256] > profiles/r3_c1_reference_cpu.json

--rows R: rasterize only the first R image rows and scale (the loop is ~minutes for the full image); stated in the output.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dgs-deblur_b200"))
import torch  # noqa: E402

from gsplat import synthetic  # noqa: E402  (this repo's scene generator: the same config-1 scene bench.py / the tests use)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=256)
    a = ap.parse_args()
    sc = synthetic.make_scene("c1", device="cpu")
    cam = sc["cameras"][0]
    N, H, W = sc["N"], sc["H"], sc["W"]
    sys.modules.pop("gsplat", None)
    for k in [k for k in sys.modules if k.startswith("gsplat.")]:
        del sys.modules[k]
    sys.path.insert(0, "/root/reference/gsplat")
    import gsplat._torch_impl as T  # the REFERENCE's module

    assert "/root/reference" in T.__file__
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    q = sc["quats"] / sc["quats"].norm(dim=-1, keepdim=True)
    viewmat = torch.eye(4)
    viewmat[:3, :4] = cam["viewmat"].reshape(-1)[:12].view(3, 4)
    bw = 16
    tile_bounds = ((W + bw - 1) // bw, (H + bw - 1) // bw, 1)
    out = {"config": "c1: 10k Gaussians, 256x256, 1 pose sample, forward only", "cores": cores, "torch_threads": torch.get_num_threads(),
           "implementation": "/root/reference/gsplat/gsplat/_torch_impl.py (unmodified, imported)"}
    zero3 = torch.zeros(3)
    t0 = time.perf_counter()
    with torch.no_grad():
        (cov3d, cov2d, xys, depths, pix_vel, radii, conics, comp, num_tiles_hit, mask) = T.project_gaussians_forward(
            sc["means"], sc["log_scales"].exp(), 1.0, q, zero3, zero3, 0.0, 0.0, viewmat, (cam["fx"], cam["fy"], cam["cx"], cam["cy"]),
            (W, H), bw, 0.01)
        t1 = time.perf_counter()
        coeffs = torch.cat((sc["sh_dc"], sc["sh_rest"]), 1)
        colors = torch.clamp(T.compute_sh_color(sc["means"] - cam["cam_pos"], coeffs) + 0.5, min=0.0)
        opac = torch.sigmoid(sc["opacity_logit"]) * comp[:, None]
        t2 = time.perf_counter()
        # tile lists: the reference's torch map_gaussian_to_intersects stops at the first invisible Gaussian
        # (_torch_impl.py:478-479 `break`), so it cannot bin this scene; the lists come from this repo's oracle port
        sys.path.insert(0, ROOT)
        from oracle import oracle as O
        O.build()
        b = O.bin_and_sort(xys.numpy(), depths.numpy(), radii.numpy(), num_tiles_hit.int().numpy(), H, W, bw)
        gids_s, bins = torch.from_numpy(b["gaussian_ids_sorted"]), torch.from_numpy(b["tile_bins"]).view(-1, 2)
        cum = torch.cumsum(num_tiles_hit, 0)
        t3 = time.perf_counter()
        rows = min(a.rows, H)
        img, Ts, fi = T.rasterize_forward(tile_bounds, (bw, bw, 1), (W, rows, 1), gids_s, bins, xys, conics, colors, opac[:, 0],
                                          sc["background"])
        t4 = time.perf_counter()
    blend = (t4 - t3) * H / rows
    out.update(projection_s=round(t1 - t0, 4), sh_s=round(t2 - t1, 4), binning_oracle_port_s=round(t3 - t2, 4), rasterize_s=round(blend, 3),
               rasterize_rows_timed=rows, intersections=int(cum[-1]), visible=int((num_tiles_hit > 0).sum()),
               total_s=round((t3 - t0) + blend, 3), images_per_s=round(1.0 / ((t3 - t0) + blend), 5))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
