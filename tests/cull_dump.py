"""Helper for test_gpu_parity.py: build the culled tile lists of one synthetic scene in THIS process (so that the
environment switches of libb200splat, which are read once per process, can be varied) and save them as .npz."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "3dgs-deblur_b200"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def culled_lists(name, n, motion, S, rs, exposure, H, W):
    import gsplat.cuda as _C
    from util_scene import cu, oracle_colors, oracle_project, scene_np

    d = scene_np(name, n=n, motion=motion, S=S, rs=rs, exposure=exposure, H=H, W=W)
    o = oracle_project(d)
    col = cu(oracle_colors(d))
    opac = cu((d["opacity"][:, 0] * o["compensation"])[:, None].astype(np.float32))
    xys, depths, pv, radii, conics, nth = (cu(o[k]) for k in ("xys", "depths", "pix_vels", "radii", "conics", "num_tiles_hit"))
    packed = _C.pack_records(xys, pv, conics, col, opac)
    total, ids, bins = _C.bin_cull(packed, depths, radii, nth, d["H"], d["W"], 16, d["S"], d["rs"], d["exposure"])
    return total, ids.cpu().numpy(), bins.cpu().numpy()


if __name__ == "__main__":
    out = sys.argv[1]
    name, n, motion, S, rs, exposure, H, W = eval(sys.argv[2])  # a literal tuple written by the test
    total, ids, bins = culled_lists(name, n, motion, S, rs, exposure, H, W)
    np.savez(out, total=total, ids=ids, bins=bins)
