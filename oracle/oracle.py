"""numpy front-end of the C oracle (oracle/splat_oracle.c).  TEST INFRASTRUCTURE ONLY.

Each function takes/returns numpy arrays with the shapes and dtypes of the
reference binding it restates (reference file:line in splat_oracle.c's header).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile liboracle.so with gcc (seconds)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "splat_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def set_threads(n=0):
    """Set (n > 0) and return the OpenMP thread count of the C oracle."""
    fn = lib().orc_set_threads
    fn.restype, fn.argtypes = C.c_int, [C.c_int]
    return int(fn(int(n)))


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def tile_bounds(H, W, bw):
    return (W + bw - 1) // bw, (H + bw - 1) // bw


def project_forward(means, scales, glob_scale, quats, lin_vel, ang_vel, rs_time, exposure, viewmat,
                    fx, fy, cx, cy, H, W, bw, clip=0.01):
    means, scales, quats = _f(means), _f(scales), _f(quats)
    n = means.shape[0]
    lv = _f(lin_vel if lin_vel is not None else np.zeros(3)).reshape(3)
    av = _f(ang_vel if ang_vel is not None else np.zeros(3)).reshape(3)
    vm = _f(np.asarray(viewmat).reshape(-1)[:12])
    out = dict(
        cov3d=np.zeros((n, 6), np.float32), xys=np.zeros((n, 2), np.float32), depths=np.zeros(n, np.float32),
        pix_vels=np.zeros((n, 2), np.float32), radii=np.zeros(n, np.int32), conics=np.zeros((n, 3), np.float32),
        compensation=np.zeros(n, np.float32), num_tiles_hit=np.zeros(n, np.int32),
    )
    lib().orc_project_forward(
        C.c_int(n), _p(means), _p(scales), C.c_float(glob_scale), _p(quats), _p(lv), _p(av),
        C.c_float(rs_time), C.c_float(exposure), _p(vm), C.c_float(fx), C.c_float(fy), C.c_float(cx), C.c_float(cy),
        C.c_int(H), C.c_int(W), C.c_int(bw), C.c_float(clip),
        _p(out["cov3d"]), _p(out["xys"]), _p(out["depths"]), _p(out["pix_vels"]), _p(out["radii"]),
        _p(out["conics"]), _p(out["compensation"]), _p(out["num_tiles_hit"]))
    return out


def project_backward(means, scales, glob_scale, quats, lin_vel, ang_vel, rs_time, exposure, viewmat, fx, fy,
                     cov3d, radii, conics, compensation, v_xy, v_depth, v_pix_vel, v_conic, v_comp):
    means, scales, quats = _f(means), _f(scales), _f(quats)
    n = means.shape[0]
    lv = _f(lin_vel if lin_vel is not None else np.zeros(3)).reshape(3)
    av = _f(ang_vel if ang_vel is not None else np.zeros(3)).reshape(3)
    vm = _f(np.asarray(viewmat).reshape(-1)[:12])
    cov3d, radii, conics, compensation = _f(cov3d), _i(radii), _f(conics), _f(compensation)
    v_xy, v_depth, v_pix_vel, v_conic, v_comp = _f(v_xy), _f(v_depth), _f(v_pix_vel), _f(v_conic), _f(v_comp)
    out = dict(v_cov2d=np.zeros((n, 3), np.float32), v_cov3d=np.zeros((n, 6), np.float32),
               v_mean3d=np.zeros((n, 3), np.float32), v_scale=np.zeros((n, 3), np.float32),
               v_quat=np.zeros((n, 4), np.float32))
    lib().orc_project_backward(
        C.c_int(n), _p(means), _p(scales), C.c_float(glob_scale), _p(quats), _p(lv), _p(av),
        C.c_float(rs_time), C.c_float(exposure), _p(vm), C.c_float(fx), C.c_float(fy),
        _p(cov3d), _p(radii), _p(conics), _p(compensation), _p(v_xy), _p(v_depth), _p(v_pix_vel), _p(v_conic),
        _p(v_comp), _p(out["v_cov2d"]), _p(out["v_cov3d"]), _p(out["v_mean3d"]), _p(out["v_scale"]), _p(out["v_quat"]))
    return out


_METHOD = {"poly": 0, "fast": 1}


def _deg_from_bases(k):
    return {1: 0, 4: 1, 9: 2, 16: 3, 25: 4}[k]


def sh_forward(method, degrees_to_use, viewdirs, coeffs):
    viewdirs, coeffs = _f(viewdirs), _f(coeffs)
    n, k, _ = coeffs.shape
    colors = np.zeros((n, 3), np.float32)
    lib().orc_sh_forward(C.c_int(_METHOD[method]), C.c_int(n), C.c_int(_deg_from_bases(k)), C.c_int(degrees_to_use),
                         _p(viewdirs), _p(coeffs), _p(colors))
    return colors


def sh_backward(method, degree, degrees_to_use, viewdirs, v_colors):
    viewdirs, v_colors = _f(viewdirs), _f(v_colors)
    n = v_colors.shape[0]
    k = (degree + 1) ** 2
    v_coeffs = np.zeros((n, k, 3), np.float32)
    lib().orc_sh_backward(C.c_int(_METHOD[method]), C.c_int(n), C.c_int(degree), C.c_int(degrees_to_use),
                          _p(viewdirs), _p(v_colors), _p(v_coeffs))
    return v_coeffs


def cov2d_bounds(cov2d):
    cov2d = _f(cov2d)
    n = cov2d.shape[0]
    conics = np.zeros((n, 3), np.float32)
    radii = np.zeros((n, 1), np.float32)
    lib().orc_cov2d_bounds(C.c_int(n), _p(cov2d), _p(conics), _p(radii))
    return conics, radii


def cumulative_intersects(num_tiles_hit):
    cum = np.cumsum(_i(num_tiles_hit), dtype=np.int32)
    return int(cum[-1]) if cum.size else 0, cum


def map_intersects(xys, depths, radii, cum_tiles_hit, tile_bounds_xy, bw):
    xys, depths, radii, cum = _f(xys), _f(depths), _i(radii), _i(cum_tiles_hit)
    n = xys.shape[0]
    m = int(cum[-1]) if n else 0
    isect = np.zeros(m, np.int64)
    gids = np.zeros(m, np.int32)
    lib().orc_map_intersects(C.c_int(n), _p(xys), _p(depths), _p(radii), _p(cum), C.c_int(tile_bounds_xy[0]),
                             C.c_int(tile_bounds_xy[1]), C.c_int(bw), _p(isect), _p(gids))
    return isect, gids


def sort_intersects(isect, gids):
    isect = np.ascontiguousarray(isect, np.int64)
    gids = _i(gids)
    ko, vo = np.zeros_like(isect), np.zeros_like(gids)
    lib().orc_sort_intersects(C.c_int(isect.shape[0]), _p(isect), _p(gids), _p(ko), _p(vo))
    return ko, vo


def tile_bin_edges(isect_sorted, num_tiles):
    isect_sorted = np.ascontiguousarray(isect_sorted, np.int64)
    bins = np.zeros((num_tiles, 2), np.int32)
    lib().orc_tile_bin_edges(C.c_int(isect_sorted.shape[0]), _p(isect_sorted), _p(bins))
    return bins


def bin_and_sort(xys, depths, radii, num_tiles_hit, H, W, bw):
    tb = tile_bounds(H, W, bw)
    m, cum = cumulative_intersects(num_tiles_hit)
    isect, gids = map_intersects(xys, depths, radii, cum, tb, bw)
    isect_s, gids_s = sort_intersects(isect, gids)
    bins = tile_bin_edges(isect_s, tb[0] * tb[1])
    return dict(num_intersects=m, cum_tiles_hit=cum, isect_ids=isect, gaussian_ids=gids,
                isect_ids_sorted=isect_s, gaussian_ids_sorted=gids_s, tile_bins=bins)


def rasterize_forward(H, W, bw, S, ids_sorted, tile_bins, xys, pix_vels, rs_time, exposure, conics, colors,
                      opacities, background, rows=None):
    ids_sorted, tile_bins = _i(ids_sorted), _i(tile_bins)
    xys, pix_vels, conics, colors = _f(xys), _f(pix_vels), _f(conics), _f(colors)
    opac, bg = _f(opacities).reshape(-1), _f(background)
    out_img = np.zeros((H, W, 3), np.float32)
    final_Ts = np.zeros((H, W, S), np.float32)
    final_idx = np.zeros((H, W, S), np.int32)
    r0, r1 = (0, H) if rows is None else rows
    lib().orc_rasterize_forward_rows(C.c_int(r0), C.c_int(r1), C.c_int(H), C.c_int(W), C.c_int(bw), C.c_int(S),
                                     _p(ids_sorted), _p(tile_bins), _p(xys), _p(pix_vels), C.c_float(rs_time),
                                     C.c_float(exposure), _p(conics), _p(colors), _p(opac), _p(bg), _p(out_img),
                                     _p(final_Ts), _p(final_idx))
    return out_img, final_Ts, final_idx


def rasterize_backward(H, W, bw, S, ids_sorted, tile_bins, xys, pix_vels, rs_time, exposure, conics, colors,
                       opacities, background, final_Ts, final_idx, v_out, v_out_alpha, rows=None):
    ids_sorted, tile_bins = _i(ids_sorted), _i(tile_bins)
    xys, pix_vels, conics, colors = _f(xys), _f(pix_vels), _f(conics), _f(colors)
    opac, bg = _f(opacities).reshape(-1), _f(background)
    final_Ts, final_idx, v_out, v_out_alpha = _f(final_Ts), _i(final_idx), _f(v_out), _f(v_out_alpha)
    n = xys.shape[0]
    out = dict(v_xy=np.zeros((n, 2), np.float32), v_xy_abs=np.zeros((n, 2), np.float32),
               v_pix_vels=np.zeros((n, 2), np.float32), v_conic=np.zeros((n, 3), np.float32),
               v_colors=np.zeros((n, 3), np.float32), v_opacity=np.zeros((n, 1), np.float32))
    r0, r1 = (0, H) if rows is None else rows
    lib().orc_rasterize_backward_rows(
        C.c_int(r0), C.c_int(r1), C.c_int(n), C.c_int(H), C.c_int(W), C.c_int(bw), C.c_int(S), _p(ids_sorted), _p(tile_bins), _p(xys),
        _p(pix_vels), C.c_float(rs_time), C.c_float(exposure), _p(conics), _p(colors), _p(opac), _p(bg),
        _p(final_Ts), _p(final_idx), _p(v_out), _p(v_out_alpha), _p(out["v_xy"]), _p(out["v_xy_abs"]),
        _p(out["v_pix_vels"]), _p(out["v_conic"]), _p(out["v_colors"]), _p(out["v_opacity"]))
    return out


def nd_rasterize_forward(H, W, bw, ids_sorted, tile_bins, xys, conics, colors, opacities, background):
    ids_sorted, tile_bins = _i(ids_sorted), _i(tile_bins)
    xys, conics, colors = _f(xys), _f(conics), _f(colors)
    opac, bg = _f(opacities).reshape(-1), _f(background)
    ch = colors.shape[1]
    out_img = np.zeros((H, W, ch), np.float32)
    final_Ts = np.zeros((H, W), np.float32)
    final_idx = np.zeros((H, W), np.int32)
    lib().orc_nd_rasterize_forward(C.c_int(H), C.c_int(W), C.c_int(bw), C.c_int(ch), _p(ids_sorted), _p(tile_bins),
                                   _p(xys), _p(conics), _p(colors), _p(opac), _p(bg), _p(out_img), _p(final_Ts),
                                   _p(final_idx))
    return out_img, final_Ts, final_idx


def nd_rasterize_backward(H, W, bw, ids_sorted, tile_bins, xys, conics, colors, opacities, background, final_Ts,
                          final_idx, v_out, v_out_alpha):
    ids_sorted, tile_bins = _i(ids_sorted), _i(tile_bins)
    xys, conics, colors = _f(xys), _f(conics), _f(colors)
    opac, bg = _f(opacities).reshape(-1), _f(background)
    final_Ts, final_idx, v_out, v_out_alpha = _f(final_Ts), _i(final_idx), _f(v_out), _f(v_out_alpha)
    n, ch = colors.shape
    out = dict(v_xy=np.zeros((n, 2), np.float32), v_xy_abs=np.zeros((n, 2), np.float32),
               v_conic=np.zeros((n, 3), np.float32), v_colors=np.zeros((n, ch), np.float32),
               v_opacity=np.zeros((n, 1), np.float32))
    lib().orc_nd_rasterize_backward(
        C.c_int(n), C.c_int(H), C.c_int(W), C.c_int(bw), C.c_int(ch), _p(ids_sorted), _p(tile_bins), _p(xys),
        _p(conics), _p(colors), _p(opac), _p(bg), _p(final_Ts), _p(final_idx), _p(v_out), _p(v_out_alpha),
        _p(out["v_xy"]), _p(out["v_xy_abs"]), _p(out["v_conic"]), _p(out["v_colors"]), _p(out["v_opacity"]))
    return out
