"""CPU executable specification of the binning's tile tests (csrc/binning.cu, csrc/blend_common.cuh), in float32 numpy:

  per_sample   -- may_touch_rect: for each blur sample, the Gaussian's centre swept over the tile's rolling-shutter window,
                  inflated by (hx, hy), against the tile's pixel-centre rectangle;
  closed_form  -- may_touch_rect_closed_form: one interval test over the sample index.

The closed form must keep every pair the per-sample test keeps (dropping a needed pair would change pixels; keeping an
extra one only costs time).  Checked here on ~2 M random and adversarial (zero / tiny / huge velocity, infinite extents,
grazing, NaN) cases; the GPU test test_closed_form_tile_test_keeps_a_superset_of_the_per_sample_test checks the kernels
themselves."""
import numpy as np
import pytest

f32 = np.float32


def per_sample(x, y, vx, vy, hx, hy, X0, X1, Y0, Y1, r0, r1, exposure, S):
    keep = np.zeros(x.shape, bool)
    with np.errstate(invalid="ignore", over="ignore"):
        for s in range(S):
            b = f32((f32(s) / f32(S - 1) - f32(0.5)) * f32(exposure)) if S > 1 else f32(0)
            t0, t1 = (b + r0).astype(f32), (b + r1).astype(f32)
            ax, bx, ay, by = t0 * vx, t1 * vx, t0 * vy, t1 * vy
            cx0, cx1 = x + np.fmin(ax, bx), x + np.fmax(ax, bx)
            cy0, cy1 = y + np.fmin(ay, by), y + np.fmax(ay, by)
            out = (cx0 - hx > X1) | (cx1 + hx < X0) | (cy0 - hy > Y1) | (cy1 + hy < Y0)
            keep |= ~out
    return keep & ~(hx < 0)


def _time_window(c, v, lo, hi):
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        a, b = (lo - c - f32(0.01)).astype(f32), (hi - c + f32(0.01)).astype(f32)
        iv = (f32(1.0) / v).astype(f32)
        p, q = (a * iv).astype(f32), (b * iv).astype(f32)
        t_lo, t_hi = np.fmin(p, q), np.fmax(p, q)
        t_lo = (t_lo - (f32(1e-5) * np.abs(t_lo) + f32(1e-9))).astype(f32)
        t_hi = (t_hi + (f32(1e-5) * np.abs(t_hi) + f32(1e-9))).astype(f32)
        zero = v == 0
        ok = ~(a > 0) & ~(b < 0)
        t_lo = np.where(zero, np.where(ok, -np.inf, np.inf), t_lo).astype(f32)
        t_hi = np.where(zero, np.where(ok, np.inf, -np.inf), t_hi).astype(f32)
    return t_lo, t_hi


def closed_form(x, y, vx, vy, hx, hy, X0, X1, Y0, Y1, r0, r1, exposure, S):
    with np.errstate(invalid="ignore", over="ignore"):
        xl, xh = _time_window(x, vx, (X0 - hx).astype(f32), (X1 + hx).astype(f32))
        yl, yh = _time_window(y, vy, (Y0 - hy).astype(f32), (Y1 + hy).astype(f32))
        L, U = (np.fmax(xl, yl) - r1).astype(f32), (np.fmin(xh, yh) - r0).astype(f32)
        if S == 1 or not exposure > 0:
            keep = ~(L > 0) & ~(U < 0)
        else:
            kps = f32(f32(S - 1) / f32(exposure))
            half = f32(0.5) * f32(S - 1)
            u = ((L * kps + half) - f32(1e-3)).astype(f32)
            w = ((U * kps + half) + f32(1e-3)).astype(f32)
            k_lo, k_hi = np.ceil(np.fmax(u, f32(0))), np.floor(np.fmin(w, f32(S - 1)))
            keep = ~(k_lo > k_hi)
    return keep & ~(hx < 0)


def _cases(rng, n, W=1920, H=1440):
    x, y = rng.uniform(-300, W + 300, n).astype(f32), rng.uniform(-300, H + 300, n).astype(f32)
    speed = 10.0 ** rng.uniform(-6, 5, n)
    ang = rng.uniform(0, 2 * np.pi, n)
    vx, vy = (speed * np.cos(ang)).astype(f32), (speed * np.sin(ang)).astype(f32)
    kind = rng.integers(0, 8, n)
    vx[kind == 0] = 0
    vy[kind == 1] = 0
    vx[kind == 2] = 0
    vy[kind == 2] = 0
    hx, hy = (10.0 ** rng.uniform(-2, 3, n)).astype(f32), (10.0 ** rng.uniform(-2, 3, n)).astype(f32)
    hx[kind == 3] = np.inf
    hy[kind == 3] = np.inf
    hx[kind == 4] = -1
    tx, ty = rng.integers(0, W // 16, n), rng.integers(0, H // 16, n)
    # half of the cases: aim the Gaussian so that it just grazes its tile (the interesting boundary)
    graze = rng.random(n) < 0.5
    x = np.where(graze, (tx * 16 + rng.uniform(-1, 17, n) + np.sign(rng.normal(size=n)) * hx * rng.uniform(0.98, 1.02, n)), x).astype(f32)
    X0, X1 = (tx * 16 + 0.5).astype(f32), (np.minimum(W, tx * 16 + 16) - 0.5).astype(f32)
    Y0, Y1 = (ty * 16 + 0.5).astype(f32), (np.minimum(H, ty * 16 + 16) - 0.5).astype(f32)
    return x, y, vx, vy, hx, hy, X0, X1, Y0, Y1


@pytest.mark.parametrize("S,exposure,rs", [(5, 1 / 60, 0.0), (5, 1 / 60, 1 / 50), (10, 1 / 60, 1 / 50), (1, 0.0, 1 / 50),
                                           (1, 0.0, 0.0), (2, 0.3, -0.02), (7, 1e-4, 0.0)])
def test_closed_form_keeps_a_superset(S, exposure, rs):
    rng = np.random.default_rng(S * 100 + int(exposure * 1e4))
    n = 300_000
    c = _cases(rng, n)
    Y0, Y1 = c[8], c[9]
    H = 1440
    ra, rb = f32(rs) * (Y0 / f32(H) - f32(0.5)), f32(rs) * (Y1 / f32(H) - f32(0.5))
    eps = f32(4e-6) * f32(abs(rs))
    r0, r1 = (np.fmin(ra, rb) - eps).astype(f32), (np.fmax(ra, rb) + eps).astype(f32)
    ref = per_sample(*c, r0, r1, exposure, S)
    fast = closed_form(*c, r0, r1, exposure, S)
    lost = ref & ~fast
    assert not lost.any(), f"{int(lost.sum())} needed pairs dropped, e.g. case {int(np.flatnonzero(lost)[0])}"
    extra = int((fast & ~ref).sum())
    assert extra <= 0.03 * max(int(ref.sum()), 1) + 50, (extra, int(ref.sum()))  # conservative, but not by much


def test_nan_inputs_keep_the_pair():
    """Overlapping in y, far away in x: dropped.  With a NaN in any x-side quantity no comparison can exclude the pair,
    so both tests keep it (the blend's exact per-pixel tests then decide, as in the reference)."""
    one = lambda v: np.array([v], f32)
    args = dict(x=one(100), y=one(8), vx=one(5), vy=one(0.5), hx=one(3), hy=one(3), X0=one(0.5), X1=one(15.5), Y0=one(0.5),
                Y1=one(15.5), r0=one(0), r1=one(0))
    assert not per_sample(**args, exposure=1 / 60, S=5)[0] and not closed_form(**args, exposure=1 / 60, S=5)[0]
    for k in ("x", "vx", "hx"):
        bad = dict(args)
        bad[k] = one(np.nan)
        assert per_sample(**bad, exposure=1 / 60, S=5)[0], k
        assert closed_form(**bad, exposure=1 / 60, S=5)[0], k


# ---- the packed record's cull data (csrc/blend_common.cuh: make_record) ---------------------------------------

def record_extents(a, b, c, opac):
    """thr = ln(255 opac); (hx, hy) = padded half extents of {sigma <= thr}; -1 = can never contribute; inf = unbounded."""
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        thr = np.log(f32(255) * opac).astype(f32)
        det = (a.astype(np.float64) * c.astype(np.float64) - b.astype(np.float64) ** 2).astype(f32)  # double, like make_record
        tm = (f32(2) * (thr * f32(1.00001) + f32(1e-4)) / det).astype(f32)
        hx = (np.sqrt(tm * c) * f32(1.00001) + f32(1e-3)).astype(f32)
        hy = (np.sqrt(tm * a) * f32(1.00001) + f32(1e-3)).astype(f32)
        never = (thr < 0) | (opac <= 0)
        bounded = (det > 0) & (a > 0) & (c > 0)
        hx = np.where(never, f32(-1), np.where(bounded, hx, f32(np.inf))).astype(f32)
        hy = np.where(never, f32(-1), np.where(bounded, hy, f32(np.inf))).astype(f32)
        thr = np.where(never, f32(-1), thr).astype(f32)
    return thr, hx, hy


def test_record_extents_never_exclude_a_contributing_pixel():
    """Whenever the reference's exact per-pixel test lets a Gaussian contribute (sigma >= 0 and alpha = min(.999, opac
    exp(-sigma)) >= 1/255, forward.cu:411-419), the offset lies inside the record's box and sigma is below the kernels'
    `thr + 1e-4` short-cut -- so neither the per-warp cull, nor the tile cull, nor the ex2 skip can drop it."""
    rng = np.random.default_rng(7)
    n = 2_000_000
    s1, s2 = 10.0 ** rng.uniform(-0.5, 2.5, n), 10.0 ** rng.uniform(-0.5, 2.5, n)  # principal std devs in pixels
    th = rng.uniform(0, np.pi, n)
    ca, sa = np.cos(th), np.sin(th)
    a = (ca * ca / s1 ** 2 + sa * sa / s2 ** 2).astype(f32)
    c = (sa * sa / s1 ** 2 + ca * ca / s2 ** 2).astype(f32)
    b = (ca * sa * (1 / s1 ** 2 - 1 / s2 ** 2)).astype(f32)
    kind = rng.integers(0, 10, n)
    b[kind == 0] *= f32(3)          # some indefinite conics (the reference blends them where sigma >= 0)
    a[kind == 1] = -a[kind == 1]
    opac = np.where(kind == 2, 10.0 ** rng.uniform(-4, -2, n), rng.uniform(0.003, 1.0, n)).astype(f32)
    thr, hx, hy = record_extents(a, b, c, opac)
    # offsets: on and around the alpha = 1/255 contour (direction random), plus uniform ones
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        phi = rng.uniform(0, 2 * np.pi, n)
        ux, uy = np.cos(phi), np.sin(phi)
        q = 0.5 * (a * ux * ux + c * uy * uy) + b * ux * uy
        rad = np.sqrt(np.maximum(thr, 0) / np.where(q > 0, q, np.nan)) * rng.uniform(0.9, 1.05, n)
        rad = np.where(np.isfinite(rad), rad, rng.uniform(0, 200, n))
        dx, dy = (rad * ux).astype(f32), (rad * uy).astype(f32)
        sigma = (f32(0.5) * (a * dx * dx + c * dy * dy) + b * dx * dy).astype(f32)
        alpha = np.minimum(f32(0.999), opac * np.exp(-sigma.astype(np.float64)))
        contributes = (sigma >= 0) & (alpha >= 1.0 / 255.0 * (1 - 1e-6))  # a hair generous: covers the ex2.approx error
    assert int(contributes.sum()) > 200_000
    inside = (np.abs(dx) <= hx) & (np.abs(dy) <= hy)
    bad = contributes & ~inside
    assert not bad.any(), f"{int(bad.sum())} contributing offsets outside the cull box, e.g. {int(np.flatnonzero(bad)[0])}"
    bad2 = contributes & (sigma > thr + f32(1e-4))
    assert not bad2.any(), f"{int(bad2.sum())} contributing offsets above the sigma short-cut"
    # and the box is tight where float32 can tell (aspect ratio <= 5, so a*c - b*b does not cancel): exact extent + 1 %
    ok = (hx > 0) & np.isfinite(hx) & (kind > 2) & (np.maximum(s1, s2) <= 5 * np.minimum(s1, s2))
    det = a.astype(np.float64) * c - b.astype(np.float64) ** 2
    with np.errstate(invalid="ignore", divide="ignore"):
        exact_hx = np.sqrt(2 * (np.maximum(thr, 0).astype(np.float64) + 2e-4) * c / det)  # (+ the 1e-4 sigma margin)
    assert int(ok.sum()) > 100_000 and np.all(hx[ok] <= exact_hx[ok] * 1.01 + 0.01)


def test_record_extents_hold_at_the_tangent_points_of_needle_splats():
    """Needle-shaped splats (eigenvalue ratio up to 1e6, any rotation): the offsets where the alpha = 1/255 contour touches
    its bounding box are the ones a too-small box would lose first.  a*c - b*b cancels there, so make_record forms it in
    double; with the float32 product the box comes out up to percents too small."""
    rng = np.random.default_rng(11)
    n = 500_000
    s1 = 10.0 ** rng.uniform(-0.5, 0.5, n)
    s2 = s1 * 10.0 ** rng.uniform(1.5, 3.0, n)
    th = np.where(rng.random(n) < 0.5, np.pi / 4, rng.uniform(0, np.pi, n))
    ca, sa = np.cos(th), np.sin(th)
    a = (ca * ca / s1 ** 2 + sa * sa / s2 ** 2).astype(f32)
    c = (sa * sa / s1 ** 2 + ca * ca / s2 ** 2).astype(f32)
    b = (ca * sa * (1 / s1 ** 2 - 1 / s2 ** 2)).astype(f32)
    opac = rng.uniform(0.01, 1.0, n).astype(f32)
    thr, hx, hy = record_extents(a, b, c, opac)
    A, B, C = a.astype(np.float64), b.astype(np.float64), c.astype(np.float64)
    det = A * C - B * B
    ok = (det > 0) & (thr > 0) & np.isfinite(hx)
    assert int(ok.sum()) > 300_000
    with np.errstate(invalid="ignore", divide="ignore"):
        t = np.maximum(thr.astype(np.float64), 0) * rng.uniform(0.97, 1.0, n)   # on / just inside the contour
        ex = np.sqrt(2 * t * C / det)                                            # extreme x of {sigma = t}
        for dx, dy in ((ex, -B / C * ex), (-B / A * np.sqrt(2 * t * A / det), np.sqrt(2 * t * A / det))):
            dx32, dy32 = dx.astype(f32), dy.astype(f32)
            sigma = (f32(0.5) * (a * dx32 * dx32 + c * dy32 * dy32) + b * dx32 * dy32).astype(f32)
            alpha = np.minimum(f32(0.999), opac * np.exp(-sigma.astype(np.float64)))
            contributes = ok & (sigma >= 0) & (alpha >= 1.0 / 255.0)
            inside = (np.abs(dx32) <= hx) & (np.abs(dy32) <= hy)
            assert int(contributes.sum()) > 50_000
            assert not (contributes & ~inside).any(), int((contributes & ~inside).sum())
