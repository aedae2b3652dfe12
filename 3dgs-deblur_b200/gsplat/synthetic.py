"""Seeded synthetic stand-ins for the BASELINE.json configs.

No dataset is reachable offline, so every benchmark / parity scene is generated:
room box [-4,4]^3, means ~ U(box minus the free space |p|_inf < 1.5 the cameras move in --
without it a few hundred screen-filling "floaters" saturate every pixel after ~30 list entries
and the blend degenerates), log-scales ~ N(ln 0.02, 0.5^2), quats = normalised
N(0,1)^4, opacity logit ~ N(0, 2^2), SH dc ~ U(-1,1), rest ~ N(0, 0.1^2) (K=16,
degrees_to_use=3), camera position ~ U([-1,1]^3) with a random orientation, 90 deg hfov
(fx = fy = W/2), block_width 16, background ~ U(0,1)^3, a fixed random target image.
Velocities: omega ~ N(0, 0.7^2) rad/s, v_lin ~ N(0, 0.3^2) units/s (camera frame).
Everything is drawn on the CPU from torch.Generator(1000 + config index) and then
moved, so a scene is identical on every device.
"""
import math

import torch

CONFIGS = {
    # name: (index, N, H, W, blur_samples, exposure, rolling_shutter_time)
    "c1": (1, 10_000, 256, 256, 1, 0.0, 0.0),
    "c2": (2, 300_000, 800, 800, 5, 1.0 / 60.0, 0.0),
    "c3_rs": (3, 500_000, 720, 1280, 1, 0.0, 1.0 / 50.0),
    "c3_rs10": (3, 500_000, 720, 1280, 10, 1.0 / 60.0, 1.0 / 50.0),
    "c4": (4, 1_500_000, 1440, 1920, 5, 1.0 / 60.0, 1.0 / 50.0),
    "c5": (5, 2_000_000, 800, 800, 5, 1.0 / 60.0, 0.0),
}


FREE_SPACE = 1.5  # half-width of the empty cube around the origin (cameras live in [-1,1]^3)


def _rand_rotation(gen):
    q = torch.randn(4, generator=gen)
    q = q / q.norm()
    w, x, y, z = q.tolist()
    return torch.tensor(
        [
            [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
            [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
            [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
        ],
        dtype=torch.float32,
    )


def make_camera(gen, H, W):
    """World-to-camera (3,4) row-major view matrix + pinhole intrinsics (gsplat axes)."""
    pos = torch.rand(3, generator=gen) * 2 - 1
    R_c2w = _rand_rotation(gen)
    R = R_c2w.T.contiguous()
    t = -R @ pos
    viewmat = torch.cat([R, t[:, None]], dim=1).contiguous()
    return dict(viewmat=viewmat, fx=W / 2.0, fy=W / 2.0, cx=W / 2.0, cy=H / 2.0, cam_pos=pos)


def make_scene(name="c2", device="cpu", n_override=None, seed_offset=0, n_cameras=1, sh_k=16):
    idx, N, H, W, S, exposure, rs = CONFIGS[name]
    if n_override is not None:
        N = int(n_override)
    gen = torch.Generator().manual_seed(1000 + idx + seed_offset)
    cand = (torch.rand(2 * N + 64, 3, generator=gen) * 2 - 1) * 4.0
    means = cand[cand.abs().max(dim=1).values >= FREE_SPACE][:N].contiguous()
    assert means.shape[0] == N
    log_scales = math.log(0.02) + 0.5 * torch.randn(N, 3, generator=gen)
    quats = torch.randn(N, 4, generator=gen)
    opacity_logit = 2.0 * torch.randn(N, 1, generator=gen)
    sh_dc = torch.rand(N, 1, 3, generator=gen) * 2 - 1
    sh_rest = 0.1 * torch.randn(N, sh_k - 1, 3, generator=gen)
    background = torch.rand(3, generator=gen)
    cams = []
    for _ in range(n_cameras):
        cam = make_camera(gen, H, W)
        cam["lin_vel"] = 0.3 * torch.randn(3, generator=gen)
        cam["ang_vel"] = 0.7 * torch.randn(3, generator=gen)
        cam["target"] = torch.rand(H, W, 3, generator=gen)
        cams.append(cam)
    scene = dict(
        name=name, N=N, H=H, W=W, blur_samples=S, exposure_time=exposure, rolling_shutter_time=rs,
        block_width=16, clip_thresh=0.01, sh_degree=3 if sh_k == 16 else {1: 0, 4: 1, 9: 2, 25: 4}[sh_k],
        means=means, log_scales=log_scales, quats=quats, opacity_logit=opacity_logit,
        sh_dc=sh_dc, sh_rest=sh_rest, background=background, cameras=cams,
    )
    return to_device(scene, device)


def to_device(obj, device):
    if torch.is_tensor(obj):
        return obj.to(device)
    if isinstance(obj, dict):
        return {k: to_device(v, device) for k, v in obj.items()}
    if isinstance(obj, list):
        return [to_device(v, device) for v in obj]
    return obj
