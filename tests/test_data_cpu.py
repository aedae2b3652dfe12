"""CPU: gsplat.data.load_transforms against the reference's own dataparser (fixtures from tests/golden/make_golden_data.py:
nerfstudio_dataparser.py run on fabricated datasets), and the camera conversion of Splatfacto.get_outputs."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3dgs-deblur_b200"))
GOLD = os.path.join(ROOT, "tests", "golden")


def _materialise(tmp_path, meta):
    """Write the fabricated dataset the fixture was parsed from (blank images: only their size is ever read)."""
    from PIL import Image
    for sub in ("images", "images_2"):
        os.makedirs(tmp_path / sub)
        for fr in meta["frames"]:
            Image.fromarray(np.zeros((240, 320, 3), np.uint8)).save(tmp_path / sub / os.path.basename(fr["file_path"]))
    json.dump(meta, open(tmp_path / "transforms.json", "w"))


@pytest.mark.parametrize("case", ["data_case1", "data_case2", "data_case3"])
def test_load_transforms_matches_reference_dataparser(case, tmp_path):
    from gsplat.data import load_transforms
    spec = json.load(open(os.path.join(GOLD, case + ".json")))
    gold = np.load(os.path.join(GOLD, case + ".npz"))
    _materialise(tmp_path, spec["meta"])
    for split in ("train", "val"):
        out = load_transforms(str(tmp_path), split=split, **spec["config"])
        pre = split + "_"
        assert [os.path.relpath(p, str(tmp_path)) for p in out["image_filenames"]] == list(gold[pre + "image_filenames"])
        np.testing.assert_allclose(out["camera_to_worlds"], gold[pre + "camera_to_worlds"], rtol=0, atol=2e-6)
        for k in ("fx", "fy", "cx", "cy"):
            np.testing.assert_allclose(out[k], gold[pre + k], rtol=1e-6)
        assert np.array_equal(out["height"], gold[pre + "height"]) and np.array_equal(out["width"], gold[pre + "width"])
        if pre + "velocities" in gold:
            np.testing.assert_allclose(out["velocities"], gold[pre + "velocities"], rtol=2e-6, atol=1e-7)
            assert out["exposure_time"] == pytest.approx(float(gold[pre + "exposure_time"]))
            assert out["rolling_shutter_time"] == pytest.approx(float(gold[pre + "rolling_shutter_time"]))
        else:
            assert out["velocities"] is None and out["exposure_time"] is None
        assert out["dataparser_scale"] == pytest.approx(float(gold[pre + "dataparser_scale"]), rel=1e-6)
        np.testing.assert_allclose(out["dataparser_transform"], gold[pre + "dataparser_transform"], rtol=0, atol=2e-6)


def test_load_transforms_errors(tmp_path):
    from gsplat.data import load_transforms
    spec = json.load(open(os.path.join(GOLD, "data_case1.json")))
    meta = spec["meta"]
    _materialise(tmp_path, meta)
    with pytest.raises(ValueError):
        load_transforms(str(tmp_path), split="nope", downscale_factor=1)
    with pytest.raises(ValueError):
        load_transforms(str(tmp_path), orientation_method="pca")
    del meta["frames"][3]["camera_angular_velocity"]
    json.dump(meta, open(tmp_path / "transforms.json", "w"))
    with pytest.raises(AssertionError):
        load_transforms(str(tmp_path))


def test_to_gsplat_camera_is_the_splatfacto_camera_block():
    """splatfacto.py:733-747,799-800: flip y/z of the camera axes, invert analytically, rotate the velocities by the flip."""
    from gsplat.data import to_gsplat_camera
    g = torch.Generator().manual_seed(0)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    if torch.det(q) < 0:
        q[:, 2] = -q[:, 2]
    t = torch.randn(3, generator=g)
    c2w = torch.cat([q, t[:, None]], dim=1)
    vel = torch.randn(6, generator=g)
    cam = to_gsplat_camera(c2w, vel)
    R_edit = torch.diag(torch.tensor([1.0, -1.0, -1.0]))
    c2w_gs = torch.eye(4)
    c2w_gs[:3, :3] = q @ R_edit
    c2w_gs[:3, 3] = t
    torch.testing.assert_close(cam["viewmat"], torch.linalg.inv(c2w_gs), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(cam["viewmat"][:3, :3] @ t + cam["viewmat"][:3, 3], torch.zeros(3), rtol=0, atol=1e-6)
    torch.testing.assert_close(cam["lin_vel"], R_edit @ vel[:3])
    torch.testing.assert_close(cam["ang_vel"], R_edit @ vel[3:])
    torch.testing.assert_close(cam["cam_pos"], t)
    with pytest.raises(RuntimeError):
        from gsplat.data import ImagePrefetcher
        ImagePrefetcher([torch.zeros(4, 4, 3, dtype=torch.uint8)], [torch.zeros(21)], "cpu")


def test_load_ply_points_ascii_and_binary(tmp_path):
    """sparse_pc.ply seed cloud: both encodings, then [xyz 1] @ T^T * scale (nerfstudio_dataparser.py:469-491)."""
    from gsplat.data import load_ply_points
    rng = np.random.default_rng(5)
    xyz = rng.normal(size=(37, 3)).astype(np.float32)
    rgb = rng.integers(0, 256, (37, 3)).astype(np.uint8)
    header = "ply\nformat {fmt} 1.0\ncomment test\nelement vertex 37\nproperty float x\nproperty float y\nproperty float z\n" \
             "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face 0\nproperty list uchar int vertex_indices\nend_header\n"
    a = tmp_path / "a.ply"
    with open(a, "w") as f:
        f.write(header.format(fmt="ascii"))
        for p_, c_ in zip(xyz, rgb):
            f.write("%r %r %r %d %d %d\n" % (float(p_[0]), float(p_[1]), float(p_[2]), c_[0], c_[1], c_[2]))
    b = tmp_path / "b.ply"
    rec = np.zeros(37, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")])
    rec["x"], rec["y"], rec["z"] = xyz.T
    rec["red"], rec["green"], rec["blue"] = rgb.T
    with open(b, "wb") as f:
        f.write(header.format(fmt="binary_little_endian").encode())
        f.write(rec.tobytes())
    T = np.concatenate([np.linalg.qr(rng.normal(size=(3, 3)))[0], rng.normal(size=(3, 1))], axis=1).astype(np.float32)
    want = (np.concatenate([xyz, np.ones((37, 1), np.float32)], 1) @ T.T) * 0.37
    for path in (a, b):
        got_xyz, got_rgb = load_ply_points(str(path), T, 0.37)
        np.testing.assert_allclose(got_xyz, want, rtol=1e-5, atol=1e-6)
        assert np.array_equal(got_rgb, rgb)
    raw_xyz, _ = load_ply_points(str(b))
    np.testing.assert_array_equal(raw_xyz, xyz)
    bad = tmp_path / "bad.ply"
    bad.write_text("plx\n")
    with pytest.raises(ValueError):
        load_ply_points(str(bad))


def test_load_image_u8_and_compositing(tmp_path):
    """Greyscale -> 3 channels, RGBA kept or composited over alpha_color (base_dataset.py:61-110), per-step compositing
    against a background (splatfacto.py:912-923)."""
    from PIL import Image
    from gsplat.data import composite_u8, load_image_u8
    rng = np.random.default_rng(1)
    rgba = rng.integers(0, 256, (6, 5, 4)).astype(np.uint8)
    Image.fromarray(rgba, "RGBA").save(tmp_path / "a.png")
    Image.fromarray(rgba[:, :, 0], "L").save(tmp_path / "g.png")
    Image.fromarray(rgba[:, :, :3], "RGB").save(tmp_path / "c.png")
    keep = load_image_u8(str(tmp_path / "a.png"))
    assert keep.dtype == torch.uint8 and np.array_equal(keep.numpy(), rgba)
    assert np.array_equal(load_image_u8(str(tmp_path / "c.png")).numpy(), rgba[:, :, :3])
    grey = load_image_u8(str(tmp_path / "g.png")).numpy()
    assert grey.shape == (6, 5, 3) and np.array_equal(grey[:, :, 1], rgba[:, :, 0])
    col = torch.tensor([0.2, 0.5, 1.0])
    flat = load_image_u8(str(tmp_path / "a.png"), alpha_color=col)
    f = torch.from_numpy(rgba)  # the reference's uint8 formula, base_dataset.py:102-110
    want = torch.clamp(f[:, :, :3] * (f[:, :, -1:] / 255.0) + 255.0 * col * (1.0 - f[:, :, -1:] / 255.0), min=0, max=255).to(torch.uint8)
    assert torch.equal(flat, want)
    bg = torch.tensor([0.1, 0.9, 0.3])
    tgt = composite_u8(keep, bg)
    a = f[..., -1:].float() / 255
    torch.testing.assert_close(tgt, a * (f[..., :3].float() / 255) + (1 - a) * bg)
    assert torch.equal(composite_u8(torch.from_numpy(rgba[:, :, :3].copy()), bg), torch.from_numpy(rgba[:, :, :3].copy()).float() / 255)
    with pytest.raises(ValueError):
        load_image_u8(str(tmp_path / "a.png"), alpha_color=[0.0, 2.0, 0.0])
