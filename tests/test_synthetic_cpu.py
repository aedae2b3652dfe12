"""CPU: the seeded scene generator (SURVEY 8d stand-ins) is deterministic and has the documented structure."""
import torch

from gsplat import synthetic


def test_scene_is_seed_deterministic_and_device_independent_layout():
    a = synthetic.make_scene("c1")
    b = synthetic.make_scene("c1")
    for k in ("means", "log_scales", "quats", "opacity_logit", "sh_dc", "sh_rest", "background"):
        assert torch.equal(a[k], b[k]), k
    assert a["means"].shape == (10_000, 3) and a["sh_rest"].shape == (10_000, 15, 3) and a["sh_dc"].shape == (10_000, 1, 3)
    assert a["H"] == 256 and a["W"] == 256 and a["blur_samples"] == 1
    # free space around the cameras, box [-4, 4]^3
    assert float(a["means"].abs().max()) <= 4.0 and float(a["means"].abs().max(dim=1).values.min()) >= synthetic.FREE_SPACE
    cam = a["cameras"][0]
    R = cam["viewmat"][:, :3]
    assert torch.allclose(R @ R.T, torch.eye(3), atol=1e-5)  # world-to-camera rotation is orthonormal
    assert torch.allclose(-R.T @ cam["viewmat"][:, 3], cam["cam_pos"], atol=1e-5)


def test_config_table_matches_baseline_json():
    c = synthetic.CONFIGS
    assert c["c2"][1:5] == (300_000, 800, 800, 5) and c["c4"][1:4] == (1_500_000, 1440, 1920) and c["c5"][1] == 2_000_000
    assert c["c3_rs"][4] == 1 and c["c3_rs10"][4] == 10
    other = synthetic.make_scene("c2", n_override=100, seed_offset=1, n_cameras=3)
    assert len(other["cameras"]) == 3 and other["cameras"][0]["target"].shape == (800, 800, 3)
