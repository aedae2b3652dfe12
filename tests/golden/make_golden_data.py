"""Generates tests/golden/data_case*.{json,npz}: fabricated nerfstudio datasets (transforms.json with this fork's
exposure_time / rolling_shutter_time / per-frame camera velocities) parsed by the REFERENCE's own dataparser
(nerfstudio/data/dataparsers/nerfstudio_dataparser.py), imported from /root/reference in the build container.
`viser` (viewer dependency, absent here) is stubbed: the parser only imports it for an unrelated box class.

    python tests/golden/make_golden_data.py        # needs /root/reference; the fixtures it writes are committed
"""
import json
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/nerfstudio")
for m in ("viser", "viser.transforms"):
    sys.modules[m] = types.ModuleType(m)
sys.modules["viser"].transforms = sys.modules["viser.transforms"]

from PIL import Image  # noqa: E402
from nerfstudio.data.dataparsers.nerfstudio_dataparser import NerfstudioDataParserConfig  # noqa: E402


def random_pose(rng):
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    # bias the camera y axes upwards so that the "up" orientation is well defined
    R[:, 1] = 0.6 * R[:, 1] + 0.4 * np.array([0.1, 0.9, 0.3])
    Rq, _ = np.linalg.qr(R)
    if np.linalg.det(Rq) < 0:
        Rq[:, 2] = -Rq[:, 2]
    T = np.eye(4)
    T[:3, :3] = Rq
    T[:3, 3] = rng.uniform(-3, 3, 3) + np.array([5.0, -2.0, 1.0])
    return T


def make_case(case, rng):
    n = case["n"]
    frames = []
    names = [f"images/frame_{i:05d}.png" for i in rng.permutation(n)]  # stored out of order: the parser sorts by name
    for name in names:
        fr = dict(file_path=name, transform_matrix=random_pose(rng).tolist())
        if case["velocities"]:
            fr["camera_linear_velocity"] = rng.normal(0, 0.3, 3).tolist()
            fr["camera_angular_velocity"] = rng.normal(0, 0.7, 3).tolist()
        if case["per_frame_intrinsics"]:
            fr.update(fl_x=float(rng.uniform(300, 500)), fl_y=float(rng.uniform(300, 500)), cx=float(rng.uniform(150, 170)),
                      cy=float(rng.uniform(110, 130)), w=320, h=240)
        frames.append(fr)
    meta = dict(frames=frames)
    if not case["per_frame_intrinsics"]:
        meta.update(fl_x=410.5, fl_y=409.25, cx=161.0, cy=119.5, w=320, h=240)
    if case["velocities"]:
        meta.update(exposure_time=1 / 60, rolling_shutter_time=1 / 50)
    meta.update(case.get("extra_meta", {}))
    return meta


CASES = [
    dict(name="data_case1", n=23, velocities=True, per_frame_intrinsics=False, config={}),
    dict(name="data_case2", n=17, velocities=False, per_frame_intrinsics=True,
         config=dict(orientation_method="none", center_method="none", auto_scale_poses=False, eval_mode="interval",
                     eval_interval=4, scale_factor=0.5, downscale_factor=2)),
    dict(name="data_case3", n=9, velocities=True, per_frame_intrinsics=False, config=dict(eval_mode="all"),
         extra_meta=dict(applied_transform=[[0, 1, 0, 0.5], [1, 0, 0, -1.0], [0, 0, -1, 2.0]], applied_scale=0.25)),
]


def main():
    rng = np.random.default_rng(2024)
    for case in CASES:
        meta = make_case(case, rng)
        with tempfile.TemporaryDirectory() as tmp:
            for sub in ("images", "images_2"):
                os.makedirs(os.path.join(tmp, sub))
                for fr in meta["frames"]:
                    Image.fromarray(np.zeros((240, 320, 3), np.uint8)).save(os.path.join(tmp, sub, os.path.basename(fr["file_path"])))
            json.dump(meta, open(os.path.join(tmp, "transforms.json"), "w"))
            out = {}
            for split in ("train", "val"):
                parser = NerfstudioDataParserConfig(data=Path(tmp), **case["config"]).setup()
                o = parser.get_dataparser_outputs(split=split)
                cam = o.cameras
                pre = split + "_"
                out[pre + "image_filenames"] = np.array([os.path.relpath(str(p), tmp) for p in o.image_filenames])
                out[pre + "camera_to_worlds"] = cam.camera_to_worlds.numpy()
                for k in ("fx", "fy", "cx", "cy", "height", "width"):
                    out[pre + k] = getattr(cam, k).numpy().reshape(-1)
                if cam.velocities is not None:
                    out[pre + "velocities"] = cam.velocities.numpy()
                out[pre + "dataparser_scale"] = np.float64(o.dataparser_scale)
                out[pre + "dataparser_transform"] = o.dataparser_transform.numpy()
                md = cam.metadata or {}
                out[pre + "exposure_time"] = np.float64(md.get("exposure_time", np.nan))
                out[pre + "rolling_shutter_time"] = np.float64(md.get("rolling_shutter_time", np.nan))
        json.dump(dict(meta=meta, config=case["config"]), open(os.path.join(HERE, case["name"] + ".json"), "w"))
        np.savez_compressed(os.path.join(HERE, case["name"] + ".npz"), **out)
        print(case["name"], {k: v.shape for k, v in out.items() if k.startswith("train_")})


if __name__ == "__main__":
    main()
