"""Generates tests/golden/loss_target.npz: the caller's ground-truth preparation and L1 term as the REFERENCE computes them.

`SplatfactoModel.get_gt_img` and `.composite_with_background` (nerfstudio/models/splatfacto.py:900-923) are the
reference's own methods, imported from /root/reference and called on a stand-in `self` (they read `self.device` and
`self._downscale_if_required` only); the three lines of `get_loss_dict` between them and the loss value -- the
`min_rgb_level` clamp (:952-953), the mask products (:957-964) and the L1 mean (:966) -- are statements inside a method
that also needs pytorch_msssim, so they are executed here verbatim on the methods' outputs.  The gamma step is
`torch.clamp(rgb, max=1.0) ** (1.0 / gamma)` (:879-880).  viser / torchmetrics / pytorch_msssim / nerfacc (absent here,
none on this path) are stubbed as in tests/test_splatfacto_caller_cpu.py.

    python tests/golden/make_golden_loss.py        # needs /root/reference; the fixture it writes is committed
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "3dgs-deblur_b200"))
sys.path.insert(0, "/root/reference/nerfstudio")
from test_splatfacto_caller_cpu import _StubFinder  # noqa: E402

sys.meta_path.insert(0, _StubFinder())
import nerfstudio.models.splatfacto as sf  # noqa: E402


def main():
    g = torch.Generator().manual_seed(7)
    H, W = 23, 31  # odd sizes: the kernel's 4-pixel vector path and its tail both run
    me = types.SimpleNamespace(device=torch.device("cpu"), _downscale_if_required=lambda im: im)
    out = {}
    for tag, ch in (("rgb", 3), ("rgba", 4)):
        img = torch.randint(0, 256, (H, W, ch), generator=g, dtype=torch.uint8)
        if ch == 4:  # fully transparent / opaque pixels and everything between
            img[:5, :, 3] = 0
            img[5:10, :, 3] = 255
        background = torch.rand(3, generator=g)
        linear = torch.rand(H, W, 3, generator=g) * 1.2  # some values above the clamp of the gamma step
        mask = (torch.rand(H, W, 1, generator=g) > 0.3)
        gt = sf.SplatfactoModel.composite_with_background(me, sf.SplatfactoModel.get_gt_img(me, img), background)
        out[f"{tag}_image"] = img.numpy()
        out[f"{tag}_background"] = background.numpy()
        out[f"{tag}_linear"] = linear.numpy()
        out[f"{tag}_mask"] = mask.numpy()
        out[f"{tag}_gt"] = gt.numpy()
        for level in (0.0, 12.0):
            for use_mask in (False, True):
                for gamma in (None, 2.2):
                    pred = linear.clone().requires_grad_(True)
                    pred_img = torch.clamp(pred, max=1.0) ** (1.0 / gamma) if gamma else pred  # :879-880
                    gt_img = gt
                    if level > 0:
                        gt_img = gt_img.clamp(min=level / 255.0)  # :952-953
                    if use_mask:
                        m = mask
                        gt_img = gt_img * m  # :963
                        pred_img = pred_img * m  # :964
                    Ll1 = torch.abs(gt_img - pred_img).mean()  # :966
                    Ll1.backward()
                    key = f"{tag}_l{int(level)}_m{int(use_mask)}_g{0 if gamma is None else 1}"
                    out[key + "_loss"] = np.float32(Ll1.item())
                    out[key + "_grad"] = pred.grad.numpy()
                    out[key + "_target"] = gt_img.numpy()
    np.savez_compressed(os.path.join(HERE, "loss_target.npz"), **out)
    print("wrote", os.path.join(HERE, "loss_target.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
