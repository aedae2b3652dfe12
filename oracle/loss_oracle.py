"""CPU oracle for the L1 term of the caller's loss with its ground-truth preparation -- TEST INFRASTRUCTURE ONLY (imported by
tests/ only).

Restates, in float32 numpy with one rounding per reference operation (each torch op of the reference is its own kernel):
  * `get_gt_img` (nerfstudio/models/splatfacto.py:900-910): uint8 -> float32 / 255;
  * `composite_with_background` (:912-923): `alpha * rgb + (1 - alpha) * background` for 4-channel images;
  * `gt_img.clamp(min=min_rgb_level / 255.0)` when `min_rgb_level > 0` (:952-953; the Python double is cast to float32);
  * the mask products `gt * mask`, `pred * mask` (:957-964);
  * the gamma step on the linear render, `clamp(rgb, max=1) ** (1 / gamma)` (:879-880);
  * `Ll1 = abs(gt - pred).mean()` (:966) and its gradient w.r.t. the linear render (torch's abs / mul / pow / clamp VJPs:
    sign(pred - gt) / n * mask * (1/gamma) x^(1/gamma - 1) for x <= 1, 0 above).
Pinned: tests/test_loss_oracle_cpu.py holds it to tests/golden/loss_target.npz, produced by the REFERENCE's own methods
(tests/golden/make_golden_loss.py) -- prepared target bit for bit, loss and gradient to float rounding of `pow`.
"""
import numpy as np


def prepare_target(image_u8, background=None, min_rgb_level=0.0, mask=None):
    """(H, W, 3|4) uint8 -> the float32 (H, W, 3) image the reference's L1 / SSIM terms are taken against."""
    img = image_u8.astype(np.float32) / np.float32(255.0)
    if img.shape[2] == 4:
        alpha = img[..., 3:4]
        bg = np.asarray(background, np.float32).reshape(1, 1, 3)
        gt = (alpha * img[..., :3]).astype(np.float32) + ((np.float32(1.0) - alpha).astype(np.float32) * bg).astype(np.float32)
        gt = gt.astype(np.float32)
    else:
        gt = img
    if min_rgb_level > 0:
        gt = np.maximum(gt, np.float32(min_rgb_level / 255.0))
    if mask is not None:
        gt = (gt * np.asarray(mask, np.float32).reshape(gt.shape[0], gt.shape[1], 1)).astype(np.float32)
    return gt.astype(np.float32)


def l1_loss(pred_linear, image_u8, gamma=None, background=None, min_rgb_level=0.0, mask=None):
    """-> (loss, d loss / d pred_linear, prepared target), float32."""
    gt = prepare_target(image_u8, background, min_rgb_level, mask)
    x = np.asarray(pred_linear, np.float32)
    if gamma:
        inv = np.float32(1.0 / gamma)
        c = np.minimum(x, np.float32(1.0))
        y = np.power(c, inv, dtype=np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            dy = np.where(x <= 1.0, inv * np.power(c, inv - np.float32(1.0), dtype=np.float32), np.float32(0.0)).astype(np.float32)
    else:
        y, dy = x, np.ones_like(x)
    if mask is not None:
        m = np.asarray(mask, np.float32).reshape(x.shape[0], x.shape[1], 1)
        y = (y * m).astype(np.float32)
        dy = (dy * m).astype(np.float32)
    d = y - gt
    loss = np.float32(np.abs(d).mean(dtype=np.float64))
    grad = (np.sign(d) / np.float32(d.size) * dy).astype(np.float32)
    return loss, grad, gt
