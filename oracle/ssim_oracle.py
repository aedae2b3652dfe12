"""CPU oracle for the SSIM term of the photometric loss -- TEST INFRASTRUCTURE ONLY (imported by tests/ only).

The reference computes `1 - SSIM(gt, pred)` with the third-party package `pytorch-msssim`
(nerfstudio/pyproject.toml:67, unpinned; `SSIM(data_range=1.0, size_average=True, channel=3)`,
nerfstudio/models/splatfacto.py:32,260,958).  The package is NOT vendored under /root/reference and is not installed
in this image, so this file restates its published algorithm (pytorch_msssim/ssim.py, v1.0.0: `_fspecial_gauss_1d`,
`gaussian_filter`, `_ssim`, `ssim`) in plain torch; gradients come from autograd.  **Parity unpinned**: there is no
reference output or fixture to check this restatement against; it is anchored on the package's documented defaults
(win_size 11, win_sigma 1.5, K = (0.01, 0.03), valid padding, filtering along H then W, mean over channel and space,
`nonnegative_ssim=False`) and on the reference's call site; tests/test_ssim_oracle_cpu.py cross-checks it against an
independent scipy.ndimage implementation of the textbook definition.
"""
import torch
import torch.nn.functional as F


def gauss_window(size=11, sigma=1.5, dtype=torch.float32):
    coords = torch.arange(size, dtype=dtype) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def gaussian_filter(x, win):
    """x (N,C,H,W); separable valid-padding filtering along H, then W (dimensions shorter than the window are skipped)."""
    C = x.shape[1]
    out = x
    w = win.to(x.dtype).view(1, 1, 1, -1).repeat(C, 1, 1, 1)
    if x.shape[2] >= win.numel():
        out = F.conv2d(out, w.transpose(2, 3), stride=1, padding=0, groups=C)
    if x.shape[3] >= win.numel():
        out = F.conv2d(out, w, stride=1, padding=0, groups=C)
    return out


def ssim_hwc(pred, target, data_range=1.0, K=(0.01, 0.03)):
    """Mean SSIM of two (H, W, C) images, as SSIM(data_range, size_average=True, channel=C)(X[None], Y[None])."""
    X = target.permute(2, 0, 1)[None]
    Y = pred.permute(2, 0, 1)[None]
    win = gauss_window(dtype=torch.float32).to(X.dtype)  # the library builds the window in float32, then casts it
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = gaussian_filter(X, win), gaussian_filter(Y, win)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    sigma1_sq = gaussian_filter(X * X, win) - mu1_sq
    sigma2_sq = gaussian_filter(Y * Y, win) - mu2_sq
    sigma12 = gaussian_filter(X * Y, win) - mu1_mu2
    cs_map = (2 * sigma12 + C2) / (sigma1_sq + sigma2_sq + C2)
    ssim_map = ((2 * mu1_mu2 + C1) / (mu1_sq + mu2_sq + C1)) * cs_map
    return torch.flatten(ssim_map, 2).mean(-1).mean()


def photometric_loss(pred, target, ssim_lambda=0.2):
    """splatfacto.py:957-975 main_loss."""
    return (1 - ssim_lambda) * torch.abs(target - pred).mean() + ssim_lambda * (1 - ssim_hwc(pred, target))
