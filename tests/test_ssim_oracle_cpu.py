"""CPU: oracle/ssim_oracle.py (the restatement of pytorch_msssim the GPU loss tests are checked against) versus an
independent implementation of the same published definition (Wang et al. 2004 SSIM with an 11-tap sigma-1.5 Gaussian
window, K = (0.01, 0.03), valid region only) built on scipy.ndimage -- no shared code with the oracle.  pytorch_msssim
itself is not installable here (parity of the restatement against the package stays unpinned, see the oracle's header);
this at least pins the restatement to the textbook algorithm."""
import numpy as np
import pytest
import torch
from scipy.ndimage import correlate1d

from oracle import ssim_oracle as SO


def _ssim_scipy(pred, target):
    """(H, W, C) float64 arrays -> mean SSIM over channels and the valid (H-10) x (W-10) region."""
    x = np.arange(11, dtype=np.float64) - 5
    w = np.exp(-x * x / (2 * 1.5 ** 2))
    w /= w.sum()

    def filt(img):  # separable correlation, then crop to the positions whose whole window is inside the image
        out = correlate1d(correlate1d(img, w, axis=0, mode="constant"), w, axis=1, mode="constant")
        return out[5:-5, 5:-5]

    C1, C2 = 0.01 ** 2, 0.03 ** 2
    vals = []
    for c in range(pred.shape[2]):
        X, Y = target[:, :, c], pred[:, :, c]
        mu1, mu2 = filt(X), filt(Y)
        s1, s2, s12 = filt(X * X) - mu1 * mu1, filt(Y * Y) - mu2 * mu2, filt(X * Y) - mu1 * mu2
        m = ((2 * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1)) * ((2 * s12 + C2) / (s1 + s2 + C2))
        vals.append(m.mean())
    return float(np.mean(vals))


@pytest.mark.parametrize("H,W,C", [(40, 56, 3), (11, 11, 1), (64, 27, 4)])
def test_oracle_matches_an_independent_scipy_implementation(H, W, C):
    g = torch.Generator().manual_seed(H + W + C)
    t = torch.rand(H, W, C, generator=g, dtype=torch.float64)
    p = (t + 0.1 * torch.randn(H, W, C, generator=g, dtype=torch.float64)).clamp(0, 1)
    got = float(SO.ssim_hwc(p, t))
    want = _ssim_scipy(p.numpy(), t.numpy())
    # the oracle builds the window in float32 (like the package), scipy here in float64: agreement to ~1e-7
    assert abs(got - want) < 5e-7, (got, want)
    assert abs(float(SO.ssim_hwc(t, t)) - 1.0) < 1e-12
    loss = float(SO.photometric_loss(p, t, 0.2))
    assert abs(loss - (0.8 * float((t - p).abs().mean()) + 0.2 * (1 - got))) < 1e-12
