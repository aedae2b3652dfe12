"""CPU: oracle/loss_oracle.py (restatement of the caller's target preparation + L1 term) against
tests/golden/loss_target.npz, which tests/golden/make_golden_loss.py generated with the REFERENCE model's own
get_gt_img / composite_with_background (splatfacto.py:900-923) and its clamp / mask / L1 / gamma lines."""
import os

import numpy as np
import pytest

from oracle import loss_oracle as LO

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_target.npz"))


@pytest.mark.parametrize("tag", ["rgb", "rgba"])
def test_loss_oracle_matches_the_reference_golden(tag):
    img, bg, lin, mask = G[f"{tag}_image"], G[f"{tag}_background"], G[f"{tag}_linear"], G[f"{tag}_mask"]
    assert np.array_equal(LO.prepare_target(img, bg), G[f"{tag}_gt"])
    for level in (0, 12):
        for use_mask in (0, 1):
            for gm in (0, 1):
                key = f"{tag}_l{level}_m{use_mask}_g{gm}"
                loss, grad, target = LO.l1_loss(lin, img, 2.2 if gm else None, bg, float(level), mask if use_mask else None)
                assert np.array_equal(target, G[key + "_target"]), key
                assert abs(float(loss) - float(G[key + "_loss"])) < 5e-7, key
                np.testing.assert_allclose(grad, G[key + "_grad"], rtol=2e-5, atol=1e-9, err_msg=key)


def test_golden_covers_the_cases_the_kernel_branches_on():
    a = G["rgba_image"][..., 3]
    assert (a == 0).any() and (a == 255).any() and ((a > 0) & (a < 255)).any()      # transparent, opaque, in between
    assert (G["rgb_linear"] > 1.0).any() and (G["rgb_linear"] < 1.0).any()          # both sides of the gamma step's clamp
    assert 0.1 < G["rgb_mask"].mean() < 0.9
    assert (G["rgb_image"].shape[0] * G["rgb_image"].shape[1]) % 4 != 0             # the kernel's 4-pixel path AND its tail
