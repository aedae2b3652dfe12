"""GPU: gsplat.data.ImagePrefetcher delivers every (image, camera) pair intact and in order while copies overlap compute."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_image_prefetcher_order_and_contents():
    from gsplat.data import ImagePrefetcher
    g = torch.Generator().manual_seed(3)
    images = [torch.randint(0, 256, (96, 128, 3), dtype=torch.uint8, generator=g) for _ in range(5)]
    cams = [torch.randn(21, generator=g) for _ in range(5)]
    pf = ImagePrefetcher(images, cams, "cuda")
    assert pf.bytes_per_step == 96 * 128 * 3 + 84
    order = [0, 3, 1, 4, 2, 2, 0, 1]
    busy = torch.zeros(1 << 22, device="cuda")
    for rep in range(2):  # start() may be called again
        pf.start(order[0])
        for k, idx in enumerate(order):
            nxt = order[k + 1] if k + 1 < len(order) else None
            img, cam = pf.get(next_index=nxt)
            got_img, got_cam = img.clone(), cam.clone()  # consumed on the compute stream ...
            for _ in range(4):
                busy.add_(1.0)  # ... while the next copy is in flight
            pf.done()
            assert torch.equal(got_img.cpu(), images[idx]) and torch.equal(got_cam.cpu(), cams[idx])
    with pytest.raises(ValueError):
        ImagePrefetcher(images, cams[:-1], "cuda")
    with pytest.raises(ValueError):
        ImagePrefetcher([images[0].float()], [cams[0]], "cuda")
