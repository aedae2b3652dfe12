"""GPU parity of the fused loss kernels against the CPU oracles: L1 vs torch, SSIM / photometric loss vs
oracle/ssim_oracle.py (fp64 autograd restatement of pytorch_msssim; parity of that restatement is unpinned, see its
header).  Tolerances: values 5e-6 absolute (fp32 filtering of 121 taps and E[x^2] - mu^2 in fp32 vs fp64: the library
itself is 1e-7 .. 2e-6 off the fp64 result on these images); cotangents 1e-4 of the tensor's max magnitude + 1e-3
relative (same conditioning)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ssim_oracle as SO

if torch.cuda.is_available():
    from gsplat.losses import l1_loss, photometric_loss, ssim


def _pair(H, W, C, seed, smooth=True):
    g = torch.Generator().manual_seed(seed)
    t = torch.rand(H, W, C, generator=g)
    if smooth:  # image-like content: low-pass noise, so local variances are small and SSIM is far from 0
        k = torch.ones(C, 1, 5, 5) / 25
        t = torch.nn.functional.conv2d(t.permute(2, 0, 1)[None], k, padding=2, groups=C)[0].permute(1, 2, 0).contiguous()
    p = (t + 0.1 * torch.randn(H, W, C, generator=g)).clamp(0, 1).contiguous()
    return p, t


@pytest.mark.parametrize("H,W,C,smooth", [(800, 800, 3, True), (37, 53, 3, True), (11, 11, 3, False), (64, 27, 1, True),
                                          (48, 48, 4, False), (16, 200, 3, True)])
def test_ssim_and_photometric_loss_vs_oracle(H, W, C, smooth):
    p, t = _pair(H, W, C, seed=H * 1000 + W + C, smooth=smooth)
    p64 = p.double().requires_grad_(True)
    ref_ssim = SO.ssim_hwc(p64, t.double())
    (g_ssim,) = torch.autograd.grad(ref_ssim, p64)
    ref_loss = SO.photometric_loss(p64, t.double(), 0.2)
    (g_loss,) = torch.autograd.grad(ref_loss * 1.7, p64)

    pc = p.cuda().requires_grad_(True)
    tc = t.cuda()
    out = ssim(pc, tc)
    (g,) = torch.autograd.grad(out, pc)
    assert abs(float(out) - float(ref_ssim)) <= 5e-6
    tol = 1e-4 * float(g_ssim.abs().max())
    torch.testing.assert_close(g.cpu().double(), g_ssim, rtol=1e-3, atol=tol)

    loss = photometric_loss(pc, tc, 0.2)
    (gl,) = torch.autograd.grad(loss * 1.7, pc)
    assert abs(float(loss) - float(ref_loss)) <= 5e-6
    # the L1 part of the cotangent is +-1/n exactly; at |p - t| ~ 0 fp32 and fp64 may pick different signs: none here
    torch.testing.assert_close(gl.cpu().double(), g_loss, rtol=1e-3, atol=1e-4 * float(g_loss.abs().max()))
    # deterministic
    assert torch.equal(ssim(pc, tc), out) and torch.equal(photometric_loss(pc, tc, 0.2), loss)
    # lambda = 0 is the plain L1, lambda = 1 is 1 - SSIM
    torch.testing.assert_close(photometric_loss(pc, tc, 0.0), l1_loss(pc, tc), rtol=1e-6, atol=0)
    torch.testing.assert_close(photometric_loss(pc, tc, 1.0), 1 - out, rtol=1e-6, atol=1e-7)


def test_ssim_identities_and_errors():
    p, t = _pair(40, 56, 3, seed=1)
    pc, tc = p.cuda(), t.cuda()
    assert abs(float(ssim(tc, tc)) - 1.0) < 1e-6  # SSIM(x, x) = 1
    assert abs(float(ssim(pc, tc)) - float(ssim(tc, pc))) < 1e-6  # symmetric
    with pytest.raises(ValueError):
        ssim(pc[:10], tc[:10])  # shorter than the 11-tap window
    with pytest.raises(ValueError):
        ssim(pc, tc[:, :50].contiguous())
    with pytest.raises(RuntimeError):
        ssim(p, t)  # CPU tensors: no CPU path
    with pytest.raises(RuntimeError):
        photometric_loss(pc.double(), tc.double())


def test_l1_loss_with_the_gamma_step_folded_in():
    """l1_loss(linear, target, gamma) == mean |clamp(linear, max=1) ** (1 / gamma) - target| (splatfacto.py:879-880 then :957)
    and its autograd cotangent w.r.t. the LINEAR image, incl. the zero slope above the clamp."""
    from gsplat.losses import l1_loss

    g = torch.Generator(device="cuda").manual_seed(3)
    for shape in ((800, 800, 3), (37, 53, 3)):
        lin = (torch.rand(shape, device="cuda", generator=g) * 1.4 + 0.01).requires_grad_(True)  # some pixels above 1
        with torch.no_grad():
            lin.view(-1)[1] = 1.0  # exactly at the clamp: the gradient still flows (torch's clamp backward is inclusive)
        target = torch.rand(shape, device="cuda", generator=g)
        ref = (torch.clamp(lin, max=1.0) ** (1.0 / 2.2) - target).abs().mean()
        (g_ref,) = torch.autograd.grad(ref * 2.0, lin)
        out = l1_loss(lin, target, gamma=2.2)
        (g_out,) = torch.autograd.grad(out * 2.0, lin)
        torch.testing.assert_close(out, ref, rtol=3e-6, atol=0)
        torch.testing.assert_close(g_out, g_ref, rtol=2e-5, atol=1e-12)
        assert float(g_out[lin.detach() > 1.0].abs().max()) == 0.0 and float(g_out.view(-1)[1].abs()) > 0.0
    with pytest.raises(ValueError):
        l1_loss(lin, target, gamma=0.0)


@pytest.mark.parametrize("tag", ["rgb", "rgba"])
def test_l1_on_the_uint8_dataset_image_vs_reference_golden(tag):
    """b200_l1_loss_u8 against tests/golden/loss_target.npz -- get_gt_img / composite_with_background of the REFERENCE
    model (imported by tests/golden/make_golden_loss.py) followed by its clamp / mask / L1 lines: prepared target bit
    for bit, loss to 5e-7, cotangent w.r.t. the (linear) render to float rounding of pow."""
    import os

    import numpy as np

    from gsplat.losses import l1_loss_and_grad
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_target.npz"))
    img = torch.from_numpy(G[f"{tag}_image"]).cuda()
    bg = torch.from_numpy(G[f"{tag}_background"]).cuda()
    linear = torch.from_numpy(G[f"{tag}_linear"]).cuda()
    mask = torch.from_numpy(G[f"{tag}_mask"]).cuda()
    for level in (0, 12):
        for use_mask in (0, 1):
            for gm in (0, 1):
                key = f"{tag}_l{level}_m{use_mask}_g{gm}"
                loss, grad, target = l1_loss_and_grad(linear, img, 2.2 if gm else None, background=bg, min_rgb_level=float(level),
                                                      mask=mask if use_mask else None, return_target=True)
                assert torch.equal(target.cpu(), torch.from_numpy(G[key + "_target"])), key
                assert abs(float(loss) - float(G[key + "_loss"])) < 5e-7, key
                ref = torch.from_numpy(G[key + "_grad"])
                # (sign flips where |pred - target| is a rounding error of pow: none on this fixture)
                torch.testing.assert_close(grad.cpu(), ref, rtol=2e-5, atol=1e-9, msg=key)
    # autograd form, and the trainer-facing photometric loss on the same uint8 image
    p = linear.clone().requires_grad_(True)
    out = l1_loss(p, img, 2.2, background=bg, min_rgb_level=12.0, mask=mask)
    out.backward()
    torch.testing.assert_close(p.grad.cpu(), torch.from_numpy(G[f"{tag}_l12_m1_g1_grad"]), rtol=2e-5, atol=1e-9)
    pred = linear.clamp(max=1.0)
    want = photometric_loss(pred, torch.from_numpy(G[f"{tag}_l12_m0_g0_target"]).cuda(), 0.2)
    got = photometric_loss(pred, img, 0.2, background=bg, min_rgb_level=12.0)
    assert torch.equal(want, got)
    with pytest.raises(ValueError):
        l1_loss(linear, torch.from_numpy(G[f"{tag}_l0_m0_g0_target"]).cuda(), mask=mask)  # options need the uint8 image
    if tag == "rgba":
        with pytest.raises(ValueError):
            l1_loss(linear, img)  # RGBA without a background colour


def test_pipelined_trainer_takes_the_uint8_image_without_a_conversion_pass():
    """gsplat.dp hands the dataset's uint8 image to gsplat.losses.l1_loss as it is: same loss as the float image."""
    import gsplat.synthetic as synthetic
    from gsplat.dp import FlatGaussians, PipelinedTrainer

    losses = []
    for as_u8 in (False, True):
        sc = synthetic.make_scene("c2", device="cuda", n_override=20000, n_cameras=1)
        sc.update(H=128, W=160, fx=80.0, fy=80.0, cx=80.0, cy=64.0)
        cam = sc["cameras"][0]
        cam.update(fx=80.0, fy=80.0, cx=80.0, cy=64.0, vel0=torch.cat([cam["lin_vel"], cam["ang_vel"]]))
        u8 = (torch.rand(128, 160, 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) * 255).to(torch.uint8)
        model = FlatGaussians(sc, "cuda", n_cameras=1, optimize_velocities=True)
        tr = PipelinedTrainer(model, sc, use_graphs=False)
        tr.prepare(cam, 0)
        loss = tr.train_step(u8 if as_u8 else u8.float() / 255)
        losses.append(float(loss))
        tr.finish()
    assert abs(losses[0] - losses[1]) < 1e-7, losses
