import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "3dgs-deblur_b200")):
    sys.path.insert(0, p)
import torch
import gsplat.synthetic as synthetic
from gsplat import dp, rasterize_gaussians, spherical_harmonics
from gsplat.losses import l1_loss
from gsplat.sh import coeff_grad_sink

mode = sys.argv[1] if len(sys.argv) > 1 else "global"
if len(sys.argv) > 2 and sys.argv[2] == "nomt":
    torch.autograd.set_multithreading_enabled(False)
sc = synthetic.make_scene("c2", device="cuda", n_override=20000, n_cameras=1)
sc.update(H=128, W=160, fx=80.0, fy=80.0, cx=80.0, cy=64.0)
cam = sc["cameras"][0]
model = dp.FlatGaussians(sc, "cuda", n_cameras=1, optimize_velocities=True, sh_layout="block")
main = torch.cuda.Stream()
status = torch.zeros(4, dtype=torch.int32, device="cuda")
st = dict(cam=torch.cat([cam["viewmat"].reshape(-1), cam["lin_vel"], cam["ang_vel"], cam["cam_pos"]]).contiguous(), cam_index=torch.zeros(1, dtype=torch.int64, device="cuda"))
target = cam["target"][:128, :160].contiguous()


def attempt(name, fn, warm=2):
    try:
        with torch.cuda.stream(main):
            for _ in range(warm):
                fn()
                model.flat_grad.zero_()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=main, capture_error_mode=mode):
                fn()
            g.replay()
            torch.cuda.synchronize()
        print("OK  ", name, flush=True)
    except Exception as e:
        print("FAIL", name, repr(e)[:160].replace("\n", " "), flush=True)
        try:
            torch.cuda.synchronize()
        except Exception as e2:
            print("   (sync after failure:", repr(e2)[:100], ")")


with torch.cuda.stream(main):
    geo = dp.geometry_phase(model, st, sc, 1 << 20, status)
    torch.cuda.synchronize()

p = model.params
x = torch.randn(1000, device="cuda", requires_grad=True)
attempt("plain torch: (x*2).sum().backward()", lambda: (x * 2).sum().backward())
attempt("leaf param view: means.sum().backward()", lambda: p["means"].sum().backward())
attempt("exp/normalise glue backward", lambda: (torch.exp(p["log_scales"]).sum() + (p["quats"] / p["quats"].norm(dim=-1, keepdim=True)).sum()).backward())
attempt("index_select of cam rows backward", lambda: model.cam_vel.index_select(0, st["cam_index"])[0].sum().backward())
dirs = (p["means"].detach() - geo["cam_pos"]).contiguous()
attempt("sh fwd+bwd (autograd accumulate)", lambda: spherical_harmonics(3, dirs, model.sh_coeffs()).sum().backward())
sink = model.flat_grad[model.sh_start:model.sh_start + model.N * 48]


def sh_sink():
    with coeff_grad_sink(sink):
        c = spherical_harmonics(3, dirs, model.sh_coeffs())
    c.sum().backward()


attempt("sh fwd+bwd (sink)", sh_sink)
pred = torch.rand(128, 160, 3, device="cuda", requires_grad=True)
attempt("l1 loss fwd+bwd", lambda: l1_loss(pred, target).backward())
cols = torch.rand(model.N, 3, device="cuda", requires_grad=True)


def rast(backward, geo_grad):
    g_ = geo if geo_grad else {k: (v.detach() if torch.is_tensor(v) else v) for k, v in geo.items()}
    rgb, alpha = rasterize_gaussians(g_["xys"], g_["depths"], g_["pix_vels"], g_["radii"], g_["conics"], g_["num_tiles_hit"], cols,
                                     g_["opacities"], 128, 160, 16, rolling_shutter_time=sc["rolling_shutter_time"],
                                     exposure_time=sc["exposure_time"], blur_samples=g_["blur"], background=sc["background"],
                                     return_alpha=True, prepared=geo["prep"])
    if backward:
        rgb.sum().backward(retain_graph=geo_grad)


attempt("rasterize fwd only", lambda: rast(False, False))
attempt("rasterize fwd+bwd, geometry detached", lambda: rast(True, False))
attempt("rasterize fwd+bwd through the projection (graph of phase A retained)", lambda: rast(True, True))


def geo_and_shade():
    g2 = dp.geometry_phase(model, st, sc, 1 << 20, status)
    dp.shading_phase(model, g2, sc, target, l1_loss, 3)


attempt("A + B in ONE capture", geo_and_shade)
