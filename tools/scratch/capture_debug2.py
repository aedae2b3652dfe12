import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for t in ["sh_fwd_nograd", "sh_fwd_grad", "sh_bwd_c", "ident_fn", "sh_fb", "sh_fb_lazyoff", "l1_fb", "proj_fb", "rast_f", "rast_fb"]:
        env = dict(os.environ)
        if t.endswith("_lazyoff"):
            env["CUDA_MODULE_LOADING"] = "EAGER"
        r = subprocess.run([sys.executable, __file__, t], capture_output=True, text=True, env=env)
        out = (r.stdout + r.stderr).strip().splitlines()
        print(t, "->", [l for l in out if l.startswith(("OK", "FAIL"))] or out[-3:], flush=True)
    sys.exit(0)
for p in (ROOT, os.path.join(ROOT, "3dgs-deblur_b200")):
    sys.path.insert(0, p)
import torch
import gsplat.synthetic as synthetic
import gsplat.cuda as _C
from gsplat import dp, rasterize_gaussians, spherical_harmonics, project_gaussians
from gsplat.losses import l1_loss

t = sys.argv[1].replace("_lazyoff", "")
sc = synthetic.make_scene("c2", device="cuda", n_override=20000, n_cameras=1)
sc.update(H=128, W=160, fx=80.0, fy=80.0, cx=80.0, cy=64.0)
cam = sc["cameras"][0]
model = dp.FlatGaussians(sc, "cuda", n_cameras=1, optimize_velocities=True, sh_layout="block")
main = torch.cuda.Stream()
status = torch.zeros(4, dtype=torch.int32, device="cuda")
st = dict(cam=torch.cat([cam["viewmat"].reshape(-1), cam["lin_vel"], cam["ang_vel"], cam["cam_pos"]]).contiguous(), cam_index=torch.zeros(1, dtype=torch.int64, device="cuda"))
target = cam["target"][:128, :160].contiguous()
p = model.params
dirs = (p["means"].detach() - cam["cam_pos"]).contiguous()
v = torch.randn(model.N, 3, device="cuda")


class Ident(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x * 1.0

    @staticmethod
    def backward(ctx, g):
        return g * 2.0


with torch.cuda.stream(main):
    geo = dp.geometry_phase(model, st, sc, 1 << 20, status)
torch.cuda.synchronize()
cols = torch.rand(model.N, 3, device="cuda", requires_grad=True)
pred = torch.rand(128, 160, 3, device="cuda", requires_grad=True)


def rast(backward):
    g_ = {k: (v_.detach() if torch.is_tensor(v_) else v_) for k, v_ in geo.items()}
    rgb, alpha = rasterize_gaussians(g_["xys"], g_["depths"], g_["pix_vels"], g_["radii"], g_["conics"], g_["num_tiles_hit"], cols,
                                     g_["opacities"], 128, 160, 16, rolling_shutter_time=sc["rolling_shutter_time"],
                                     exposure_time=sc["exposure_time"], blur_samples=g_["blur"], background=sc["background"],
                                     return_alpha=True, prepared=geo["prep"])
    if backward:
        rgb.sum().backward()


def proj_fb():
    q = p["quats"] / p["quats"].norm(dim=-1, keepdim=True)
    out = project_gaussians(p["means"], torch.exp(p["log_scales"]), 1, q, cam["lin_vel"], cam["ang_vel"], sc["rolling_shutter_time"],
                            sc["exposure_time"], cam["viewmat"], 80.0, 80.0, 80.0, 64.0, 128, 160, 16)
    (out[0].sum() + out[4].sum()).backward()


def no_grad_sh():
    with torch.no_grad():
        spherical_harmonics(3, dirs, model.sh_coeffs())


fns = {
    "sh_fwd_nograd": no_grad_sh,
    "sh_fwd_grad": lambda: spherical_harmonics(3, dirs, model.sh_coeffs()),
    "sh_bwd_c": lambda: _C.compute_sh_backward("fast", model.N, 3, 3, dirs, v),
    "ident_fn": lambda: Ident.apply(p["means"]).sum().backward(),
    "sh_fb": lambda: spherical_harmonics(3, dirs, model.sh_coeffs()).sum().backward(),
    "l1_fb": lambda: l1_loss(pred, target).backward(),
    "proj_fb": proj_fb,
    "rast_f": lambda: rast(False),
    "rast_fb": lambda: rast(True),
}
fn = fns[t]
try:
    with torch.cuda.stream(main):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=main):
            fn()
        g.replay()
        torch.cuda.synchronize()
    print("OK", t)
except Exception as e:
    print("FAIL", t, repr(e)[:300].replace("\n", " "))
