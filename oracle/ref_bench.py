"""`bench.py --impl refgpu`: the train step of bench.py on the UNMODIFIED reference CUDA kernels (oracle/_ref),
same scene, same loss, same fused Adam, 1 GPU.  BENCH INFRASTRUCTURE ONLY (see ref_ops.py)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def measure(config, n_override, n_img, steps, warmup, variant="cuda", breakdown=False):
    """Train-step images/s of bench.py's workload on the UNMODIFIED reference CUDA kernels (oracle/_ref), 1 GPU.
    variant: "cuda"   velocities constant -> the reference's CUDA projection (its faster mode);
             "torch"  velocities require grad (train.py's default) -> the reference's PyTorch projection path;
             "static" zero camera velocities -> no blur-inflated radii, hence none of the phantom tile-0 entries that make
                      one CTA of the reference walk ~24k extra list entries per sample."""
    sys.path.insert(0, os.path.join(ROOT, "3dgs-deblur_b200"))
    import ref_ops
    from gsplat import synthetic
    from gsplat.dp import FlatGaussians

    dev = torch.device("cuda", torch.cuda.current_device())
    scene = synthetic.make_scene(config, device="cpu", n_override=n_override, n_cameras=n_img)
    scene_dev = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in scene.items() if k != "cameras"}
    zero = torch.zeros(3, device=dev)
    cams = []
    for c in scene["cameras"]:
        lin, ang = (zero, zero) if variant == "static" else (c["lin_vel"].to(dev), c["ang_vel"].to(dev))
        if variant == "torch":
            lin, ang = lin.clone().requires_grad_(True), ang.clone().requires_grad_(True)
        cams.append(dict(viewmat=c["viewmat"].to(dev), fx=c["fx"], fy=c["fy"], cx=c["cx"], cy=c["cy"], cam_pos=c["cam_pos"].to(dev),
                         lin_vel=lin, ang_vel=ang))
    targets = [c["target"].to(dev) for c in scene["cameras"]]
    model = FlatGaussians(scene_dev, dev)
    params = model.parameters() + ([t for c in cams for t in (c["lin_vel"], c["ang_vel"])] if variant == "torch" else [])
    opt = torch.optim.Adam(params, lr=1e-4, eps=1e-15, fused=True)

    def step(k):
        i = k % n_img
        opt.zero_grad(set_to_none=False) if variant == "torch" else model.zero_grad()
        rgb, alpha = ref_ops.render(model, cams[i], scene_dev, torch_projection=(variant == "torch"))
        loss = (rgb - targets[i]).abs().mean()
        loss.backward()
        opt.step()
        return loss

    for k in range(warmup):
        step(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(steps):
        step(k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"value": 1000.0 / ms, "unit": "images/s", "ms_per_step": ms, "steps": steps, "variant": variant}
    if breakdown:
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for k in range(3):
                step(k)
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:8]
        out["kernels_ms_per_step"] = {r.key[:60]: round(r.device_time_total / 3 / 1000.0, 4) for r in rows}
    return out


def run(args):
    sys.path.insert(0, os.path.join(ROOT, "3dgs-deblur_b200"))
    import ref_ops
    from gsplat import synthetic
    from gsplat.dp import FlatGaussians

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    n_img = args.images
    scene = synthetic.make_scene(args.config, device="cpu", n_override=args.n, n_cameras=n_img)
    scene_dev = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in scene.items() if k != "cameras"}
    cams = [dict(viewmat=c["viewmat"].to(dev), fx=c["fx"], fy=c["fy"], cx=c["cx"], cy=c["cy"], cam_pos=c["cam_pos"].to(dev),
                 lin_vel=c["lin_vel"].to(dev), ang_vel=c["ang_vel"].to(dev)) for c in scene["cameras"]]
    targets = [c["target"].to(dev) for c in scene["cameras"]]
    model = FlatGaussians(scene_dev, dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, eps=1e-15, fused=True)

    def step(k):
        i = k % n_img
        model.zero_grad()
        rgb, alpha = ref_ops.render(model, cams[i], scene_dev)
        loss = (rgb - targets[i]).abs().mean()
        loss.backward()
        opt.step()
        return loss

    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        step(k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps

    # per-kernel breakdown of one step with the torch profiler (names of the reference's own kernels)
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for k in range(3):
            step(k)
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:12]
    kernels = {r.key[:60]: round(r.device_time_total / 3 / 1000.0, 4) for r in rows}
    out = {"impl": "refgpu", "metric": "train images/sec (fwd+bwd) at N=5 blur samples", "value": 1000.0 / ms, "unit": "images/s",
           "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
           "config": {"workload": f"{args.config}: {scene['N']} Gaussians, {scene['W']}x{scene['H']}, S={scene['blur_samples']}",
                      "arm": "unmodified reference gsplat CUDA kernels (oracle/_ref, -O3 --use_fast_math, sm_100), velocities constant "
                             "(the reference's faster mode), same L1 loss + fused Adam"},
           "kernels_ms_per_step": kernels}
    import __main__ as _bench_main  # bench.py owns stdout: it prints exactly one JSON line at exit

    _bench_main._emit(out)
