"""Image-sharded data-parallel training dispatcher for the rasterizer hot path.

Replaces the reference's one-random-image-per-step dispatcher (nerfstudio/data/datamanagers/
full_images_datamanager.py:287-304) + the DDP wrap Splatfacto cannot actually use
(pipelines/base_pipeline.py:282-284; SURVEY section 5 "distributed") for this path:

  * all Gaussian parameters (means 3, log-scales 3, quats 4, SH dc 3, SH rest 3(K-1), opacity 1 = 59
    floats per Gaussian at K=16) live in ONE flat fp32 buffer, their gradients in a second one;
  * every rank renders its own image of the step (rank r takes image step*R + r) through the public
    gsplat operators -- the render block of splatfacto.py:816-880 -- and back-propagates;
  * the flat gradient buffer is averaged across ranks once per step, as two contiguous slices (SH block / the rest) so
    that the exchange overlaps the backward pass and the optimizer (NCCL over NVLink on GPUs, gloo in the CPU tests),
    then the same fused Adam update on every rank, so replicas stay bit-identical;
  * per-camera parameters (velocities) are disjoint rows: their gradients ride in the same buffer.

The densification statistics (xys.absgrad norms, visibility counts, max 2D radius;
splatfacto.py:417-434) are exposed by `reduce_densify_stats` with the matching sum / sum / max
reductions so a caller can keep topology changes identical across ranks.
"""
import math
import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

# flat layout: the geometry/opacity rows (11 floats per Gaussian) and the per-camera rows first, the SH coefficients
# (48 of the 59 floats at K = 16) last, so "everything but SH" and "SH" are each ONE contiguous slice to exchange
FIELDS = ("means", "log_scales", "quats", "opacity_logit", "sh_dc", "sh_rest")


class FlatGaussians:
    """Gaussian parameters as views into one flat buffer (+ a flat gradient buffer of the same layout)."""

    def __init__(self, scene: Dict, device, n_cameras: int = 0, optimize_velocities: bool = False, sh_layout: str = "split"):
        """sh_layout "split": sh_dc (N,1,3) and sh_rest (N,K-1,3) are two parameters like Splatfacto's features_dc /
        features_rest (the render block concatenates them every step, splatfacto.py:840-842; gsplat.fused reads them
        separately).  "block": ONE (N,K,3) parameter `sh` in the layout spherical_harmonics consumes -- no per-step
        torch.cat (a 57 MB copy at 300k Gaussians) and no cat backward; `sh_dc` / `sh_rest` remain readable as views."""
        if sh_layout not in ("split", "block"):
            raise ValueError("sh_layout must be 'split' or 'block'")
        N = scene["means"].shape[0]
        K = scene["sh_rest"].shape[1] + 1
        self.N, self.K, self.sh_layout = N, K, sh_layout
        widths = dict(means=3, log_scales=3, quats=4, opacity_logit=1, sh_dc=3, sh_rest=3 * (K - 1))
        shapes = dict(means=(N, 3), log_scales=(N, 3), quats=(N, 4), opacity_logit=(N, 1), sh_dc=(N, 1, 3),
                      sh_rest=(N, K - 1, 3))
        extra = 6 * n_cameras if optimize_velocities else 0
        # pad the camera rows so the SH coefficients start on a 16-byte boundary (the slice-wise device-state Adam and
        # the SH kernels' vector accesses want that)
        pad = (-(N * 11 + extra)) % 4
        total = N * sum(widths.values()) + extra + pad
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.flat_grad = torch.zeros_like(self.flat)
        self.params: Dict[str, torch.Tensor] = {}
        self.slices: Dict[str, tuple] = {}
        self.cam_vel: Optional[torch.Tensor] = None

        def take(name, off, numel, shape):
            view = self.flat[off:off + numel].view(shape)
            p = view.requires_grad_(True)
            p.grad = self.flat_grad[off:off + numel].view(shape)
            self.slices[name] = (off, off + numel)
            return p

        off = 0
        for name in FIELDS:
            if name == "sh_dc":  # camera rows (+ alignment padding) sit between the geometry block and the SH block
                if extra:
                    self.cam_vel = take("cam_vel", off, extra, (n_cameras, 6))
                off += extra + pad
                if sh_layout == "block":
                    blk = self.flat[off:off + N * 3 * K].view(N, K, 3)
                    blk[:, :1].copy_(scene["sh_dc"].to(device).reshape(N, 1, 3))
                    blk[:, 1:].copy_(scene["sh_rest"].to(device).reshape(N, K - 1, 3))
                    self.params["sh"] = take("sh", off, N * 3 * K, (N, K, 3))
                    self.slices["sh_dc"] = self.slices["sh_rest"] = self.slices["sh"]
                    off += N * 3 * K
                    break
            self.flat[off:off + N * widths[name]].view(shapes[name]).copy_(scene[name].to(device).reshape(shapes[name]))
            self.params[name] = take(name, off, N * widths[name], shapes[name])
            off += N * widths[name]
        assert off == total
        self.sh_start = self.slices["sh_dc"][0]  # flat[sh_start:] = all SH coefficients
        self.floats_per_gaussian = sum(widths.values())

    @property
    def views(self) -> Dict[str, torch.Tensor]:
        """Block layout: Splatfacto's two SH tensors as read-only (detached) views of the one block."""
        sh = self.params["sh"].detach()
        return {"sh_dc": sh[:, :1], "sh_rest": sh[:, 1:]}

    def rebind_leaves(self):
        """Replace every parameter by a fresh leaf tensor over the same storage (same .grad views).  Autograd ties a
        leaf's gradient accumulator to the stream that is current when the accumulator is first created; a trainer that
        captures its backward pass into a CUDA graph on its own stream calls this (on that stream) so that no accumulator
        made earlier on the legacy default stream forces the capture to synchronise with it."""
        for name in list(self.params):
            old = self.params[name]
            new = old.detach().requires_grad_(True)
            new.grad = old.grad
            self.params[name] = new
        if self.cam_vel is not None:
            new = self.cam_vel.detach().requires_grad_(True)
            new.grad = self.cam_vel.grad
            self.cam_vel = new

    def _reallocate(self, new_n: int):
        """New flat / gradient buffers for `new_n` Gaussians with the same field order, camera rows and layout rules; the
        parameter views are rebuilt (contents uninitialised: gsplat.densify.Densifier fills them).  Anything that cached
        pointers into the old buffers (optimizer, captured graphs) must be rebound by the caller."""
        K, dev = self.K, self.flat.device
        n_cam = 0 if self.cam_vel is None else self.cam_vel.shape[0]
        extra = 6 * n_cam
        pad = (-(new_n * 11 + extra)) % 4
        widths = dict(means=3, log_scales=3, quats=4, opacity_logit=1, sh_dc=3, sh_rest=3 * (K - 1))
        shapes = dict(means=(new_n, 3), log_scales=(new_n, 3), quats=(new_n, 4), opacity_logit=(new_n, 1), sh_dc=(new_n, 1, 3),
                      sh_rest=(new_n, K - 1, 3))
        total = new_n * sum(widths.values()) + extra + pad
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros_like(self.flat)
        self.params, self.slices, self.cam_vel, self.N = {}, {}, None, new_n

        def take(name, off, numel, shape):
            pr = self.flat[off:off + numel].view(shape).requires_grad_(True)
            pr.grad = self.flat_grad[off:off + numel].view(shape)
            self.slices[name] = (off, off + numel)
            return pr

        off = 0
        for name in FIELDS:
            if name == "sh_dc":
                if extra:
                    self.cam_vel = take("cam_vel", off, extra, (n_cam, 6))
                off += extra + pad
                if self.sh_layout == "block":
                    self.params["sh"] = take("sh", off, new_n * 3 * K, (new_n, K, 3))
                    self.slices["sh_dc"] = self.slices["sh_rest"] = self.slices["sh"]
                    off += new_n * 3 * K
                    break
            self.params[name] = take(name, off, new_n * widths[name], shapes[name])
            off += new_n * widths[name]
        assert off == total
        self.sh_start = self.slices["sh_dc"][0]

    def sh_coeffs(self) -> torch.Tensor:
        """(N, K, 3) coefficients as spherical_harmonics wants them."""
        if self.sh_layout == "block":
            return self.params["sh"]
        return torch.cat((self.params["sh_dc"], self.params["sh_rest"]), dim=1)  # splatfacto.py:840-842

    def parameters(self) -> List[torch.Tensor]:
        ps = list(self.params.values())
        if self.cam_vel is not None:
            ps.append(self.cam_vel)
        return ps

    def zero_grad(self):
        self.flat_grad.zero_()


def render(model: FlatGaussians, cam: Dict, scene: Dict, cam_index: int = 0, sh_degree_to_use: int = 3):
    """The Splatfacto render block (splatfacto.py:816-880) on the public operators.  Returns (rgb, alpha, xys)."""
    from gsplat import project_gaussians, rasterize_gaussians, spherical_harmonics

    p = model.params
    if model.cam_vel is not None:
        vel = model.cam_vel[cam_index] + cam["vel0"]
        lin, ang = vel[:3].unsqueeze(0), vel[3:].unsqueeze(0)
    else:
        lin, ang = cam["lin_vel"].unsqueeze(0), cam["ang_vel"].unsqueeze(0)
    H, W, bw = scene["H"], scene["W"], scene["block_width"]
    quats = p["quats"] / p["quats"].norm(dim=-1, keepdim=True)
    xys, depths, pix_vels, radii, conics, comp, num_tiles_hit, _ = project_gaussians(
        p["means"], torch.exp(p["log_scales"]), 1, quats, lin, ang, scene["rolling_shutter_time"],
        scene["exposure_time"], cam["viewmat"], cam["fx"], cam["fy"], cam["cx"], cam["cy"], H, W, bw)
    colors = model.sh_coeffs()
    viewdirs = p["means"].detach() - cam["cam_pos"]
    rgbs = torch.clamp(spherical_harmonics(sh_degree_to_use, viewdirs, colors) + 0.5, min=0.0)
    opacities = torch.sigmoid(p["opacity_logit"]) * comp[:, None]  # "antialiased" mode, splatfacto.py:853-854
    blur = scene["blur_samples"] if scene["exposure_time"] > 0 else 1
    rgb, alpha = rasterize_gaussians(xys, depths, pix_vels, radii, conics, num_tiles_hit, rgbs, opacities, H, W, bw,
                                     rolling_shutter_time=scene["rolling_shutter_time"],
                                     exposure_time=scene["exposure_time"], blur_samples=blur,
                                     background=scene["background"], return_alpha=True)
    return rgb, alpha, xys, radii


def render_fused(model: FlatGaussians, cam: Dict, scene: Dict, cam_index: int = 0, sh_degree_to_use: int = 3,
                 write_grads_in_place: bool = True):
    """Same render through gsplat.fused.render_gaussians (raw parameters in, one operator).  With
    `write_grads_in_place` the backward kernel overwrites the model's flat gradient buffer directly."""
    from gsplat.fused import render_gaussians

    p = model.params
    if model.cam_vel is not None:
        vel = model.cam_vel[cam_index] + cam["vel0"]
        lin, ang = vel[:3], vel[3:]
    else:
        lin, ang = cam["lin_vel"], cam["ang_vel"]
    sink = {k: v.grad for k, v in p.items()} if write_grads_in_place else None
    blur = scene["blur_samples"] if scene["exposure_time"] > 0 else 1
    rgb, alpha, info = render_gaussians(
        p["means"], p["log_scales"], p["quats"], p["opacity_logit"], p["sh_dc"], p["sh_rest"], cam["viewmat"],
        cam["cam_pos"], lin, ang, cam["fx"], cam["fy"], cam["cx"], cam["cy"], scene["H"], scene["W"],
        scene["block_width"], scene["background"], rolling_shutter_time=scene["rolling_shutter_time"],
        exposure_time=scene["exposure_time"], blur_samples=blur, sh_degree_to_use=sh_degree_to_use, grad_sink=sink)
    return rgb, alpha, info


class ImageShardedTrainer:
    """One process per GPU; rank r renders image (step * world + r) % n_images; one gradient exchange per step."""

    def __init__(self, model: FlatGaussians, scene: Dict, lr: float = 1e-3, group=None, overlap_sh: bool = True,
                 fused: bool = False, loss_fn=None, sh_chunks: int = 1, optimizer: str = "b200"):
        self.model, self.scene = model, scene
        if loss_fn is None:  # fused L1 (value + cotangent in one kernel); CUDA only, like the operators themselves
            from gsplat.losses import l1_loss as loss_fn
        self.loss_fn = loss_fn
        self.fused = fused  # render through gsplat.fused (caller-modified path) instead of the three drop-in operators
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.group = group
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.world = dist.get_world_size(group) if self.distributed else 1
        # The exchange is split where the backward pass splits: the SH coefficients are 48 of the 59 floats per Gaussian
        # and their gradient is final as soon as the SH backward has run -- before the projection backward.  That part
        # of the flat gradient buffer is reduced asynchronously from a post-accumulate hook, so most of the step's
        # exchange overlaps the rest of the backward pass; the remaining 11 floats per Gaussian (+ camera rows) follow
        # when backward returns.  Exchange and update are pipelined over contiguous CHUNKS of the flat buffers (the
        # geometry block + `sh_chunks` pieces of the SH block): chunk i is updated as soon as its allreduce has landed,
        # while the allreduce of chunk i+1 is still on the wire.  Adam is per-element and every chunk sees the same step
        # count and learning rate, so this is the same update as one optimizer over all parameters (a per-group
        # learning rate would need chunk boundaries on parameter boundaries).  Measured at 2 GPUs: every extra collective
        # costs more fixed latency on the NCCL stream than its pipelining hides (969 / 910 images/s with 1 / 3 SH
        # chunks), hence the default of one SH chunk + the geometry chunk.
        # optimizer = "b200": gsplat.optim.FlatAdam (one kernel per slice, clears the gradient slice in the same pass);
        # "torch": torch.optim.Adam per chunk (what the CPU/gloo tests of this host logic inject).
        if optimizer not in ("b200", "torch"):
            raise ValueError("optimizer must be 'b200' or 'torch'")
        self.flat_adam = None
        if optimizer == "b200":
            from gsplat.optim import FlatAdam
            self.flat_adam = FlatAdam(model.flat, model.flat_grad, lr=lr, eps=1e-15)
        fused_adam = model.flat.is_cuda
        self.step_idx = 0
        # NCCL averages in the collective itself; gloo (CPU tests) only sums, so the 1/R scale is a separate pass there
        self._avg = self.distributed and dist.get_backend(group) == "nccl"
        self._op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self._sh_works = None
        self._sh_seen = 0
        if self.distributed:
            total, lo = model.flat.numel(), model.sh_start
            n_sh = max(1, int(sh_chunks))
            cuts = [lo + (total - lo) * k // n_sh for k in range(n_sh + 1)]
            bounds = [(0, lo)] + [(cuts[k], cuts[k + 1]) for k in range(n_sh) if cuts[k + 1] > cuts[k]]
            self._chunks = []  # (gradient slice, (begin, end) | torch Adam over an alias of the matching parameter slice)
            for a, b_ in bounds:
                if self.flat_adam is not None:
                    self._chunks.append((model.flat_grad[a:b_], (a, b_)))
                    continue
                alias = model.flat[a:b_].detach().requires_grad_(True)  # shares storage with the parameter views
                alias.grad = model.flat_grad[a:b_]
                self._chunks.append((alias.grad, torch.optim.Adam([alias], lr=lr, eps=1e-15, fused=fused_adam)))
        elif self.flat_adam is None:  # nothing to overlap with: one multi-tensor launch over the parameter views
            self.opt = torch.optim.Adam(model.parameters(), lr=lr, eps=1e-15, fused=fused_adam)
        # the fused operator writes every gradient in one kernel: nothing is final early, but the chunked
        # exchange / update pipeline still applies
        self.overlap_sh = bool(overlap_sh and self.distributed and not self.fused)
        if self.overlap_sh:
            self._sh_params = ("sh",) if model.sh_layout == "block" else ("sh_dc", "sh_rest")
            for name in self._sh_params:
                model.params[name].register_post_accumulate_grad_hook(self._on_sh_grad)

    def _reduce_async(self, chunk_ids):
        return [dist.all_reduce(self._chunks[i][0], op=self._op, group=self.group, async_op=True) for i in chunk_ids]

    def _on_sh_grad(self, _param):
        self._sh_seen += 1
        if self._sh_seen == len(self._sh_params):  # every SH parameter's gradient has landed in the flat buffer
            self._sh_works = self._reduce_async(range(1, len(self._chunks)))

    def image_index(self, step: int, n_images: int) -> int:
        return (step * self.world + self.rank) % n_images

    def train_step(self, cam: Dict, target: torch.Tensor, cam_index: int = 0):
        """fwd + L1 loss + bwd (+ gradient exchange) + Adam.  Returns the (device) loss tensor; no host sync."""
        m = self.model
        clear = self.flat_adam is None  # FlatAdam leaves the gradient buffer zeroed behind it
        if self.fused:
            if clear and m.cam_vel is not None:
                m.cam_vel.grad.zero_()  # the Gaussian rows are overwritten by the fused backward kernel
            rgb, alpha, info = render_fused(m, cam, self.scene, cam_index)
            xys, radii = None, info["radii"]
        else:
            if clear:
                m.zero_grad()
            rgb, alpha, xys, radii = render(m, cam, self.scene, cam_index)
        loss = self.loss_fn(rgb, target)
        self._sh_seen, self._sh_works = 0, None
        loss.backward()
        if self.distributed:
            # gradients of the R images are averaged (each rank's loss is a per-image mean)
            n = len(self._chunks)
            if self._sh_works is not None:  # SH chunks already on the wire (hook): geometry chunk goes last
                order = list(range(1, n)) + [0]
                works = self._sh_works + self._reduce_async([0])
            else:
                order = list(range(n))
                works = self._reduce_async(order)
            if self.flat_adam is not None:
                self.flat_adam.begin_step()
            for i, w in zip(order, works):
                g, opt = self._chunks[i]
                w.wait()
                if self.flat_adam is not None:  # the 1/R of a summing backend is folded into the update
                    self.flat_adam.update(opt[0], opt[1], 1.0 if self._avg else 1.0 / self.world, True)
                    continue
                if not self._avg:
                    g.mul_(1.0 / self.world)
                opt.step()
        elif self.flat_adam is not None:
            self.flat_adam.step()
        else:
            self.opt.step()
        self.step_idx += 1
        self.last_xys, self.last_radii = xys, radii
        return loss

    def reduce_densify_stats(self, grad_norm_sum: torch.Tensor, vis_counts: torch.Tensor, max_2d: torch.Tensor):
        """sum / sum / max reductions of the densification statistics (splatfacto.py:417-434)."""
        if self.distributed:
            dist.all_reduce(grad_norm_sum, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(vis_counts, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(max_2d, op=dist.ReduceOp.MAX, group=self.group)
        return grad_norm_sum, vis_counts, max_2d


# ------------------------------------------------------------------------------------------------------------------
# Pipelined trainer: no host sync on the path, two CUDA graphs per camera signature, the gradient exchange and the SH
# update of step k hidden behind the geometry of image k+1.
# ------------------------------------------------------------------------------------------------------------------

def geometry_phase(model: FlatGaussians, st: Dict, scene: Dict, capacity: int, status: torch.Tensor):
    """Phase A of a step -- everything that depends only on the GEOMETRY parameters (means, scales, quaternions,
    opacities, camera rows): the projection block of splatfacto.py:816-834 on the public operator, the opacity
    activation (:853-856) and the tile lists (gsplat.rasterize.prepare_lists: capacity mode, no host sync).
    `st` holds the step's camera as device tensors: cam (21 floats: viewmat 12 | lin_vel 3 | ang_vel 3 | cam_pos 3)
    and cam_index (int64[1])."""
    from gsplat import project_gaussians
    from gsplat.rasterize import prepare_lists

    p = model.params
    cam = st["cam"]
    viewmat, vel0, cam_pos = cam[:12].view(3, 4), cam[12:18], cam[18:21]
    if model.cam_vel is not None:  # per-camera velocity rows (disjoint across images): row cam_index + dataset value
        vel = model.cam_vel.index_select(0, st["cam_index"])[0] + vel0
    else:
        vel = vel0
    lin, ang = vel[:3].unsqueeze(0), vel[3:].unsqueeze(0)
    H, W, bw = scene["H"], scene["W"], scene["block_width"]
    rs, ex = scene["rolling_shutter_time"], scene["exposure_time"]
    quats = p["quats"] / p["quats"].norm(dim=-1, keepdim=True)
    xys, depths, pix_vels, radii, conics, comp, num_tiles_hit, _ = project_gaussians(
        p["means"], torch.exp(p["log_scales"]), 1, quats, lin, ang, rs, ex, viewmat, scene["fx"], scene["fy"], scene["cx"],
        scene["cy"], H, W, bw)
    opacities = torch.sigmoid(p["opacity_logit"]) * comp[:, None]  # "antialiased" mode, splatfacto.py:853-854
    blur = scene["blur_samples"] if ex > 0 else 1
    prep = prepare_lists(xys, depths, pix_vels, radii, conics, num_tiles_hit, opacities, H, W, bw, rs, ex, blur,
                         capacity=capacity, status=status)
    return dict(xys=xys, depths=depths, pix_vels=pix_vels, radii=radii, conics=conics, num_tiles_hit=num_tiles_hit,
                opacities=opacities, prep=prep, cam_pos=cam_pos, blur=blur)


def shading_phase(model: FlatGaussians, geo: Dict, scene: Dict, target: torch.Tensor, loss_fn, sh_degree_to_use: int = 3):
    """Phase B -- needs the SH parameters: colours (splatfacto.py:840-852), blend on the prepared lists, loss, and the
    backward pass through BOTH phases.  With the block layout the SH gradient goes straight into the flat gradient
    buffer (gsplat.sh.coeff_grad_sink).  Returns the (device) loss."""
    from gsplat import rasterize_gaussians, spherical_harmonics
    from gsplat.sh import coeff_grad_sink

    p = model.params
    H, W, bw = scene["H"], scene["W"], scene["block_width"]
    viewdirs = p["means"].detach() - geo["cam_pos"]
    sink = model.flat_grad[model.sh_start:model.sh_start + model.N * model.K * 3] if model.sh_layout == "block" else None
    with coeff_grad_sink(sink):
        rgbs = torch.clamp(spherical_harmonics(sh_degree_to_use, viewdirs, model.sh_coeffs()) + 0.5, min=0.0)
    rgb, alpha = rasterize_gaussians(geo["xys"], geo["depths"], geo["pix_vels"], geo["radii"], geo["conics"],
                                     geo["num_tiles_hit"], rgbs, geo["opacities"], H, W, bw,
                                     rolling_shutter_time=scene["rolling_shutter_time"], exposure_time=scene["exposure_time"],
                                     blur_samples=geo["blur"], background=scene["background"], return_alpha=True,
                                     prepared=geo["prep"])
    loss = loss_fn(rgb, target)
    loss.backward()
    return loss.detach()


def fused_geometry_phase(model: FlatGaussians, st: Dict, scene: Dict, capacity: int, status: torch.Tensor):
    """Phase A straight on the C ABI, raw parameters in: ONE kernel (b200_fused_geometry_forward: exp, quaternion
    normalisation, projection, sigmoid * compensation, blend record -- what `geometry_phase` spends 11 launches on, 8 of
    them PyTorch glue) + the capacity-mode binning.  Same interface and the same numbers as `geometry_phase` up to
    float rounding of the activations.  No autograd graph: `fused_shading_phase` differentiates with
    b200_fused_preprocess_backward."""
    import gsplat.cuda as _C
    from gsplat import _lib
    from gsplat._lib import check, ptr, stream

    p = model.params
    cam = st["cam"]
    with torch.no_grad():
        if model.cam_vel is not None:  # per-camera velocity rows (disjoint across images): row cam_index + dataset value
            vel = model.cam_vel.detach().index_select(0, st["cam_index"])[0] + cam[12:18]
        else:
            vel = cam[12:18]
        H, W, bw = scene["H"], scene["W"], scene["block_width"]
        rs, ex = float(scene["rolling_shutter_time"]), float(scene["exposure_time"])
        blur = scene["blur_samples"] if ex > 0 else 1
        n, dev = model.N, model.flat.device
        lib = _lib.load()
        with _lib.on_device(dev):
            packed = torch.empty((n * lib.b200_packed_record_bytes(),), dtype=torch.uint8, device=dev)
            depths = torch.empty((n,), dtype=torch.float32, device=dev)
            radii = torch.empty((n,), dtype=torch.int32, device=dev)
            nth = torch.empty((n,), dtype=torch.int32, device=dev)
            check(lib.b200_fused_geometry_forward(
                n, ptr(p["means"]), ptr(p["log_scales"]), ptr(p["quats"]), ptr(p["opacity_logit"]), ptr(cam), ptr(vel[:3]),
                ptr(vel[3:]), rs, ex, float(scene["fx"]), float(scene["fy"]), float(scene["cx"]), float(scene["cy"]), H, W, bw,
                _CLIP_THRESH, ptr(packed), ptr(depths), ptr(radii), ptr(nth), stream()))
            ids, bins = _C.bin_cull_capacity(packed, depths, radii, nth, H, W, bw, blur, rs, ex, capacity, status)
    return dict(packed=packed, depths=depths, radii=radii, num_tiles_hit=nth, ids=ids, bins=bins, vel=vel, cam=cam,
                cam_index=st["cam_index"], status=status, blur=blur, absgrad=None)


def fused_shading_phase(model: FlatGaussians, geo: Dict, scene: Dict, target: torch.Tensor, loss_fn, sh_degree_to_use: int = 3):
    """Phase B on the C ABI: SH colours into the records (b200_fused_colors_forward), blend, `loss_fn` (autograd sees
    only loss_fn(rgb, target)), blend backward, and ONE kernel for everything behind it (b200_fused_preprocess_backward:
    SH, clamp, sigmoid * compensation, projection VJP, exp, normalisation) that writes the model's flat gradient buffer
    row by row -- no accumulate passes, no memsets.  The gradient rows are OVERWRITTEN (the buffer is zero between steps:
    FlatAdam clears what it consumes); the camera-velocity row is accumulated."""
    import gsplat.cuda as _C
    from gsplat import _lib
    from gsplat._lib import check, ptr, stream

    if model.sh_layout != "split":
        raise ValueError("fused_shading_phase reads sh_dc / sh_rest apart: build the model with sh_layout='split'")
    p = model.params
    H, W, bw = scene["H"], scene["W"], scene["block_width"]
    rs, ex = float(scene["rolling_shutter_time"]), float(scene["exposure_time"])
    fx, fy, cx, cy = (float(scene[k]) for k in ("fx", "fy", "cx", "cy"))
    n, K, S, dev = model.N, model.K, geo["blur"], model.flat.device
    cam, vel, packed, radii = geo["cam"], geo["vel"], geo["packed"], geo["radii"]
    cam_pos = cam[18:21]
    bg = scene["background"]
    lib = _lib.load()
    with torch.no_grad(), _lib.on_device(dev):
        check(lib.b200_fused_colors_forward(n, ptr(p["means"]), ptr(p["sh_dc"]), ptr(p["sh_rest"]), K, sh_degree_to_use,
                                            ptr(cam_pos), ptr(radii), ptr(packed), stream()))
        rgb, Ts, fi = _C.blend_forward_packed(H, W, bw, S, geo["ids"], geo["bins"], packed, rs, ex, bg, status=geo["status"])
    from gsplat.losses import l1_loss, l1_loss_and_grad
    if loss_fn is l1_loss:  # the default: value and cotangent from the one kernel
        loss, v_rgb = l1_loss_and_grad(rgb, target)
    else:
        rgb.requires_grad_(True)
        with torch.enable_grad():
            loss = loss_fn(rgb, target)
        (v_rgb,) = torch.autograd.grad(loss, rgb)
    with torch.no_grad(), _lib.on_device(dev):
        v_xy, v_abs, v_pix, v_conic, v_col, v_op = _C.blend_backward_packed(
            n, H, W, bw, S, geo["ids"], geo["bins"], packed, rs, ex, bg, Ts, fi, v_rgb, None)
        g = {name: model.flat_grad[a:b_] for name, (a, b_) in model.slices.items()}
        g_vel = torch.empty(6, dtype=torch.float32, device=dev) if model.cam_vel is not None else None
        check(lib.b200_fused_preprocess_backward(
            n, ptr(p["means"]), ptr(p["log_scales"]), ptr(p["quats"]), ptr(p["opacity_logit"]), ptr(p["sh_dc"]),
            ptr(p["sh_rest"]), K, sh_degree_to_use, ptr(cam), ptr(cam_pos), ptr(vel[:3]), ptr(vel[3:]), rs, ex, fx, fy, cx, cy,
            H, W, bw, _CLIP_THRESH, ptr(packed), ptr(radii), ptr(v_xy), ptr(v_pix), ptr(v_conic), ptr(v_col), ptr(v_op),
            ptr(g["means"]), ptr(g["log_scales"]), ptr(g["quats"]), ptr(g["opacity_logit"]), ptr(g["sh_dc"]), ptr(g["sh_rest"]),
            ptr(g_vel[:3]) if g_vel is not None else None, ptr(g_vel[3:]) if g_vel is not None else None, None, stream()))
        if g_vel is not None:
            g["cam_vel"].view(-1, 6).index_add_(0, geo["cam_index"], g_vel.view(1, 6))
    geo["absgrad"] = v_abs   # the `xys.absgrad` side channel of the drop-in operator (densification statistic)
    return loss.detach()


_CLIP_THRESH = 0.01  # project_gaussians' default near plane (gsplat/gsplat/project_gaussians.py:31)


def balanced_assignment(costs, world: int):
    """Cost-aware batching for synchronous data parallelism.  A step lasts as long as its slowest rank, and images differ
    in cost (tile-list length: +-20 % at BASELINE config 2), so a random group of `world` images wastes the difference to
    the group's maximum on every other rank.  costs: one number per image (e.g. the list entries phase A reports in
    status[1]); returns steps[j][r] = index of the image rank r renders in step j, every image used exactly once per
    pass, each step made of images adjacent in the cost order (length-bucketing, as sequence models do).  The step ORDER
    is then shuffled deterministically so consecutive steps do not walk the cost ramp."""
    n = len(costs)
    if n % world:
        raise ValueError("balanced_assignment: the number of images must be a multiple of the world size")
    order = sorted(range(n), key=lambda i: (-float(costs[i]), i))
    steps = [order[j * world:(j + 1) * world] for j in range(n // world)]
    # a fixed permutation of the steps: stride through them with a step count coprime to their number
    m = len(steps)
    stride = next((s_ for s_ in range(max(2, m // 2) | 1, 2 * m + 3, 2) if math.gcd(s_, m) == 1), 1) if m > 2 else 1
    return [steps[(k * stride) % m] for k in range(m)]


class PipelinedTrainer:
    """Image-sharded trainer on the drop-in operators with nothing on the path waiting for the host.

    A step is cut where the parameter dependencies cut it:
      A(k)  geometry of image k: projection + opacity + tile lists   -- reads the geometry rows only
      B(k)  SH colours, blend, loss, backward through A and B         -- reads the SH block too
    and the host queues, per step k:  B(k) | exchange(geometry grads), exchange(SH grads) | Adam(geometry) | A(k+1) ||
    Adam(SH) on a side stream.  The SH block is 48 of the 59 floats per Gaussian: its allreduce and its update run behind
    the projection + binning of the NEXT image, which only needs the geometry rows -- exact, because the SH parameters
    are first read by B(k+1) (which waits for the side stream).  On one GPU the same ordering overlaps the bandwidth-bound
    SH update with the latency-bound binning.

    No host sync: the tile lists are sized from a running high-water mark (gsplat.rasterize.prepare_lists); if an image
    needs more, its lists are incomplete -- the device raises a flag, the (rank-reduced) flag VETOES that step's
    optimizer update on the device (gsplat.optim.FlatAdam device state), and the host, which polls a pinned copy of the
    status words one step late, grows the capacity and reports the step in `vetoed` so the caller repeats the image.
    With `use_graphs` A and B are captured as CUDA graphs per camera signature (image size, intrinsics, blur settings,
    SH degree, capacity): the ~60 launches of a step become two graph launches; per-step inputs (camera, target) live in
    static device buffers.

    API:  prepare(cam, cam_index)            -- A for the first image (and after anything that invalidates it)
          train_step(target, next_cam=None, next_index=0) -> device loss
    `cam` = dict(viewmat (3,4), lin_vel (3), ang_vel (3), cam_pos (3)) tensors on any device, or `cam_floats`
    (21 floats in that order, e.g. a pinned host row from gsplat.data.ImagePrefetcher)."""

    def __init__(self, model: FlatGaussians, scene: Dict, lr: float = 1e-3, group=None, loss_fn=None, use_graphs: bool = True,
                 capacity: Optional[int] = None, sh_degree_to_use: int = 3, geometry_fn=None, shading_fn=None,
                 optimizer: str = "b200", sh_chunks: int = 1, operators: Optional[str] = None):
        """operators: "dropin" = the two phases on the reference-compatible operators under autograd (`geometry_phase` /
        `shading_phase`); "fused" = on the fused raw-parameter kernels (`fused_geometry_phase` / `fused_shading_phase`:
        same numbers, ~45 fewer launches per step); None = fused where it applies (CUDA, sh_layout "split"), else drop-in."""
        self.model, self.scene = model, dict(scene)
        if loss_fn is None:
            from gsplat.losses import l1_loss as loss_fn
        self.loss_fn = loss_fn
        self.sh_degree = sh_degree_to_use
        fusable = model.flat.device.type == "cuda" and model.sh_layout == "split"
        if operators is None:
            operators = "fused" if fusable else "dropin"
        if operators not in ("fused", "dropin"):
            raise ValueError("operators must be 'fused', 'dropin' or None")
        if operators == "fused" and not fusable and geometry_fn is None:
            raise ValueError("operators='fused' needs a CUDA model with sh_layout='split'")
        self.operators = operators
        phases = (fused_geometry_phase, fused_shading_phase) if operators == "fused" else (geometry_phase, shading_phase)
        self.geometry_fn = geometry_fn or phases[0]   # (tests inject CPU stand-ins with the same interface)
        self.shading_fn = shading_fn or phases[1]
        self.distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.group = group
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.world = dist.get_world_size(group) if self.distributed else 1
        dev = model.flat.device
        self.device = dev
        self.cuda = dev.type == "cuda"
        self.use_graphs = bool(use_graphs and self.cuda)
        self._avg = self.distributed and dist.get_backend(group) == "nccl"
        self._op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        # static per-step inputs / outputs
        self.st = dict(cam=torch.zeros(21, dtype=torch.float32, device=dev), cam_index=torch.zeros(1, dtype=torch.int64, device=dev))
        self.status = torch.zeros(4, dtype=torch.int32, device=dev)       # [overflow, entries, max entries, reference isects]
        self.flag = self.status[0:1]                                      # this step's veto (reduced with MAX across ranks)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.capacity = capacity
        self.vetoed: List[int] = []     # step indices whose update was skipped (the caller repeats those images)
        self.after_backward = None      # optional callable(absgrad (N,2), radii (N,)) run between phase B and the next phase A
        self.graph_kernel_launches = 0  # libb200splat kernels replayed from captured graphs (the C launch counter only
        #                                 sees launches made through the C ABI, i.e. eager ones and those DURING capture)
        self.steps = 0
        self._geo = None
        self._graphs: Dict[tuple, Dict] = {}
        self._flag_work = None
        self._pending_sh = None
        self.optimizer = optimizer
        # SH update launch shape (measured at c2, 1 GPU, r2q: resident-CTA-limited background kernels were SLOWER -- 1.33 ms
        # per step with 1-2 CTAs per SM against 1.30 with one short CTA per 256 vectors -- so 0 = short CTAs is the default)
        self.sh_update_ctas = int(os.environ.get("B200_SH_UPDATE_CTAS", "0"))
        self.sh_update_pieces = int(os.environ.get("B200_SH_UPDATE_PIECES", "8"))
        # the SH block is exchanged in `sh_chunks` pieces so that the update of piece i runs while piece i+1 is on the wire
        lo, total = model.sh_start, model.flat.numel()
        n_sh = max(1, int(sh_chunks)) if self.distributed else 1
        cuts = [lo + ((total - lo) * k // n_sh) // 4 * 4 for k in range(n_sh)] + [total]
        self._sh_bounds = [(cuts[k], cuts[k + 1]) for k in range(n_sh) if cuts[k + 1] > cuts[k]]
        if optimizer == "b200":
            from gsplat.optim import FlatAdam
            self.adam = FlatAdam(model.flat, model.flat_grad, lr=lr, eps=1e-15).use_device_state()
        elif optimizer == "torch":  # CPU / gloo tests of this host logic: same slices, torch's Adam, host-side veto
            self.adam = None
            lo, total = model.sh_start, model.flat.numel()
            self._torch_opts = []
            for a, b_ in ((0, lo), (lo, total)):
                alias = model.flat[a:b_].detach().requires_grad_(True)
                alias.grad = model.flat_grad[a:b_]
                self._torch_opts.append(torch.optim.Adam([alias], lr=lr, eps=1e-15))
        else:
            raise ValueError("optimizer must be 'b200' or 'torch'")
        if self.cuda:
            # Everything this trainer queues runs on its own stream `main` (callers on another stream are ordered before /
            # after each call).  Autograd ties every leaf's gradient accumulation to the stream that was current when the
            # accumulator was first used; if that were the legacy default stream, the backward pass inside a graph capture
            # would have to synchronise with it, which invalidates the capture.
            # `main` outranks `side`: the SH update (one bandwidth-bound kernel of many short CTAs) then fills the SMs the
            # next image's projection / binning kernels leave idle instead of queueing its whole grid ahead of them
            self.main = torch.cuda.Stream(device=dev, priority=-1)
            self.side = torch.cuda.Stream(device=dev, priority=0)
            model.rebind_leaves()  # gradient accumulators are (re)created on `main` by the first step
            self.ev_prepared, self.ev_sh_done, self.ev_poll = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
            self._host_words = torch.zeros(2, 24, dtype=torch.int32).pin_memory()   # status (4) | adam state (16) | quat flag
            self._host_ev = [torch.cuda.Event(), torch.cuda.Event()]
            self._host_seen = [False, False]
            self._vetoed_seen = 0

    # ---- camera / capacity --------------------------------------------------------------------------------------
    def _set_camera(self, cam, cam_index):
        if torch.is_tensor(cam):  # 21 floats: viewmat 12 | lin_vel 3 | ang_vel 3 | cam_pos 3 (host pinned or device)
            row = cam.reshape(-1)[:21]
        else:
            d0 = cam["viewmat"].device
            row = torch.cat([cam["viewmat"].reshape(-1)[:12], cam["lin_vel"].reshape(-1)[:3].to(d0),
                             cam["ang_vel"].reshape(-1)[:3].to(d0), cam["cam_pos"].reshape(-1)[:3].to(d0)]).float()
            for k in ("fx", "fy", "cx", "cy"):  # intrinsics are launch constants (part of the graph key)
                if k in cam:
                    self.scene[k] = float(cam[k])
        self.st["cam"].copy_(row, non_blocking=True)
        self.st["cam_index"].fill_(int(cam_index))

    def _size_capacity(self):
        """First use: one synchronous probe of the entry count with a generous capacity; later growth comes from the
        polled high-water mark."""
        if self.capacity is not None:
            return
        if not self.cuda:
            self.capacity = 1
            return
        probe = max(1 << 20, 8 * self.model.N)
        with torch.no_grad():
            self.geometry_fn(self.model, self.st, self.scene, probe, self.status)
        torch.cuda.current_stream().synchronize()
        entries = int(self.status[1])
        if int(self.status[0]):
            entries = max(entries, probe)
        self.status.zero_()
        self.capacity = self._round_capacity(int(1.6 * entries) + 65536)

    @staticmethod
    def _round_capacity(n):
        return (int(n) + 65535) // 65536 * 65536

    # ---- phases (eager or captured) ------------------------------------------------------------------------------
    def _key(self):
        s = self.scene
        return (s["H"], s["W"], s["fx"], s["fy"], s["cx"], s["cy"], s["block_width"], s["blur_samples"], s["exposure_time"],
                s["rolling_shutter_time"], self.sh_degree, self.capacity, self.model.N)

    def _entry(self, target_shape_dtype):
        key = self._key() + (target_shape_dtype,)
        e = self._graphs.get(key)
        if e is None:
            e = self._graphs[key] = dict(gA=None, gB=None, geo=None, target=None, warm=0)
        return e

    def _lib_launches(self):
        from gsplat import _lib
        return _lib.load().b200_launch_count() if self.cuda else 0

    def _run_A(self, e):
        if self.use_graphs and e["gA"] is not None:
            e["gA"].replay()
            self.graph_kernel_launches += e["nA"]
            self._geo = e["geo"]
            return
        if self.use_graphs and e["warm"] >= 2:  # two eager warm-up rounds (allocator, lazy inits), then capture
            g = torch.cuda.CUDAGraph()
            n0 = self._lib_launches()
            with torch.cuda.graph(g, stream=self.main):
                e["geo"] = self.geometry_fn(self.model, self.st, self.scene, self.capacity, self.status)
            e["gA"], e["nA"] = g, self._lib_launches() - n0   # libb200splat kernels recorded in the graph
            g.replay()  # (capture does not execute)
            self._geo = e["geo"]
            return
        self._geo = self.geometry_fn(self.model, self.st, self.scene, self.capacity, self.status)
        e["geo_eager"] = True

    def _run_B(self, e, target):
        if self.use_graphs and e["gB"] is not None:
            e["target"].copy_(target, non_blocking=True)
            e["gB"].replay()
            self.graph_kernel_launches += e["nB"]
            return
        if self.use_graphs and e["gA"] is not None and e["warm"] >= 2:
            # B is captured against the tensors A's graph writes (static addresses), with a static target buffer
            e["target"] = torch.empty_like(target)
            e["target"].copy_(target)
            g = torch.cuda.CUDAGraph()
            n0 = self._lib_launches()
            with torch.cuda.graph(g, stream=self.main):
                tgt = self._as_target(e["target"])
                self.loss.copy_(self.shading_fn(self.model, e["geo"], self.scene, tgt, self.loss_fn, self.sh_degree))
            e["gB"], e["nB"] = g, self._lib_launches() - n0
            g.replay()
            return
        tgt = self._as_target(target)
        self.loss.copy_(self.shading_fn(self.model, self._geo, self.scene, tgt, self.loss_fn, self.sh_degree))
        e["warm"] += 1

    def _as_target(self, target):
        """The dataset's uint8 image goes to the loss as it is when the loss converts it in its own kernel
        (gsplat.losses.l1_loss / photometric_loss); any other loss gets the float image the reference builds
        (get_gt_img, splatfacto.py:906-907)."""
        if target.dtype != torch.uint8 or getattr(self.loss_fn, "accepts_uint8", False):
            return target
        return target.float() / 255

    # ---- public API ----------------------------------------------------------------------------------------------
    class _OnMain:
        """Run a block on the trainer's stream, ordered after the caller's current stream and before its later work."""

        def __init__(self, tr):
            self.tr, self.ctx, self.cur = tr, None, None

        def __enter__(self):
            tr = self.tr
            if not tr.cuda:
                return
            self.cur = torch.cuda.current_stream(tr.device)
            if self.cur != tr.main:
                tr.main.wait_stream(self.cur)
                self.ctx = torch.cuda.stream(tr.main)
                self.ctx.__enter__()

        def __exit__(self, *exc):
            if self.ctx is not None:
                self.ctx.__exit__(*exc)
                self.cur.wait_stream(self.tr.main)
            return False

    def prepare(self, cam, cam_index: int = 0):
        """Stage the camera of the image the next train_step trains on (phase A itself is queued by train_step: at once
        for the first image, behind the previous step's geometry update afterwards)."""
        with self._OnMain(self):
            self._prepare(cam, cam_index)

    def _prepare(self, cam, cam_index):
        self._set_camera(cam, cam_index)
        self._size_capacity()
        self._prepared = True
        self._A_done = False

    def _ensure_A(self, e):
        if not self._A_done:
            self._run_A(e)
            self._A_done = True
            if self.distributed:  # this image's overflow flag, agreed on by all ranks before anyone's Adam reads it
                self._flag_work = dist.all_reduce(self.flag, op=dist.ReduceOp.MAX, group=self.group, async_op=True)

    def train_step(self, target: torch.Tensor, next_cam=None, next_index: int = 0):
        """Trains on the prepared image (B, exchange, update) and, given `next_cam`, runs A for the next image while the
        SH block is still being exchanged / updated.  Returns the device loss (a static tensor: copy it if you keep it)."""
        assert getattr(self, "_prepared", False), "call prepare(cam, cam_index) before the first train_step"
        with self._OnMain(self):
            return self._train_step(target, next_cam, next_index)

    def _train_step(self, target, next_cam, next_index):
        m = self.model
        e = self._entry((tuple(target.shape), target.dtype))
        self._ensure_A(e)
        if self.cuda and self._pending_sh is not None:
            torch.cuda.current_stream().wait_event(self.ev_sh_done)   # SH parameters of the previous step are final
            self._pending_sh = None
        self._run_B(e, target)
        if self.after_backward is not None:
            # e.g. gsplat.densify.Densifier.accumulate: this image's radii and |d loss / d xy| (the `xys.absgrad` side channel)
            # are both valid exactly here -- phase A of the next image overwrites the projection outputs below
            absgrad = self._geo["absgrad"] if "absgrad" in self._geo else self._geo["xys"].absgrad
            self.after_backward(absgrad, self._geo["radii"])
        lo, total = m.sh_start, m.flat.numel()
        works = None
        if self.distributed:
            works = [dist.all_reduce(m.flat_grad[:lo], op=self._op, group=self.group, async_op=True)]
            works += [dist.all_reduce(m.flat_grad[a:b_], op=self._op, group=self.group, async_op=True) for a, b_ in self._sh_bounds]
            self._flag_work.wait()
        scale = 1.0 if (self._avg or not self.distributed) else 1.0 / self.world
        veto_host = False
        if self.adam is not None:
            self.adam.prepare(self.flag)
            if self.cuda:
                self.ev_prepared.record()
        else:  # torch optimizer (CPU tests): the veto is read on the host
            veto_host = bool(int(self.flag[0]) != 0)
            self.flag.zero_()
            if veto_host:
                self.vetoed.append(self.steps)
        # the next image's camera row is staged while the geometry gradients are on the wire (B, its last reader, is queued)
        self._prepared = False
        if next_cam is not None:
            self._prepare(next_cam, next_index)
        if works is not None:
            works[0].wait()
        self._update(0, lo, scale, veto_host, 0)
        self._poll_host()
        # ---- next image's geometry (needs only the rows just updated), SH update beside it
        if self.cuda:
            with torch.cuda.stream(self.side):
                self.side.wait_event(self.ev_prepared)
                for j, (a, b_) in enumerate(self._sh_bounds):
                    if works is not None:
                        works[1 + j].wait()
                    self._update(a, b_, scale, veto_host, 1)
                self.ev_sh_done.record(self.side)
            self._pending_sh = True
            if next_cam is not None:
                self._ensure_A(self._entry((tuple(target.shape), target.dtype)))
        else:
            for j in range(len(self._sh_bounds)):
                if works is not None:
                    works[1 + j].wait()
            self._update(lo, total, scale, veto_host, 1)
        self.steps += 1
        return self.loss

    def _update(self, a, b_, scale, veto_host, which):
        if self.adam is not None:
            if which == 1 and self.sh_update_pieces > 1:
                # the SH slice is updated on the side stream beside the next image's projection / binning.  One launch would
                # queue its whole grid (tens of thousands of short CTAs) ahead of whatever arrives later at the same
                # priority -- the memset nodes of the binning graph carry none, r2p timeline: the depth sort waited 31 us
                # for one -- so it goes out in pieces: a late arrival waits for one piece at most
                n = self.sh_update_pieces
                cuts = [a + ((b_ - a) * k // n) // 4 * 4 for k in range(n)] + [b_]
                for lo_, hi_ in zip(cuts[:-1], cuts[1:]):
                    if hi_ > lo_:
                        self.adam.update_state(lo_, hi_, scale, True, background_ctas=self.sh_update_ctas)
                return
            self.adam.update_state(a, b_, scale, True, background_ctas=self.sh_update_ctas if which == 1 else 0)
            return
        g = self.model.flat_grad[a:b_]
        if not veto_host:
            if scale != 1.0:
                g.mul_(scale)
            self._torch_opts[which].step()
        g.zero_()

    def on_resize(self):
        """The model's buffers were reallocated (gsplat.densify.Densifier.refine): drop the captured graphs (they address
        the old buffers), recompute the exchange slices, put the new leaves on this trainer's stream.  The next step
        re-sizes nothing else: the list capacity follows the polled high-water mark as before."""
        self.finish()
        self._graphs.clear()
        self._geo = None
        self._prepared = False
        m = self.model
        lo, total = m.sh_start, m.flat.numel()
        n_sh = len(self._sh_bounds)
        cuts = [lo + ((total - lo) * k // n_sh) // 4 * 4 for k in range(n_sh)] + [total]
        self._sh_bounds = [(cuts[k], cuts[k + 1]) for k in range(n_sh) if cuts[k + 1] > cuts[k]]
        if self.cuda:
            with torch.cuda.stream(self.main):
                m.rebind_leaves()

    def reserve(self, entries: int):
        """Make the tile lists hold at least `entries` (+ 30 % headroom) from now on.  A caller that knows its images' entry
        counts (a measuring pass, a previous epoch) calls this once -- on every rank with the SAME number -- so that the
        lists never grow mid-training: a growth means new CUDA graphs (two eager steps, then ~10 ms of capture), and in a
        synchronous data-parallel step one rank's capture stalls all of them."""
        want = self._round_capacity(int(1.3 * int(entries)) + 65536)
        if self.capacity is None or want > self.capacity:
            self.capacity = want

    def steady(self, target_shape_dtype) -> bool:
        """True when the next step of this camera signature replays captured graphs (or graphs are off): nothing left to
        warm up.  `target_shape_dtype` = (tuple(target.shape), target.dtype) of the images passed to train_step."""
        if not (self.use_graphs and self.cuda):
            return True
        e = self._graphs.get(self._key() + (target_shape_dtype,))
        return e is not None and e["gA"] is not None and e["gB"] is not None

    def finish(self):
        """Order the caller's stream after everything queued so far, the SH update on the side stream included (call
        before reading parameters)."""
        with self._OnMain(self):
            if self.cuda and self._pending_sh is not None:
                torch.cuda.current_stream().wait_event(self.ev_sh_done)
                self._pending_sh = None

    # ---- lagged host view of the device status (overflow / vetoes / input check) ---------------------------------
    def _poll_host(self):
        if not self.cuda or self.adam is None:
            return
        slot = self.steps & 1
        other = 1 - slot
        if self._host_seen[other] and self._host_ev[other].query():
            self._consume(self._host_words[other])
            self._host_seen[other] = False
        w = self._host_words[slot]
        # the three small transfers go on the side stream (idle here: the previous SH update finished before phase B), so
        # the next image's geometry does not queue behind them; the words may then mix this step's counters with the next
        # image's (phase A runs beside the copies) -- every consumer below is monotone (running max, veto count + ring)
        self.ev_poll.record()
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ev_poll)
            w[0:4].copy_(self.status, non_blocking=True)
            w[4:20].copy_(self.adam.state, non_blocking=True)
            from gsplat import _lib
            qf = _lib._quat_flags.get(self.device.index)
            if qf is not None:
                w[20:21].copy_(qf, non_blocking=True)
            self._host_ev[slot].record(self.side)
        self._host_seen[slot] = True

    def _consume(self, w):
        words = w.tolist()
        if words[20]:
            from gsplat import _lib
            _lib.raise_if_flagged(words[20])  # deferred project_gaussians.py:69
        vetoed = words[4 + 5]
        if vetoed > self._vetoed_seen:
            ring = words[4 + 6:4 + 16]
            for j in range(self._vetoed_seen, vetoed):
                if vetoed - j <= 10:
                    self.vetoed.append(ring[j % 10])
            self._vetoed_seen = vetoed
            self.capacity = self._round_capacity(max(self.capacity, int(1.3 * words[2]) + 65536))  # new graphs at the new size
        elif words[2] > 0.9 * self.capacity:  # growing scene: enlarge before it overflows
            self.capacity = self._round_capacity(int(1.3 * words[2]) + 65536)

    def sync_status(self):
        """Blocking variant of the poll (end of training / tests): returns dict(entries_max, capacity, vetoed)."""
        if self.cuda and self.adam is not None:
            torch.cuda.synchronize(self.device)
            for slot in (0, 1):
                if self._host_seen[slot]:
                    self._consume(self._host_words[slot])
                    self._host_seen[slot] = False
            self._consume(torch.cat([self.status, self.adam.state, torch.zeros(4, dtype=torch.int32, device=self.device)]).cpu())
        return dict(entries_max=int(self.status[2]), capacity=self.capacity, vetoed=list(self.vetoed), steps=self.steps)
