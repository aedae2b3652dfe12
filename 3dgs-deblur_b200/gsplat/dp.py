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
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

# flat layout: the geometry/opacity rows (11 floats per Gaussian) and the per-camera rows first, the SH coefficients
# (48 of the 59 floats at K = 16) last, so "everything but SH" and "SH" are each ONE contiguous slice to exchange
FIELDS = ("means", "log_scales", "quats", "opacity_logit", "sh_dc", "sh_rest")


class FlatGaussians:
    """Gaussian parameters as views into one flat buffer (+ a flat gradient buffer of the same layout)."""

    def __init__(self, scene: Dict, device, n_cameras: int = 0, optimize_velocities: bool = False):
        N = scene["means"].shape[0]
        K = scene["sh_rest"].shape[1] + 1
        self.N, self.K = N, K
        widths = dict(means=3, log_scales=3, quats=4, opacity_logit=1, sh_dc=3, sh_rest=3 * (K - 1))
        shapes = dict(means=(N, 3), log_scales=(N, 3), quats=(N, 4), opacity_logit=(N, 1), sh_dc=(N, 1, 3),
                      sh_rest=(N, K - 1, 3))
        extra = 6 * n_cameras if optimize_velocities else 0
        total = N * sum(widths.values()) + extra
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
            if name == "sh_dc" and extra:  # camera rows sit between the geometry block and the SH block
                self.cam_vel = take("cam_vel", off, extra, (n_cameras, 6))
                off += extra
            self.flat[off:off + N * widths[name]].view(shapes[name]).copy_(scene[name].to(device).reshape(shapes[name]))
            self.params[name] = take(name, off, N * widths[name], shapes[name])
            off += N * widths[name]
        assert off == total
        self.sh_start = self.slices["sh_dc"][0]  # flat[sh_start:] = all SH coefficients
        self.floats_per_gaussian = sum(widths.values())

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
    colors = torch.cat((p["sh_dc"], p["sh_rest"]), dim=1)
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
            for name in ("sh_dc", "sh_rest"):
                model.params[name].register_post_accumulate_grad_hook(self._on_sh_grad)

    def _reduce_async(self, chunk_ids):
        return [dist.all_reduce(self._chunks[i][0], op=self._op, group=self.group, async_op=True) for i in chunk_ids]

    def _on_sh_grad(self, _param):
        self._sh_seen += 1
        if self._sh_seen == 2:  # both halves of the cat() have landed in the flat buffer
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
