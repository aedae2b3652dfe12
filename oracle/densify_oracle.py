"""CPU restatement of the reference's adaptive density control.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows nerfstudio/models/splatfacto.py: after_train :408-434, refinement_after :443-531, cull_gaussians :533-566,
split_gaussians :568-611, dup_gaussians :613-622, optimizer surgery :352-406 -- on plain tensors (a dict of the six
Gaussian parameter tensors + a dict of their Adam moments), float32 like the reference.  Pinned against the reference
itself by tests/test_densify_cpu.py, which runs the real SplatfactoModel.refinement_after here (CPU) on the same inputs."""
import math

import torch


def accumulate(stats, absgrad, radii, H, W):
    """after_train: stats = dict(grad_norm, vis_counts, max_2d) or {} before the first image."""
    visible = radii > 0
    grads = absgrad.norm(dim=-1)
    if not stats:
        stats["grad_norm"] = grads.clone()
        stats["vis_counts"] = torch.ones_like(grads)
        stats["max_2d"] = torch.zeros_like(grads)
    else:
        stats["vis_counts"][visible] += 1
        stats["grad_norm"][visible] += grads[visible]
    stats["max_2d"][visible] = torch.maximum(stats["max_2d"][visible], radii[visible].float() / float(max(H, W)))
    return stats


def _quat_to_rotmat(q):
    w, x, y, z = torch.unbind(q / q.norm(dim=-1, keepdim=True), dim=-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                        2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                        2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)


def refine(params, moments, stats, cfg, step, num_train_data, last_size, z=None):
    """params: dict(means, log_scales, quats, opacity_logit, + any further per-Gaussian tensors); moments: dict name ->
    (exp_avg, exp_avg_sq) or None; z: (n_split_samples * n_splits, 3) normal samples (drawn with torch.randn if None).
    Returns (params, moments, info) -- new tensors, inputs untouched."""
    if step <= cfg.warmup_length:
        return params, moments, None
    reset_interval = cfg.reset_alpha_every * cfg.refine_every
    do_densify = step < cfg.stop_split_at and step % reset_interval > num_train_data + cfg.refine_every
    params = {k: v.clone() for k, v in params.items()}
    moments = None if moments is None else {k: (a.clone(), b.clone()) for k, (a, b) in moments.items()}
    n = params["means"].shape[0]
    max_2d = stats.get("max_2d") if stats else None
    info = None

    def cull_mask(p, max2d, extra=None):
        culls = torch.sigmoid(p["opacity_logit"]).squeeze(-1) < cfg.cull_alpha_thresh
        if extra is not None:
            culls = culls | extra
        if step > cfg.refine_every * cfg.reset_alpha_every:
            toobigs = torch.exp(p["log_scales"]).max(dim=-1).values > cfg.cull_scale_thresh
            if step < cfg.stop_screen_size_at:
                toobigs = toobigs | (max2d > cfg.cull_screen_size)
            culls = culls | toobigs
        return culls

    deleted = None
    if do_densify:
        avg = (stats["grad_norm"] / stats["vis_counts"]) * 0.5 * max(last_size[0], last_size[1])
        high = avg > cfg.densify_grad_thresh
        smax = params["log_scales"].exp().max(dim=-1).values
        splits = smax > cfg.densify_size_thresh
        if step < cfg.stop_screen_size_at:
            splits = splits | (max_2d > cfg.split_screen_size)
        splits = splits & high
        samps = cfg.n_split_samples
        n_splits = int(splits.sum())
        if z is None:
            z = torch.randn((samps * n_splits, 3))
        scaled = torch.exp(params["log_scales"][splits].repeat(samps, 1)) * z
        rots = _quat_to_rotmat(params["quats"][splits].repeat(samps, 1))
        split_new = {k: v[splits].repeat(samps, *([1] * (v.dim() - 1))) for k, v in params.items()}
        split_new["means"] = torch.bmm(rots, scaled[..., None]).squeeze(-1) + params["means"][splits].repeat(samps, 1)
        split_new["log_scales"] = torch.log(torch.exp(params["log_scales"][splits]) / 1.6).repeat(samps, 1)
        # split_gaussians also shrinks the PARENTS in place (:597) before the duplicate mask is formed (:469): a split
        # parent whose shrunken size falls under the threshold is duplicated as well (with the shrunken scales)
        params["log_scales"][splits] = torch.log(torch.exp(params["log_scales"][splits]) / 1.6)
        dups = (params["log_scales"].exp().max(dim=-1).values <= cfg.densify_size_thresh) & high
        dup_new = {k: v[dups] for k, v in params.items()}
        params = {k: torch.cat([params[k], split_new[k], dup_new[k]], 0) for k in params}
        n_new = samps * n_splits + int(dups.sum())
        max_2d = torch.cat([max_2d, torch.zeros(n_new)])
        if moments is not None:
            moments = {k: tuple(torch.cat([t, torch.zeros((n_new,) + t.shape[1:])], 0) for t in ab) for k, ab in moments.items()}
        splits_mask = torch.cat([splits, torch.zeros(n_new, dtype=torch.bool)])
        deleted = cull_mask(params, max_2d, splits_mask)
        info = dict(step=step, before=n, splits=n_splits, dups=int(dups.sum()), densified=True)
    elif step >= cfg.stop_split_at and cfg.continue_cull_post_densification:
        deleted = cull_mask(params, max_2d)
        info = dict(step=step, before=n, splits=0, dups=0, densified=False)
    if deleted is not None:
        params = {k: v[~deleted] for k, v in params.items()}
        if moments is not None:
            moments = {k: tuple(t[~deleted] for t in ab) for k, ab in moments.items()}
        info["after"] = params["means"].shape[0]
    if step < cfg.stop_split_at and step % reset_interval == cfg.refine_every:
        lim = math.log((cfg.cull_alpha_thresh * 2.0) / (1.0 - cfg.cull_alpha_thresh * 2.0))
        params["opacity_logit"] = torch.clamp(params["opacity_logit"], max=lim)
        if moments is not None and "opacity_logit" in moments:
            moments["opacity_logit"] = tuple(torch.zeros_like(t) for t in moments["opacity_logit"])
        info = dict(info or {}, opacity_reset=True)
    return params, moments, info
