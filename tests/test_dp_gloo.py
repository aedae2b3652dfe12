"""CPU, world_size 2, gloo: host-side logic of the image-sharded dispatcher (gsplat/dp.py).

The render itself needs the GPU library, so it is replaced by a differentiable stand-in with the same parameter
interface; what is under test is the flat-buffer layout, the image -> rank assignment, the single gradient
allreduce, replica equality after the optimizer step and the densification-statistic reductions."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _torch_l1(pred, target):
    return (pred - target).abs().mean()


def _fake_render(model, cam, scene, cam_index=0, sh_degree_to_use=3):
    p = model.params
    w = cam["w"]
    rgb = (p["means"].sum() * w + p["sh_dc"].sum() + (p["sh_rest"] ** 2).sum() + torch.sigmoid(p["opacity_logit"]).sum()
           + p["log_scales"].exp().sum() + (p["quats"] / p["quats"].norm(dim=-1, keepdim=True)).sum())
    if model.cam_vel is not None:
        rgb = rgb + (model.cam_vel[cam_index] * w).sum()
    img = rgb * torch.ones(scene["H"], scene["W"], 3)
    return img, img[..., 0], p["means"][:, :2], torch.ones(p["means"].shape[0], dtype=torch.int32)


def _worker(rank, world, port, out, sh_chunks):
    sys.path.insert(0, os.path.join(ROOT, "3dgs-deblur_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gsplat import dp, synthetic

    dp.render = _fake_render
    scene = synthetic.make_scene("c1", n_override=50, n_cameras=4)
    model = dp.FlatGaussians(scene, "cpu", n_cameras=4, optimize_velocities=True)
    assert model.flat.numel() == 50 * 59 + 24 + (-(50 * 11 + 24)) % 4 and model.floats_per_gaussian == 59  # SH rows 16-byte aligned
    # parameter views alias the flat buffer, gradient views alias the flat gradient buffer
    model.params["means"].data[0, 0] = 7.0
    assert model.flat[0] == 7.0
    tr = dp.ImageShardedTrainer(model, scene, lr=1e-2, loss_fn=_torch_l1, optimizer="torch", sh_chunks=sh_chunks)
    assert len(tr._chunks) == 1 + sh_chunks
    assert [tr.image_index(s, 4) for s in range(3)] == [(s * world + rank) % 4 for s in range(3)]
    cams = [dict(w=float(i + 1)) for i in range(4)]
    for step in range(3):
        i = tr.image_index(step, 4)
        tr.train_step(cams[i], torch.zeros(scene["H"], scene["W"], 3), i)
    # replicas identical after 3 steps although every rank saw different images
    gathered = [torch.zeros_like(model.flat) for _ in range(world)]
    dist.all_gather(gathered, model.flat)
    assert all(torch.equal(gathered[0], g_) for g_ in gathered[1:])
    # the overlapped exchange (SH slice reduced from the autograd hook) gives the same parameters as the plain one
    model2 = dp.FlatGaussians(scene, "cpu", n_cameras=4, optimize_velocities=True)
    model2.params["means"].data[0, 0] = 7.0  # same aliasing probe as the first model
    tr2 = dp.ImageShardedTrainer(model2, scene, lr=1e-2, overlap_sh=False, loss_fn=_torch_l1, optimizer="torch")
    assert tr.overlap_sh and not tr2.overlap_sh
    for step in range(3):
        i = tr2.image_index(step, 4)
        tr2.train_step(cams[i], torch.zeros(scene["H"], scene["W"], 3), i)
    assert torch.allclose(model.flat.detach(), model2.flat.detach(), rtol=0, atol=1e-6)
    # layout: geometry + opacity rows, camera rows, then ONE contiguous SH block at the tail
    assert model.slices["cam_vel"] == (50 * 11, 50 * 11 + 24) and model.sh_start == 50 * 11 + 24 + 2  # (+2: 16-byte grid)
    # camera-velocity rows are disjoint per image: rows of images nobody rendered this step keep zero grad
    last = {(2 * world + r) % 4 for r in range(world)}
    for c in range(4):
        assert (model.cam_vel.grad[c].abs().sum() > 0) == (c in last)
    # the averaged gradient equals the mean of per-rank gradients: recompute locally without the trainer
    g = torch.ones(50) * (rank + 1)
    v = torch.ones(50) * (rank + 1)
    mx = torch.ones(50) * (rank + 1)
    g, v, mx = tr.reduce_densify_stats(g, v, mx)
    tot = world * (world + 1) // 2
    assert torch.all(g == tot) and torch.all(v == tot) and torch.all(mx == world)
    if rank == 0:
        torch.save(model.flat.clone(), out)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,sh_chunks", [(2, 1), (3, 3)])
def test_image_sharded_trainer_gloo(tmp_path, world, sh_chunks):
    """world 2 with the default exchange (geometry chunk + one SH chunk) and world 3 with the SH block in 3 chunks: the
    chunked, overlapped exchange ends at the same parameters as the plain one, replicas stay identical."""
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(world, _free_port(), out, sh_chunks), nprocs=world, join=True)
    flat = torch.load(out)
    assert torch.isfinite(flat).all()


def test_single_process_trainer_matches_manual_adam():
    sys.path.insert(0, os.path.join(ROOT, "3dgs-deblur_b200"))
    from gsplat import dp, synthetic

    dp.render = _fake_render
    scene = synthetic.make_scene("c1", n_override=20)
    model = dp.FlatGaussians(scene, "cpu")
    ref = model.flat.detach().clone()
    tr = dp.ImageShardedTrainer(model, scene, lr=1e-2, loss_fn=_torch_l1, optimizer="torch")
    loss = tr.train_step(dict(w=1.0), torch.zeros(scene["H"], scene["W"], 3))
    assert torch.isfinite(loss)
    # first Adam step moves every parameter with a non-zero gradient by lr (bias-corrected), opposite to its sign
    g = model.flat_grad
    moved = (model.flat.detach() - ref)
    nz = g != 0
    assert torch.allclose(moved[nz], -1e-2 * torch.sign(g[nz]), atol=1e-6)
    assert (moved[~nz] == 0).all()


# ---- PipelinedTrainer (phase A / phase B split, exchange behind the next image's geometry) ---------------------------

def _fake_geometry(model, st, scene, capacity, status):
    """Stand-in for gsplat.dp.geometry_phase: depends on the geometry rows and the camera only."""
    p = model.params
    w = st["cam"][0]
    geo = (p["means"].sum() * w + torch.sigmoid(p["opacity_logit"]).sum() + p["log_scales"].exp().sum()
           + (p["quats"] / p["quats"].norm(dim=-1, keepdim=True)).sum())
    if model.cam_vel is not None:
        geo = geo + (model.cam_vel.index_select(0, st["cam_index"])[0] * w).sum()
    if float(w) >= 100.0:  # a camera that "needs more list entries than the capacity": the device would raise the flag
        status[0] = 1
    return dict(geo=geo)


def _fake_shading(model, geo, scene, target, loss_fn, sh_degree_to_use=3):
    sh = model.sh_coeffs()
    rgb = geo["geo"] + sh[:, :1].sum() + (sh[:, 1:] ** 2).sum()
    img = rgb * torch.ones(scene["H"], scene["W"], 3)
    loss = loss_fn(img, target)
    loss.backward()
    return loss.detach()


def _cam_row(w):
    return torch.tensor([float(w)] + [0.0] * 20)


def _pipelined_worker(rank, world, port, out):
    sys.path.insert(0, os.path.join(ROOT, "3dgs-deblur_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gsplat import dp, synthetic

    scene = synthetic.make_scene("c1", n_override=50, n_cameras=4)
    scene.update(fx=1.0, fy=1.0, cx=0.0, cy=0.0)
    tgt = torch.zeros(scene["H"], scene["W"], 3)
    # reference: the plain trainer on the same images (same math: averaged gradients, Adam eps 1e-15)
    dp.render = _fake_render
    m0 = dp.FlatGaussians(scene, "cpu", n_cameras=4, optimize_velocities=True)
    t0 = dp.ImageShardedTrainer(m0, scene, lr=1e-2, loss_fn=_torch_l1, optimizer="torch")
    for step in range(4):
        i = t0.image_index(step, 4)
        t0.train_step(dict(w=float(i + 1)), tgt, i)
    # pipelined trainer, block SH layout, fake phases with the same total function
    m1 = dp.FlatGaussians(scene, "cpu", n_cameras=4, optimize_velocities=True, sh_layout="block")
    assert m1.sh_start % 4 == 0 and m1.params["sh"].shape == (50, 16, 3)
    t1 = dp.PipelinedTrainer(m1, scene, lr=1e-2, loss_fn=_torch_l1, optimizer="torch", geometry_fn=_fake_geometry,
                             shading_fn=_fake_shading, capacity=1, sh_chunks=2)  # (two SH pieces: the chunked exchange path)
    assert len(t1._sh_bounds) == 2
    order = [t0.image_index(s, 4) for s in range(4)]
    t1.prepare(_cam_row(order[0] + 1), order[0])
    for step in range(4):
        nxt = order[step + 1] if step + 1 < 4 else None
        t1.train_step(tgt, None if nxt is None else _cam_row(nxt + 1), 0 if nxt is None else nxt)
    t1.finish()
    for k in ("means", "log_scales", "quats", "opacity_logit"):
        assert torch.allclose(m0.params[k].detach(), m1.params[k].detach(), rtol=0, atol=1e-6), k
    assert torch.allclose(m0.cam_vel.detach(), m1.cam_vel.detach(), rtol=0, atol=1e-6)
    assert torch.allclose(torch.cat((m0.params["sh_dc"], m0.params["sh_rest"]), 1).detach(), m1.params["sh"].detach(), rtol=0, atol=1e-6)
    gathered = [torch.zeros_like(m1.flat) for _ in range(world)]
    dist.all_gather(gathered, m1.flat)
    assert all(torch.equal(gathered[0], g_) for g_ in gathered[1:])  # replicas identical
    assert float(m1.flat_grad.abs().max()) == 0.0                    # gradients cleared behind the update
    # veto: ONE rank's image overflows -> the MAX-reduced flag skips the update on EVERY rank, gradients are cleared
    before = m1.flat.detach().clone()
    t1.prepare(_cam_row(100.0 if rank == 0 else 1.0), 0)
    t1.train_step(tgt)
    t1.finish()
    assert torch.equal(m1.flat.detach(), before) and float(m1.flat_grad.abs().max()) == 0.0
    assert t1.vetoed == [4] and int(t1.flag[0]) == 0
    # and the step after it applies again
    t1.prepare(_cam_row(1.0), 0)
    t1.train_step(tgt)
    t1.finish()
    assert not torch.equal(m1.flat.detach(), before)
    # reserve(): lists sized once from a known entry count (same number on every rank); never shrinks; steady() is
    # trivially true without CUDA graphs
    t1.reserve(1000)
    assert t1.capacity == 131072 and t1.steady(((scene["H"], scene["W"], 3), torch.float32))
    t1.reserve(10)
    assert t1.capacity == 131072
    if rank == 0:
        torch.save(m1.flat.clone(), out)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_pipelined_trainer_gloo(tmp_path, world):
    """world 2 / 4: the pipelined trainer (geometry phase of image k+1 issued before the SH slice of step k is updated) ends
    at the same parameters as the plain trainer, replicas stay identical, and an overflow on one rank vetoes the step on
    all of them."""
    out = str(tmp_path / "flat2.pt")
    mp.spawn(_pipelined_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert torch.isfinite(torch.load(out)).all()
