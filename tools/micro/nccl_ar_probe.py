"""All-reduce latency probe for the two message sizes of a PipelinedTrainer step (geometry block 11 N floats, SH block
48 N floats in `chunks` pieces) under whatever NCCL_* environment the caller set:
    torchrun --nproc-per-node 4 tools/micro/nccl_ar_probe.py --n 300000 --tag default
Device time per collective (CUDA events, max over ranks), alone on an idle GPU -- the trainer's exchange competes with
its own kernels, so these are lower bounds."""
import argparse
import json
import os

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=300000)
ap.add_argument("--chunks", type=int, default=2)
ap.add_argument("--tag", default="default")
ap.add_argument("--iters", type=int, default=40)
ap.add_argument("--op", default="avg", choices=["avg", "sum"])
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = {"tag": a.tag, "world": world, "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}
OP = dist.ReduceOp.AVG if a.op == "avg" else dist.ReduceOp.SUM
out["op"] = a.op
for name, numel in (("geometry", 11 * a.n), ("sh_chunk", 48 * a.n // a.chunks), ("sh_whole", 48 * a.n), ("flag", 1),
                    ("sh_reduce_scatter", 48 * a.n), ("sh_all_gather", 48 * a.n), ("geo_reduce_scatter", 11 * a.n // 32 * 32),
                    ("geo_all_gather", 11 * a.n // 32 * 32)):
    buf = torch.randn(numel, device="cuda") if numel > 1 else torch.zeros(1, dtype=torch.int32, device="cuda")
    shard = numel // world
    mine = buf[rank * shard:(rank + 1) * shard] if numel > 1 else None   # in place, like the trainer

    def call():
        if name.endswith("reduce_scatter"):
            dist.reduce_scatter_tensor(mine, buf, op=OP)
        elif name.endswith("all_gather"):
            dist.all_gather_into_tensor(buf, mine)
        else:
            dist.all_reduce(buf, op=OP if numel > 1 else dist.ReduceOp.MAX)

    for _ in range(5):
        call()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / a.iters * 1e3], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    us = float(t)
    out[name] = {"bytes": 4 * numel, "us": round(us, 1), "busbw_GBs": round(2 * (world - 1) / world * 4 * numel / us / 1e3, 1)}
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
