"""Per-stage device times of one sharded step (events on the shared stream), async and sync mode.
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/sharded_stage_times.py"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from amcl3d_test_b200 import Grid3d, synth  # noqa: E402
from amcl3d_test_b200.sharded import CudaShard, ShardedFilter  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = synth.CONFIGS["C3"]
    sm = synth.make_map(cfg["extent"], cfg["resol"])
    cloud = synth.make_cloud(sm, cfg["points"], cfg["radius"])
    n_local = cfg["particles"]
    n_total = n_local * world
    P = synth.make_particles(sm, n_total, "T")[rank * n_local:(rank + 1) * n_local]
    g = Grid3d(local)
    g.computePointCloud(sm.points, cfg["resol"])
    g.computeGrid(sm.points, 0.05, sub=2)
    shard = CudaShard(g, n_local, n_total, dev)
    filt = ShardedFilter(shard, n_total)
    backup = torch.from_numpy(np.ascontiguousarray(P)).to(dev)
    shard.set_cloud(cloud)
    stream = torch.cuda.current_stream(dev)
    names = ["copy", "update", "gather", "normalize", "resample", "mean"]
    for asyn in (False, True):
        shard.set_async(asyn)
        acc = np.zeros(len(names))
        steps = 30
        for it in range(steps + 3):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
            ev[0].record(stream)
            shard.local.copy_(backup)
            ev[1].record(stream)
            filt.update()
            ev[2].record(stream)
            filt.gather()
            ev[3].record(stream)
            shard.normalize_full()
            ev[4].record(stream)
            shard.resample_full(0.37, filt.lo, filt.hi)
            ev[5].record(stream)
            filt.mean_device() if asyn else filt.mean()
            ev[6].record(stream)
            torch.cuda.synchronize()
            if it >= 3:
                acc += [ev[k].elapsed_time(ev[k + 1]) for k in range(len(names))]
        if rank == 0:
            print(f"world={world} async={asyn}: " + "  ".join(f"{n}={v / steps * 1e3:.0f}us" for n, v in zip(names, acc)) +
                  f"  total={acc.sum() / steps:.3f}ms", flush=True)
    # the bench loop: free-running, stream-ordered, one pair of events around everything
    for variant in ("bench", "bench+barrier"):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 60
        torch.cuda.synchronize()
        if world > 1 and variant == "bench+barrier":
            dist.barrier()
        e0.record(stream)
        for k in range(steps):
            shard.local.copy_(backup)
            filt.update()
            filt.resample(0.37)
            filt.mean_device()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
        if world > 1:
            g_t = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(g_t, t)
            t_all = [float(x.item()) for x in g_t]
        else:
            t_all = [float(t.item())]
        if rank == 0:
            print(f"world={world} {variant} free-running per-rank ms/step: {t_all}", flush=True)
    shard.set_async(False)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
