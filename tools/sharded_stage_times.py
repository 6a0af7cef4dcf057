"""Per-stage device times of the sharded step (peer-memory exchange), every rank: update | exact resample | mean | fast resample.
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tools/sharded_stage_times.py"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from amcl3d_test_b200 import Grid3d, synth  # noqa: E402
from amcl3d_test_b200.sharded import PeerShardedFilter  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = synth.CONFIGS["C3"]
    n_total = cfg["particles"] * world
    sm = synth.make_map(cfg["extent"], cfg["resol"])
    cloud = synth.make_cloud(sm, cfg["points"], cfg["radius"])
    g = Grid3d(local)
    g.computePointCloud(sm.points, cfg["resol"])
    g.computeGrid(sm.points, 0.05, sub=2)
    f = PeerShardedFilter(g, n_total, dev)
    P = synth.make_particles(sm, n_total, "T")[f.lo:f.hi]
    backup = torch.from_numpy(np.ascontiguousarray(P)).to(dev)
    f.set_cloud(cloud)
    stream = torch.cuda.current_stream(dev)
    steps = 30
    for mode in ("exact", "fast"):
        f.set_async(False)
        f.set_particles_device(backup)
        f.update()
        (f.resample if mode == "exact" else f.resample_fast)(0.37)
        f.mean()
        f.set_async(True)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(steps)]
        torch.cuda.synchronize()
        dist.barrier()
        for k in range(steps):
            ev[k][0].record(stream)
            f.set_particles_device(backup)
            ev[k][1].record(stream)
            f.update()
            ev[k][2].record(stream)
            (f.resample if mode == "exact" else f.resample_fast)(0.37)
            ev[k][3].record(stream)
            f.mean_device()
            ev[k][4].record(stream)
        torch.cuda.synchronize()
        f.pf.sync()
        t = np.array([[ev[k][i].elapsed_time(ev[k][i + 1]) for i in range(4)] for k in range(5, steps)])
        m = t.mean(0)
        msg = (f"[{mode}] rank {rank}: copy {m[0]:.3f} update {m[1]:.3f} resample {m[2]:.3f} (min {t[:, 2].min():.3f} max {t[:, 2].max():.3f}) "
               f"mean {m[3]:.3f} total {m.sum():.3f} ms")
        out = [None] * world
        dist.all_gather_object(out, msg)
        if rank == 0:
            print("\n".join(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
