"""Multi-GPU parity: G ranks (NCCL) must reproduce the 1-GPU / CPU result bit for bit.
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from amcl3d_test_b200 import Grid3d, ParticleFilter, synth  # noqa: E402
from amcl3d_test_b200.sharded import CudaShard, ShardedFilter, shard_bounds  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n, n_pts = 40_000, 3000
    sm = synth.make_map((20.0, 20.0, 5.0), 0.1)
    cloud = synth.make_cloud(sm, n_pts, 9.0)
    P = synth.make_particles(sm, n, "T")
    g = Grid3d(local)
    g.computePointCloud(sm.points, 0.1)
    g.computeGrid(sm.points, 0.05, sub=2)
    lo, hi = shard_bounds(n, world, rank)
    shard = CudaShard(g, hi - lo, n, dev)
    shard.local.copy_(torch.from_numpy(P[lo:hi]).to(dev))
    shard.set_cloud(cloud)
    f = ShardedFilter(shard, n)
    f.update()
    w_local = shard.local[:, 6].cpu().numpy().copy()
    f.resample(0.37)
    out_local = shard.local.cpu().numpy().copy()
    mean = f.mean()
    # single-GPU reference on this rank
    pf = ParticleFilter(g)
    pf.p_ = P
    pf.update(cloud)
    W = pf.p_
    pf.particleNormalize()
    pf.resample(0.37)
    ref_out = pf.p_
    ok = np.array_equal(w_local.view(np.uint32), W[lo:hi, 6].view(np.uint32)) and np.array_equal(
        out_local.view(np.uint32), ref_out[lo:hi].view(np.uint32))
    ref_mean = pf.getMean()
    ok = ok and np.allclose(mean[:6], ref_mean[:6], rtol=1e-5, atol=1e-6)
    # same with the range-beacon weight fused in (global sums of the fusion across ranks)
    beacons = synth.make_beacons(sm, 8)
    shard.set_beacons(beacons, 0.5, 0.53)
    shard.local.copy_(torch.from_numpy(P[lo:hi]).to(dev))
    f.update()
    f.resample(0.37)
    out_b = shard.local.cpu().numpy().copy()
    pf.p_ = P
    pf.update(cloud, beacons=beacons, alpha=0.5, sigma=0.53)
    pf.particleNormalize()
    pf.resample(0.37)
    ok_b = np.array_equal(out_b.view(np.uint32), pf.p_[lo:hi].view(np.uint32))
    if not ok_b:
        print(f"rank {rank}: beacon-fused resample differs from 1 GPU")
    ok = ok and ok_b
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"sharded_check world={world}: {'OK bit-identical to 1 GPU' if t.item() == 1 else 'MISMATCH'}")
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
