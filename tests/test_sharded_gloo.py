"""world_size-2 gloo test of the multi-GPU orchestration (amcl3d_test_b200/sharded.py) on CPU.

The ShardedFilter plumbing (block partition, all-gather of the weighted blocks, replicated sequential chain,
per-rank output slice, mean all-reduce) is exercised with the CPU oracle standing in for the CUDA shard, and
must reproduce the single-process result bit for bit (weights, resampled particles) for even and ragged splits."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


class OracleShard:
    """CPU stand-in for CudaShard: same interface, compute by the oracle (test infrastructure)."""

    def __init__(self, O, m, cloud, local_np, n_total):
        self.O, self.m, self.cloud = O, m, cloud
        self.local = torch.from_numpy(local_np.copy())
        self.full = torch.zeros((n_total, 7), dtype=torch.float32)

    def update(self):
        self.O.update(self.m, self.local.numpy(), self.cloud)

    def normalize_full(self):
        return self.O.normalize(self.m, self.full.numpy())

    def resample_full(self, u01, m0, m1):
        out, _ = self.O.resample(self.full.numpy(), u01)
        self.local.copy_(torch.from_numpy(out[m0:m1]))

    def mean_local(self):
        return self.O.mean(self.local.numpy())


def _world(n):
    from amcl3d_test_b200 import synth
    from oracle import loader as O

    sm = synth.make_map((6.0, 6.0, 3.0), 0.1)
    mn, mx = O.compute_bounds(sm.points)
    m = O.Map(mn, mx, 0.1, 0.05)
    m.set_prob(O.compute_grid2(m, sm.points))
    cloud = synth.make_cloud(sm, 300, 2.5)
    P = synth.make_particles(sm, n, "T")
    return O, m, cloud, P


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from amcl3d_test_b200.sharded import ShardedFilter, shard_bounds

    O, m, cloud, P = _world(n)
    lo, hi = shard_bounds(n, world, rank)
    shard = OracleShard(O, m, cloud, P[lo:hi], n)
    f = ShardedFilter(shard, n)
    f.update()
    w_local = shard.local.numpy()[:, 6].copy()
    wtp = f.resample(0.37)
    mean = f.mean()
    np.savez(Path(out_dir) / f"rank{rank}.npz", lo=lo, hi=hi, w=w_local, out=shard.local.numpy(), wtp=wtp, mean=mean)
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [200, 201])
def test_two_ranks_match_single_process(tmp_path, n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    O, m, cloud, P = _world(n)
    W = O.update(m, O.particles(P), cloud)
    ref_w = W[:, 6].copy()
    wtp = O.normalize(m, W)
    ref_out, _ = O.resample(W, np.float32(0.37))
    ref_mean = O.mean(ref_out)
    covered = 0
    for r in range(2):
        z = np.load(tmp_path / f"rank{r}.npz")
        lo, hi = int(z["lo"]), int(z["hi"])
        covered += hi - lo
        assert np.array_equal(z["w"].view(np.uint32), ref_w[lo:hi].view(np.uint32))
        assert np.array_equal(z["out"].view(np.uint32), ref_out[lo:hi].view(np.uint32))
        assert np.float32(z["wtp"]).view(np.uint32) == np.float32(wtp).view(np.uint32)
        assert np.allclose(z["mean"][:6], ref_mean[:6], rtol=1e-5, atol=1e-6)
        assert z["mean"][6] == ref_mean[6]
    assert covered == n


def test_shard_bounds_partition():
    from amcl3d_test_b200.sharded import shard_bounds

    for n in (0, 1, 7, 8, 1000, 1001):
        for world in (1, 2, 3, 8):
            edges = [shard_bounds(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
