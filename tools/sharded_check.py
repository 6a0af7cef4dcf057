"""Multi-GPU parity: G ranks must reproduce the 1-GPU result bit for bit (indices(G) == indices(1)).
Covers the peer-memory exchange (csrc/shard.cu, the product path) over several consecutive frames, even and ragged
partitions, with and without the fused range-beacon weight, and the NCCL all-gather formulation kept as the A/B baseline.

python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py
(run by tests/test_gpu_sharded.py when >= 2 GPUs are visible, and by bench.py --gpus N before timing)"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from amcl3d_test_b200 import Grid3d, ParticleFilter, synth  # noqa: E402
from amcl3d_test_b200.sharded import CudaShard, PeerShardedFilter, ShardedFilter, shard_bounds  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def one_gpu_frames(g, P, clouds, u01s, beacons=None):
    """The same frames on one GPU through the plain handle: per frame (weights after update, resampled set, indices, mean)."""
    pf = ParticleFilter(g)
    pf.p_ = P
    out = []
    for cloud, u01 in zip(clouds, u01s):
        if beacons is not None:
            pf.update(cloud, beacons=beacons, alpha=0.5, sigma=0.53)
        else:
            pf.update(cloud)
        W = pf.p_
        pf.particleNormalize()
        idx = pf.resample(u01, want_idx=True)
        out.append((W, pf.p_, idx, pf.getMean()))
    return out


def exact_prefix(g, W):
    """Inclusive sequential f32 prefix of the normalised weights of the one-GPU run (ParticleFilter.cpp:235-247)."""
    pf = ParticleFilter(g)
    pf.p_ = W
    pf.particleNormalize()
    return pf.prefix()


def peer_frames(g, P, clouds, u01s, dev, beacons=None):
    n = P.shape[0]
    f = PeerShardedFilter(g, n, dev)
    f.set_particles(P[f.lo:f.hi])
    if beacons is not None:
        f.set_beacons(beacons, 0.5, 0.53)
    out = []
    for cloud, u01 in zip(clouds, u01s):
        f.set_cloud(cloud)
        f.update()
        W = f.pf.p_
        idx = f.resample(u01, want_idx=True)
        out.append((W, f.pf.p_, idx, f.mean()))
    return f, out


def compare(tag, rank, lo, hi, got, ref, weights_fused=False):
    ok = True
    for k, ((W, out, idx, mean), (rW, rout, ridx, rmean)) in enumerate(zip(got, ref)):
        if not weights_fused and not np.array_equal(bits(W[:, 6]), bits(rW[lo:hi, 6])):
            print(f"rank {rank} {tag} frame {k}: weights differ from 1 GPU")
            ok = False
        if not np.array_equal(idx, ridx[lo:hi]):
            print(f"rank {rank} {tag} frame {k}: {int((idx != ridx[lo:hi]).sum())} resampled indices differ from 1 GPU")
            ok = False
        if not np.array_equal(bits(out), bits(rout[lo:hi])):
            print(f"rank {rank} {tag} frame {k}: resampled particles differ from 1 GPU")
            ok = False
        if not np.allclose(mean[:6], rmean[:6], rtol=1e-5, atol=1e-6) or mean[6] != rmean[6]:
            print(f"rank {rank} {tag} frame {k}: mean {mean} vs {rmean}")
            ok = False
    return ok


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_pts = 3000
    sm = synth.make_map((20.0, 20.0, 5.0), 0.1)
    clouds = [synth.make_cloud(sm, n_pts, 9.0, seed=synth.SEED_CLOUD + k) for k in range(3)]
    u01s = [0.37, 0.0, 0.93]
    g = Grid3d(local)
    g.computePointCloud(sm.points, 0.1)
    g.computeGrid(sm.points, 0.05, sub=2)
    beacons = synth.make_beacons(sm, 8)
    ok = True
    for n, kind in ((40_000, "T"), (40_003, "G"), (world * 4096 + 1, "T")):
        P = synth.make_particles(sm, n, kind)
        lo, hi = shard_bounds(n, world, rank)
        ref = one_gpu_frames(g, P, clouds, u01s)
        f, got = peer_frames(g, P, clouds, u01s, dev)
        ok = compare(f"peer n={n} {kind}", rank, lo, hi, got, ref) and ok
        del f
        ref_b = one_gpu_frames(g, P, clouds, u01s, beacons)
        f, got_b = peer_frames(g, P, clouds, u01s, dev, beacons)
        ok = compare(f"peer+beacons n={n} {kind}", rank, lo, hi, got_b, ref_b, weights_fused=True) and ok
        del f
    # fast (non-exact) mode: structurally valid resampled set, and how many indices differ from the exact result
    for n, kind in ((40_000, "T"), (40_003, "G")):
        P = synth.make_particles(sm, n, kind)
        lo, hi = shard_bounds(n, world, rank)
        ref = one_gpu_frames(g, P, clouds[:1], u01s[:1])
        f = PeerShardedFilter(g, n, dev)
        for rep in range(2):  # twice: the exchange slots ping-pong
            f.set_particles(P[f.lo:f.hi])
            f.set_cloud(clouds[0])
            f.update()
            idx = f.resample_fast(u01s[0], want_idx=True)
            out = f.pf.p_
            bad = int((idx != ref[0][2][lo:hi]).sum())
            tot = torch.tensor([bad], device=dev)
            dist.all_reduce(tot)
            valid = bool(np.all(np.diff(idx) >= 0) and idx.min() >= 0 and idx.max() < n and
                         np.array_equal(bits(out[:, :6]), bits(P[idx, :6])) and np.all(out[:, 6] == np.float32(1.0) / np.float32(n)))
            if not valid:
                print(f"rank {rank} fast n={n} {kind}: resampled slice is not a valid systematic resample")
                ok = False
            # a differing index must still be a defensible choice: the threshold lies within a few ulps of the EXACT
            # prefix interval of the particle the fast mode picked (the two modes only differ in how partial sums round)
            Wn = ref[0][0].copy()
            pfx = exact_prefix(g, Wn)
            um = np.float32(u01s[0]) * (np.float32(1.0) / np.float32(n)) + (np.float32(1.0) / np.float32(n)) * np.arange(lo, hi, dtype=np.float32)
            # accumulated rounding of a sequential f32 sum of n terms below 1 is a random walk of ~sqrt(n) half-ulps
            eps = np.float32(8.0 * np.sqrt(n) * 2.0 ** -24)
            upper_ok = um <= pfx[idx] + eps
            lower_ok = (idx == 0) | (um > np.where(idx > 0, pfx[np.maximum(idx - 1, 0)], np.float32(0)) - eps)
            tail = idx == n - 1
            if not np.all((upper_ok | tail) & lower_ok):
                print(f"rank {rank} fast n={n} {kind}: an index is not within {eps} (f32 summation noise) of the exact prefix interval")
                ok = False
            if rank == 0 and rep == 0:
                print(f"fast mode n={n} {kind} world={world}: {int(tot.item())} of {n} resampled indices differ from the exact mode")
            mean = f.mean()
            if not np.allclose(mean[:6], ref[0][3][:6], rtol=1e-3, atol=1e-3):
                print(f"rank {rank} fast n={n} {kind}: mean {mean} vs {ref[0][3]}")
                ok = False
        del f
    # the NCCL all-gather formulation (A/B baseline)
    n = 40_000
    P = synth.make_particles(sm, n, "T")
    lo, hi = shard_bounds(n, world, rank)
    shard = CudaShard(g, hi - lo, n, dev)
    shard.local.copy_(torch.from_numpy(P[lo:hi]).to(dev))
    shard.set_cloud(clouds[0])
    f = ShardedFilter(shard, n)
    f.update()
    f.resample(0.37)
    ref = one_gpu_frames(g, P, clouds[:1], u01s[:1])
    if not np.array_equal(bits(shard.local.cpu().numpy()), bits(ref[0][1][lo:hi])):
        print(f"rank {rank}: NCCL all-gather formulation differs from 1 GPU")
        ok = False
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"sharded_check world={world}: {'OK bit-identical to 1 GPU' if t.item() == 1 else 'MISMATCH'}")
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
