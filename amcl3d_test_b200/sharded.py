"""Particles sharded across the GPUs of one box; one process per GPU, torch.distributed.

The product path is PeerShardedFilter (bottom of the file): the exchange step runs in the library's own kernels over
peer memory (csrc/shard.cu, C-ABI amcl3d_b200_pf_shard_*); torch.distributed only carries the 64-byte CUDA IPC handles
at set-up.  ShardedFilter + CudaShard is the older NCCL all-gather formulation, kept as the A/B baseline and because
its orchestration runs on CPU with gloo.

Partitioning: the grid is replicated, rank r owns the contiguous block [r*n_local, (r+1)*n_local) of the
global particle order (contiguous blocks keep the order the reference's serial prefix needs).  The update
is embarrassingly parallel (no collective).  Normalise + resample has one real exchange step: an
all-gather of the freshly weighted particle blocks (28 B x N over NVLink); every rank then replays the
reference's sequential f32 chains over the full set -- so weights, prefix and resampled indices are
bit-identical to the 1-GPU / CPU result -- and produces only its own slice [m0, m1) of the output.
The mean is a sum of per-rank partial means (all-reduce of 7 floats).

The collective plumbing is backend-agnostic so that the same orchestration is exercised on CPU with gloo
(tests/test_sharded_gloo.py, compute supplied by the test) and on GPUs with NCCL (CudaShard below).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_total: int, world: int, rank: int):
    """Contiguous block partition; the first (n_total % world) ranks get one extra particle."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedFilter:
    """Orchestration only.  `shard` supplies the compute on this rank's memory:

    shard.local  : tensor (n_local, 7)   this rank's particle block (weights in column 6)
    shard.full   : tensor (N, 7)         gather target, same device
    shard.update()                                   -> weights of shard.local
    shard.normalize_full()                           -> particleNormalize over shard.full, returns wtp
    shard.resample_full(u01, m0, m1)                 -> writes outputs [m0, m1) into shard.local
    shard.uniform_full(S, k0, k1) -> (emitted)       -> writes outputs [k0, k1) into shard.local_out(S)
    shard.mean_local()                               -> 7 floats (weighted sums over shard.local, max w)
    """

    def __init__(self, shard, n_total: int, group=None):
        self.shard = shard
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = n_total
        self.lo, self.hi = shard_bounds(n_total, self.world, self.rank)
        assert shard.local.shape[0] == self.hi - self.lo

    # -- exchange step -----------------------------------------------------------------------------
    def gather(self):
        full, local = self.shard.full, self.shard.local
        if self.world == 1:
            full.copy_(local)
            return
        if self.n_total % self.world == 0:
            dist.all_gather_into_tensor(full, local, group=self.group)
        else:  # ragged blocks: gather equal-sized padded blocks, then compact into the global order
            pad = -(-self.n_total // self.world)
            send = torch.zeros((pad, 7), dtype=local.dtype, device=local.device)
            send[: local.shape[0]] = local
            recv = torch.empty((self.world * pad, 7), dtype=local.dtype, device=local.device)
            dist.all_gather_into_tensor(recv, send, group=self.group)
            for r in range(self.world):
                lo, hi = shard_bounds(self.n_total, self.world, r)
                full[lo:hi] = recv[r * pad : r * pad + (hi - lo)]

    # -- the path ----------------------------------------------------------------------------------
    def update(self):
        self.shard.update()

    def resample(self, u01: float):
        """update() must have run.  Global ParticleFilter::particleNormalize + resample, bit-exact."""
        self.gather()
        if getattr(self.shard, "wr_local", None) is not None:
            # range-beacon fusion needs the GLOBAL sums of both weights: gather the raw range weights too, fuse on
            # the full set (every rank, redundantly), then normalise as usual
            if self.world == 1:
                self.shard.wr_full.copy_(self.shard.wr_local)
            else:
                assert self.n_total % self.world == 0, "beacon fusion across ranks needs equal blocks"
                dist.all_gather_into_tensor(self.shard.wr_full, self.shard.wr_local, group=self.group)
            self.shard.fuse_full()
        wtp = self.shard.normalize_full()
        self.shard.resample_full(u01, self.lo, self.hi)
        return wtp

    def mean_device(self):
        """Global mean as a 7-float device tensor, nothing read back: local weighted sums on the device, one
        all-gather of 7 floats per rank, combined on the device (sums in f64, w = max)."""
        loc = self.shard.mean_local_device()
        if self.world == 1:
            return loc
        if getattr(self, "_mean_all", None) is None:
            self._mean_all = torch.empty((self.world, 7), dtype=torch.float32, device=loc.device)
        dist.all_gather_into_tensor(self._mean_all, loc.view(1, 7), group=self.group)
        sums = self._mean_all[:, :6].to(torch.float64).sum(0).to(torch.float32)
        return torch.cat([sums, self._mean_all[:, 6].max().view(1)])

    def mean(self):
        part = torch.as_tensor(np.asarray(self.shard.mean_local(), np.float64))
        if self.world > 1:
            dev = self.shard.local.device
            sums = part[:6].to(dev)
            wmax = part[6:].to(dev)
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(wmax, op=dist.ReduceOp.MAX, group=self.group)
            part = torch.cat([sums, wmax]).cpu()
        return part.numpy().astype(np.float32)


class CudaShard:
    """This rank's CUDA state: a grid replica, the local block and the gather target as torch tensors
    bound (zero-copy) to two C-ABI ParticleFilter handles."""

    def __init__(self, grid3d, n_local: int, n_total: int, device: torch.device):
        from .api import ParticleFilter

        self.device = device
        self.local = torch.zeros((n_local, 7), dtype=torch.float32, device=device)
        self.full = torch.zeros((n_total, 7), dtype=torch.float32, device=device)
        self.pf_local = ParticleFilter(grid3d)
        self.pf_full = ParticleFilter(grid3d)
        self.pf_local.bind_device(self.local.data_ptr(), n_local, n_local, keep=self.local)
        self.pf_full.bind_device(self.full.data_ptr(), n_total, n_total, keep=self.full)
        stream = torch.cuda.current_stream(device).cuda_stream
        self.pf_local.set_stream(stream)
        self.pf_full.set_stream(stream)

    def set_cloud(self, cloud):
        self.pf_local.set_cloud(cloud)

    wr_local = None

    def set_beacons(self, beacons, alpha, sigma):
        """Range-only beacon weighting (row R): (positions (nb,3), ranges (nb,)), alpha, sigma.  The range weight is
        evaluated in the update kernel; the fusion runs on the gathered set (global sums)."""
        self.beacons, self.alpha, self.sigma = beacons, alpha, sigma
        self.wr_local = torch.zeros(self.local.shape[0], dtype=torch.float32, device=self.device)
        self.wr_full = torch.zeros(self.full.shape[0], dtype=torch.float32, device=self.device)

    def update(self):
        if self.wr_local is not None:
            self.pf_local.update_split(self.beacons, self.sigma, self.wr_local.data_ptr())
        else:
            self.pf_local.update()

    def fuse_full(self):
        self.pf_full.fuse_range(self.wr_full.data_ptr(), self.alpha)

    def set_async(self, enable: bool):
        """Stream order for the stages that return nothing to the host; mean() stays the one wait of a step."""
        self.async_ = bool(enable)
        self.pf_local.set_async(enable)
        self.pf_full.set_async(enable)

    async_ = False

    def normalize_full(self):
        return self.pf_full.particleNormalize(fetch=not self.async_)

    def resample_full(self, u01, m0, m1):
        self.pf_full.resample(u01, m0=m0, m1=m1, d_out=self.local.data_ptr())

    def mean_local(self):
        return self.pf_local.getMean(exact=False)

    def mean_local_device(self):
        if getattr(self, "_mean_buf", None) is None:
            self._mean_buf = torch.zeros(7, dtype=torch.float32, device=self.device)
        self.pf_local.mean_device(self._mean_buf.data_ptr(), exact=False)
        return self._mean_buf

    def launch_count(self):
        return self.pf_local.launch_count() + self.pf_full.launch_count()


class PeerShardedFilter:
    """One rank of a particle set sharded over the GPUs of one NVLink box, exchange in our own kernels over peer
    memory (csrc/shard.cu): 4-byte weights are stored into every peer's dense copy, the sequential chains are
    replayed on every rank (bit-identical to one GPU), the selected particles are loaded from their owners.
    torch.distributed is used once, to hand the CUDA IPC handles round."""

    def __init__(self, grid3d, n_total: int, device: torch.device, group=None):
        from .api import ParticleFilter

        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = n_total
        self.device = device
        self.pf = ParticleFilter(grid3d)
        self.lo, self.hi = self.pf.shard_init(n_total, self.rank, self.world)
        assert (self.lo, self.hi) == shard_bounds(n_total, self.world, self.rank)
        if self.world > 1:
            mine = torch.frombuffer(bytearray(self.pf.shard_ipc_handle()), dtype=torch.uint8).to(device)
            allh = torch.empty((self.world, 64), dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(allh, mine.view(1, 64), group=group)
            self.pf.shard_attach_ipc([bytes(r) for r in allh.cpu().numpy()])
            dist.barrier(group=group)
        self.pf.set_stream(torch.cuda.current_stream(device).cuda_stream)
        self.beacons = None
        self.wr_local = None
        self._mean_buf = torch.zeros(8, dtype=torch.float32, device=device)

    def set_particles(self, block):
        """This rank's block [lo, hi) of the global particle order, (hi - lo, 7) float32 on the host."""
        assert block.shape[0] == self.hi - self.lo
        self.pf.p_ = block

    def set_particles_device(self, src: torch.Tensor):
        """Same from a device (or pinned host) tensor, on the current stream."""
        ptr, n = self.pf.device_ptr()
        assert src.numel() == 7 * n
        dst = _as_tensor(ptr, n, self.device)
        dst.copy_(src.view(-1), non_blocking=True)

    def particles_device(self) -> torch.Tensor:
        """Zero-copy view of p_ (valid until the next resample: the block ping-pongs inside the exchange region)."""
        ptr, n = self.pf.device_ptr()
        return _as_tensor(ptr, n, self.device).view(n, 7)

    def set_cloud(self, cloud):
        self.pf.set_cloud(cloud)

    def set_beacons(self, beacons, alpha, sigma):
        self.beacons, self.alpha, self.sigma = beacons, alpha, sigma
        self.wr_local = torch.zeros(self.hi - self.lo, dtype=torch.float32, device=self.device)

    def set_async(self, enable: bool):
        self.pf.set_async(enable)

    def update(self):
        if self.wr_local is not None:
            self.pf.update_split(self.beacons, self.sigma, self.wr_local.data_ptr())
        else:
            self.pf.update()

    def resample(self, u01: float, want_idx=False):
        if self.wr_local is not None:
            return self.pf.shard_resample(u01, self.wr_local.data_ptr(), self.alpha, want_idx=want_idx)
        return self.pf.shard_resample(u01, want_idx=want_idx)

    def resample_fast(self, u01: float, want_idx=False):
        """NON-exact mode: scalar exchanges + source-side placement (amcl3d_b200_pf_shard_resample_fast)."""
        assert self.wr_local is None, "the fast mode has no range-beacon fusion"
        return self.pf.shard_resample_fast(u01, want_idx=want_idx)

    def mean_device(self) -> torch.Tensor:
        self.pf.shard_mean(self._mean_buf.data_ptr(), fetch=False)
        return self._mean_buf[:7]

    def mean(self):
        return self.pf.shard_mean()

    def launch_count(self):
        return self.pf.launch_count()


def _as_tensor(ptr: int, n_particles: int, device: torch.device) -> torch.Tensor:
    """float32 view of 7 * n floats of library-owned device memory (CUDA array interface, no copy)."""

    class _Mem:
        pass

    m = _Mem()
    m.__cuda_array_interface__ = {"shape": (7 * n_particles,), "typestr": "<f4", "data": (ptr, False), "version": 3,
                                  "strides": None}
    return torch.as_tensor(m, device=device)
