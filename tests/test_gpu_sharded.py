"""Particles sharded over GPUs through the C-ABI peer-memory exchange (csrc/shard.cu): the sharded result
must equal the one-GPU result bit for bit -- weights, resampled indices and particles (ParticleFilter.cpp:168-196,
230-254), several frames in a row (the exchange region ping-pongs by epoch).

  * world 1 through the sharded entry points: runs on any box with one GPU;
  * several ranks inside ONE process (amcl3d_b200_pf_shard_attach_local, the path the C++ classes use): two and three
    ranks on device 0 (any box) and two devices (needs 2 GPUs);
  * one process per GPU under torch.distributed.run at 2 ranks (CUDA IPC): needs 2 GPUs; tools/sharded_check.py.
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _device_count():
    from amcl3d_test_b200 import api

    return api.device_info(0)["device_count"]


def _world(device, resol=0.1):
    from amcl3d_test_b200 import Grid3d, synth

    sm = synth.make_map((20.0, 20.0, 5.0), resol)
    g = Grid3d(device)
    g.computePointCloud(sm.points, resol)
    g.computeGrid(sm.points, 0.05, sub=2)
    return sm, g


def _one_gpu(g, P, clouds, u01s, beacons=None):
    from amcl3d_test_b200 import ParticleFilter

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
        out.append((W, pf.p_, idx, pf.getMean(), pf.getMean(exact=True)))
    return out


@pytest.mark.parametrize("with_beacons", [False, True])
def test_world_of_one_equals_the_plain_handle(O, with_beacons):
    import torch

    from amcl3d_test_b200 import synth
    from amcl3d_test_b200.sharded import PeerShardedFilter

    sm, g = _world(0)
    clouds = [synth.make_cloud(sm, 1500, 9.0, seed=synth.SEED_CLOUD + k) for k in range(3)]
    u01s = [0.37, 0.0, 0.93]
    P = synth.make_particles(sm, 10_007, "G")
    beacons = synth.make_beacons(sm, 8) if with_beacons else None
    ref = _one_gpu(g, P, clouds, u01s, beacons)
    f = PeerShardedFilter(g, P.shape[0], torch.device("cuda", 0))
    assert (f.lo, f.hi) == (0, P.shape[0])
    f.set_particles(P)
    if with_beacons:
        f.set_beacons(beacons, 0.5, 0.53)
    for k, (cloud, u01) in enumerate(zip(clouds, u01s)):
        f.set_cloud(cloud)
        f.update()
        idx = f.resample(u01, want_idx=True)
        assert np.array_equal(idx, ref[k][2])
        assert np.array_equal(bits(f.pf.p_), bits(ref[k][1]))
        mean = f.mean()
        assert np.allclose(mean[:6], ref[k][3][:6], rtol=1e-5, atol=1e-6) and mean[6] == ref[k][3][6]
    # first frame also against the CPU oracle (indices are the criterion of north_star)
    mn, mx = O.compute_bounds(sm.points)
    m = O.Map(mn, mx, 0.1, 0.05)
    m.set_prob(g.downloadGrid())
    if not with_beacons:
        W = O.update(m, O.particles(P), clouds[0])
        O.normalize(m, W)
        _, oidx = O.resample(W, np.float32(u01s[0]))
        assert np.array_equal(ref[0][2], oidx)


def test_fast_mode_in_a_world_of_one_is_the_exact_result():
    """With one rank the fast mode's rank-order combination is the sequential chain itself: same indices, same bits."""
    import torch

    from amcl3d_test_b200 import synth
    from amcl3d_test_b200.sharded import PeerShardedFilter

    sm, g = _world(0)
    cloud = synth.make_cloud(sm, 1500, 9.0)
    P = synth.make_particles(sm, 30_011, "G")
    ref = _one_gpu(g, P, [cloud, cloud], [0.37, 0.81])
    f = PeerShardedFilter(g, P.shape[0], torch.device("cuda", 0))
    f.set_particles(P)
    for k, u01 in enumerate((0.37, 0.81)):
        f.set_cloud(cloud)
        f.update()
        idx = f.resample_fast(u01, want_idx=True)
        assert np.array_equal(idx, ref[k][2])
        assert np.array_equal(bits(f.pf.p_), bits(ref[k][1]))


# rank layouts of the in-process tests: device ordinal of every rank.  Ranks that share a device run the very same exchange
# protocol (stores into, flags on and loads from the other rank's region) -- so the protocol is covered on a one-GPU box too.
LAYOUTS = {"2 ranks on device 0": [0, 0], "3 ranks on device 0": [0, 0, 0], "2 devices": [0, 1]}


def _layout(name):
    devs = LAYOUTS[name]
    if max(devs) + 1 > _device_count():
        pytest.skip(f"needs {max(devs) + 1} GPUs")
    return devs


@pytest.mark.parametrize("layout", list(LAYOUTS))
def test_ranks_in_one_process(layout):
    """The ranks are handles of one process (attach_local + stream order): what amcl3d::ParticleFilter::setDevices uses."""
    devs = _layout(layout)
    from amcl3d_test_b200 import ParticleFilter, synth
    from amcl3d_test_b200.sharded import shard_bounds

    world = len(devs)
    by_dev = {}
    for dv in sorted(set(devs)):
        sm, by_dev[dv] = _world(dv)
    clouds = [synth.make_cloud(sm, 1500, 9.0, seed=synth.SEED_CLOUD + k) for k in range(3)]
    u01s = [0.37, 0.0, 0.93]
    n = 20_001
    P = synth.make_particles(sm, n, "T")
    ref = _one_gpu(by_dev[0], P, clouds, u01s)
    pfs = [ParticleFilter(by_dev[dv]) for dv in devs]
    for r, pf in enumerate(pfs):
        lo, hi = pf.shard_init(n, r, world)
        assert (lo, hi) == shard_bounds(n, world, r)
        pf.p_ = P[lo:hi]
    ParticleFilter.shard_attach_local(pfs)
    for k, (cloud, u01) in enumerate(zip(clouds, u01s)):
        for pf in pfs:
            pf.update(cloud)
        ParticleFilter.shard_resample_local(pfs, u01)
        out = np.concatenate([pf.p_ for pf in pfs])
        assert np.array_equal(bits(out), bits(ref[k][1]))
        mean = ParticleFilter.shard_mean_local(pfs)
        assert np.allclose(mean[:6], ref[k][3][:6], rtol=1e-5, atol=1e-6) and mean[6] == ref[k][3][6]
        exact = ParticleFilter.shard_mean_exact_local(pfs)  # one f32 chain carried from rank to rank
        assert np.array_equal(bits(exact), bits(ref[k][4]))


@pytest.mark.parametrize("layout", list(LAYOUTS))
def test_particle_count_changes_between_sharded_steps(layout):
    """uniformSample changes the particle count every frame (amcl3d.cpp:192-203, ParticleFilter.cpp:256-280) and has no
    sharded form: the set is gathered into a plain handle device to device, resampled there, and scattered back over a
    re-drawn partition (shards with capacity) -- shrinking, then growing.  Bit-identical to one GPU throughout."""
    devs = _layout(layout)
    from amcl3d_test_b200 import ParticleFilter, synth
    from amcl3d_test_b200.sharded import shard_bounds

    world = len(devs)
    by_dev = {}
    for dv in sorted(set(devs)):
        sm, by_dev[dv] = _world(dv)
    clouds = [synth.make_cloud(sm, 1200, 9.0, seed=synth.SEED_CLOUD + k) for k in range(3)]
    n, cap = 6_001, 9_000
    counts = [4_099, 8_999]  # shrink, then grow to just below the capacity
    u01s = [0.11, 0.77]
    P = synth.make_particles(sm, n, "T")

    one = ParticleFilter(by_dev[0])
    one.p_ = P
    one.update(clouds[0])
    ref = []
    for S, cloud, u01 in zip(counts, clouds[1:], u01s):
        one.uniformSample(S)
        after_sample = one.p_
        one.update(cloud)
        one.particleNormalize()
        one.resample(u01)
        ref.append((after_sample, one.p_))

    home = ParticleFilter(by_dev[0])  # the plain handle the set visits
    pfs = [ParticleFilter(by_dev[dv]) for dv in devs]
    for r, pf in enumerate(pfs):
        lo, hi = pf.shard_init_capacity(n, cap, r, world)
        assert (lo, hi) == shard_bounds(n, world, r)
        pf.p_ = P[lo:hi]
    ParticleFilter.shard_attach_local(pfs)
    for pf in pfs:
        pf.update(clouds[0])
    for k, (S, cloud, u01) in enumerate(zip(counts, clouds[1:], u01s)):
        ParticleFilter.shard_gather_local(pfs, home)
        home.uniformSample(S)
        assert np.array_equal(bits(home.p_), bits(ref[k][0]))
        ParticleFilter.shard_scatter_local(home, pfs)
        for r, pf in enumerate(pfs):
            assert pf.shard_info()[:2] == shard_bounds(S, world, r) and pf.size() == shard_bounds(S, world, r)[1] - shard_bounds(S, world, r)[0]
        assert np.array_equal(bits(np.concatenate([pf.p_ for pf in pfs])), bits(ref[k][0]))
        for pf in pfs:
            pf.update(cloud)
        ParticleFilter.shard_resample_local(pfs, u01)
        assert np.array_equal(bits(np.concatenate([pf.p_ for pf in pfs])), bits(ref[k][1]))
    # misuse: beyond the capacity, fewer particles than ranks, a sharded handle as the plain one
    home.p_ = synth.make_particles(sm, cap + 1, "T")
    with pytest.raises(RuntimeError):
        ParticleFilter.shard_scatter_local(home, pfs)
    home.p_ = P[:world - 1]
    with pytest.raises(RuntimeError):
        ParticleFilter.shard_scatter_local(home, pfs)
    with pytest.raises(RuntimeError):
        ParticleFilter.shard_gather_local(pfs, pfs[0])
    with pytest.raises(RuntimeError):
        pfs[0].shard_resize(cap + 1)


def test_two_ranks_one_process_per_gpu():
    """torch.distributed.run at 2 ranks: CUDA IPC handles, peer stores/loads over NVLink; tools/sharded_check.py."""
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(ROOT / "tools" / "sharded_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "OK bit-identical to 1 GPU" in r.stdout


def test_grid_clone_gives_the_same_weights(O):
    """amcl3d_b200_grid_clone: a replica (coded cells only, and with the float cells) drives update() to the same bits; a
    codes-only replica refuses to hand out float cells it does not hold."""
    from amcl3d_test_b200 import ParticleFilter, synth
    from amcl3d_test_b200.capi import check  # noqa: F401

    sm, g = _world(0)
    cloud = synth.make_cloud(sm, 1500, 9.0)
    P = synth.make_particles(sm, 9000, "T")
    pf = ParticleFilter(g)
    pf.p_ = P
    pf.update(cloud)
    want = pf.p_
    dev = 1 if _device_count() >= 2 else 0
    for with_float in (False, True):
        r = g.clone(dev, with_float=with_float)
        assert r.info()["size"] == g.info()["size"] and r.info()["lut_n"] == g.info()["lut_n"] and r.info()["device"] == dev
        q = ParticleFilter(r)
        q.p_ = P
        q.update(cloud)
        assert np.array_equal(bits(q.p_), bits(want))
        if with_float:
            assert np.array_equal(bits(r.downloadGrid()), bits(g.downloadGrid()))
        else:
            with pytest.raises(RuntimeError):
                r.downloadGrid()


def test_shard_calls_fail_loudly_when_misused():
    """Bad geometry, calls before the peers are attached, a block that does not match the partition."""
    import ctypes as C

    from amcl3d_test_b200 import ParticleFilter, capi, synth

    sm, g = _world(0)
    L = capi.load()
    pf = ParticleFilter(g)
    assert L.amcl3d_b200_pf_shard_init(pf._h, 0, 0, 1) == capi.ERR_ARG          # no particles
    assert L.amcl3d_b200_pf_shard_init(pf._h, 100, 3, 2) == capi.ERR_ARG        # rank outside the world
    assert L.amcl3d_b200_pf_shard_init(pf._h, 100, 0, 17) == capi.ERR_ARG       # more ranks than the exchange region holds
    assert L.amcl3d_b200_pf_shard_resample(pf._h, C.c_float(0.5), 1, None, C.c_float(1.0), None) == capi.ERR_STATE  # not sharded
    assert L.amcl3d_b200_pf_shard_init(pf._h, 1000, 0, 2) == capi.OK
    assert L.amcl3d_b200_pf_shard_init(pf._h, 1000, 0, 2) == capi.ERR_STATE     # twice
    assert L.amcl3d_b200_pf_shard_resample(pf._h, C.c_float(0.5), 1, None, C.c_float(1.0), None) == capi.ERR_STATE  # peers not attached
    assert L.amcl3d_b200_pf_shard_mean(pf._h, None, None) == capi.ERR_ARG
    lo, hi, nbytes = pf.shard_info()
    assert (lo, hi) == (0, 500) and nbytes > 2 * 500 * 28 + 2 * 1000 * 4
    assert L.amcl3d_b200_pf_bind_device(pf._h, C.c_void_p(256), 1, 1) == capi.ERR_STATE  # p_ lives in the exchange region
    assert len(pf.shard_ipc_handle()) == 64


def test_host_pin_roundtrip():
    import ctypes as C

    from amcl3d_test_b200 import capi

    L = capi.load()
    buf = np.zeros(1 << 20, np.uint8)
    assert L.amcl3d_b200_host_pin(C.c_void_p(buf.ctypes.data), buf.nbytes) == capi.OK
    assert L.amcl3d_b200_host_pin(C.c_void_p(buf.ctypes.data), buf.nbytes) == capi.OK   # registering again re-registers
    assert L.amcl3d_b200_host_unpin(C.c_void_p(buf.ctypes.data)) == capi.OK
    assert L.amcl3d_b200_host_unpin(C.c_void_p(buf.ctypes.data)) != capi.OK             # nothing left to release
    assert L.amcl3d_b200_host_pin(None, 16) == capi.ERR_ARG
