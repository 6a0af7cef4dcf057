"""Seeded synthetic inputs for the amcl3d weighting path.

Everything here is host-side numpy set-up code (never timed): a "multi-floor office" occupancy map
with occupied points at voxel centres, a sensor-frame cloud cut out of it, tracking / global particle
sets, the resample offset and UWB beacons.  Seeds: map 1001, cloud 2002, particles 3003, offset 4004,
beacons 5005.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SEED_MAP, SEED_CLOUD, SEED_PARTICLES, SEED_OFFSET, SEED_BEACONS = 1001, 2002, 3003, 4004, 5005

# named shapes from BASELINE.json `configs`
CONFIGS = {
    "C1": dict(extent=(20.0, 20.0, 5.0), resol=0.1, particles=600, points=2000, radius=12.0),
    "C2": dict(extent=(100.0, 100.0, 20.0), resol=0.05, particles=0, points=0, radius=30.0),
    "C3": dict(extent=(100.0, 100.0, 20.0), resol=0.05, particles=100_000, points=10_000, radius=30.0),
    "C4": dict(extent=(100.0, 100.0, 20.0), resol=0.05, particles=1_000_000, points=20_000, radius=30.0),
    "C5": dict(extent=(400.0, 400.0, 40.0), resol=0.1, particles=1_000_000, points=20_000, radius=30.0),
    "C5s": dict(extent=(400.0, 400.0, 4.0), resol=0.1, particles=1_000_000, points=20_000, radius=30.0),
}


@dataclass
class SynthMap:
    points: np.ndarray  # (P,3) float32 map points (voxel centres + the two extreme corners)
    extent: tuple
    resol: float
    size: tuple  # voxels per axis

    @property
    def n_voxels(self) -> int:
        return int(self.size[0]) * int(self.size[1]) * int(self.size[2])


def _occupied_voxels(size, resol, seed):
    """int32 (K,3) occupied voxel indices of the office map."""
    sx, sy, sz = size
    per_m = int(round(1.0 / resol))
    parts = []

    def grid3(xs, ys, zs):
        g = np.stack(np.meshgrid(xs, ys, zs, indexing="ij"), axis=-1).reshape(-1, 3)
        return g.astype(np.int32)

    ax, ay, az = np.arange(sx), np.arange(sy), np.arange(sz)
    # floor slabs every 3 m
    slabs = np.arange(0, sz, 3 * per_m)
    parts.append(grid3(ax, ay, slabs))
    # perimeter walls
    parts.append(grid3(np.array([0, sx - 1]), ay, az))
    parts.append(grid3(ax, np.array([0, sy - 1]), az))
    # interior walls every 5 m with 1 m wide, 2 m high door gaps centred between crossings
    wall = 5 * per_m
    door_lo, door_hi = int(2.0 * per_m), int(3.0 * per_m)  # gap [2 m, 3 m) of every 5 m span
    z_in_door = ((az % (3 * per_m)) >= 1) & ((az % (3 * per_m)) <= 2 * per_m)
    for axis, (n_a, n_b) in enumerate(((sx, sy), (sy, sx))):
        planes = np.arange(wall, n_a - 1, wall)
        along = np.arange(n_b)
        in_gap = ((along % wall) >= door_lo) & ((along % wall) < door_hi)
        solid = np.ones((along.size, az.size), bool)
        solid[np.ix_(in_gap, z_in_door)] = False
        bi, zi = np.nonzero(solid)
        for pl in planes:
            col = np.full(bi.size, pl, np.int32)
            if axis == 0:
                parts.append(np.stack([col, along[bi], az[zi]], 1).astype(np.int32))
            else:
                parts.append(np.stack([along[bi], col, az[zi]], 1).astype(np.int32))
    # Bernoulli(0.002) clutter: draw the count, then uniform voxels (duplicates merge below)
    rng = np.random.RandomState(seed)
    n_vox = sx * sy * sz
    k = int(rng.binomial(n_vox, 0.002)) if n_vox < 2**31 else int(round(n_vox * 0.002))
    clutter = np.stack([rng.randint(0, sx, k), rng.randint(0, sy, k), rng.randint(0, sz, k)], 1).astype(np.int32)
    parts.append(clutter)
    vox = np.concatenate(parts, 0)
    key = (vox[:, 2].astype(np.int64) * sy + vox[:, 1]) * sx + vox[:, 0]
    _, first = np.unique(key, return_index=True)
    return vox[np.sort(first)]


def make_map(extent=(20.0, 20.0, 5.0), resol=0.1, seed=SEED_MAP) -> SynthMap:
    size = tuple(int(round(e / resol)) for e in extent)
    vox = _occupied_voxels(size, resol, seed)
    centres = ((vox.astype(np.float64) + 0.5) * resol).astype(np.float32)
    corners = np.array([[0.0, 0.0, 0.0], list(extent)], np.float32)  # pins getMinMax3D to the nominal box
    pts = np.concatenate([corners, centres], 0)
    return SynthMap(points=np.ascontiguousarray(pts), extent=tuple(extent), resol=resol, size=size)


def gt_pose(extent, resol):
    """Ground truth: map centre, 1.2 m above the slab nearest mid-height, yaw 0.3 rad."""
    slab_z = 3.0 * np.floor((extent[2] / 2.0) / 3.0)
    return np.array([extent[0] / 2.0, extent[1] / 2.0, slab_z + 1.2, 0.0, 0.0, 0.3], np.float32)


def _rot(roll, pitch, yaw):
    sr, cr, sp, cp, sy, cy = np.sin(roll), np.cos(roll), np.sin(pitch), np.cos(pitch), np.sin(yaw), np.cos(yaw)
    return np.array(
        [
            [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
            [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
            [-sp, cp * sr, cp * cr],
        ],
        np.float64,
    )


def make_cloud(smap: SynthMap, n_points: int, radius: float, seed=SEED_CLOUD, leaf=0.1, noise=0.02) -> np.ndarray:
    """(M,3) float32 sensor-frame cloud, ordered like pcl::VoxelGrid output (voxel index, x fastest)."""
    rng = np.random.RandomState(seed)
    gt = gt_pose(smap.extent, smap.resol).astype(np.float64)
    pts = smap.points[2:].astype(np.float64)
    d2 = ((pts - gt[:3]) ** 2).sum(1)
    near = np.nonzero(d2 <= radius * radius)[0]
    if near.size < n_points:
        raise ValueError(f"only {near.size} occupied points within {radius} m, need {n_points}")
    pick = rng.choice(near, n_points, replace=False)
    world = pts[pick]
    R = _rot(gt[3], gt[4], gt[5])
    sensor = (world - gt[:3]) @ R  # R^T (w - t) written row-wise
    sensor = sensor + rng.normal(0.0, noise, sensor.shape)
    v = np.floor(sensor / leaf).astype(np.int64)
    v -= v.min(0)
    dims = v.max(0) + 1
    key = (v[:, 2] * dims[1] + v[:, 1]) * dims[0] + v[:, 0]
    order = np.argsort(key, kind="stable")
    return np.ascontiguousarray(sensor[order].astype(np.float32))


def make_particles(smap: SynthMap, n: int, kind="T", seed=SEED_PARTICLES) -> np.ndarray:
    """(n,7) float32 [x,y,z,roll,pitch,yaw,w].  kind 'T' tracking cluster, 'G' global uniform."""
    rng = np.random.RandomState(seed)
    p = np.zeros((n, 7), np.float32)
    if kind == "T":
        gt = gt_pose(smap.extent, smap.resol).astype(np.float64)
        dev = np.array([0.5, 0.5, 0.4, 0.2, 0.2, 0.4])  # ParticleFilter.h:56-61 (pitch typo read as 0.2)
        p[:, :6] = (gt + rng.normal(0.0, 1.0, (n, 6)) * dev).astype(np.float32)
    elif kind == "G":
        ext = np.array(smap.extent)
        lo, hi = -0.01 * ext, 1.01 * ext  # 2 % inflated box: ~4 % of particles fail isIntoMap
        p[:, :3] = (lo + rng.uniform(0.0, 1.0, (n, 3)) * (hi - lo)).astype(np.float32)
        p[:, 5] = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    else:
        raise ValueError(kind)
    p[:, 6] = np.float32(1.0 / max(n, 1))
    return p


def resample_offset(seed=SEED_OFFSET) -> np.float32:
    """The U(0,1) draw that resample() multiplies by 1/N (ParticleFilter.cpp:234)."""
    return np.float32(np.random.RandomState(seed).uniform(0.0, 1.0))


def make_beacons(smap: SynthMap, n=8, seed=SEED_BEACONS, noise=0.1):
    """(pos (n,3) f64, range (n,) f64): anchors uniform in the box, range to GT + N(0, 0.1)."""
    rng = np.random.RandomState(seed)
    pos = rng.uniform(0.0, 1.0, (n, 3)) * np.array(smap.extent)
    gt = gt_pose(smap.extent, smap.resol).astype(np.float64)[:3]
    rng_m = np.sqrt(((pos - gt) ** 2).sum(1)) + rng.normal(0.0, noise, n)
    return pos, rng_m
