"""Probe: update-kernel time when the cloud is evaluated in a spatially coherent (Morton) order."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amcl3d_test_b200 import Grid3d, ParticleFilter, synth


def morton(c, cell):
    q = np.clip(np.floor(c / cell).astype(np.int64) + 2048, 0, 4095)
    key = np.zeros(len(c), np.int64)
    for b in range(12):
        for a in range(3):
            key |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return key


extent, resol = (100.0, 100.0, 20.0), 0.05
sm = synth.make_map(extent, resol)
g = Grid3d(0); g.computePointCloud(sm.points, resol); g.computeGrid(sm.points, 0.05, sub=2)
cloud = synth.make_cloud(sm, 10000, 30.0)
pf = ParticleFilter(g)
rng = np.random.RandomState(0)
orders = {"voxelgrid(original)": np.arange(len(cloud)), "random": rng.permutation(len(cloud))}
for cell in (0.25, 1.0, 4.0):
    orders[f"morton{cell}"] = np.argsort(morton(cloud, cell), kind="stable")
orders["by_range"] = np.argsort(np.linalg.norm(cloud, axis=1))
orders["by_azimuth"] = np.argsort(np.arctan2(cloud[:, 1], cloud[:, 0]))
for kind in ("T", "G"):
    P = synth.make_particles(sm, 100000, kind)
    for name, o in orders.items():
        pf.p_ = P
        pf.set_cloud(cloud[o])
        best = min((pf.update(), pf.last_kernel_ms())[1] for _ in range(4))
        print(f"{kind} {name:22s}: {best:8.3f} ms  {1e9/best/1e6:7.1f} G evals/s")
