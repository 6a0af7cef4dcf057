"""Probe: how fast is the update kernel when every particle only sweeps a short chunk of the cloud
(emulates perfect lock-step across CTAs)?  python tools/chunk_probe.py"""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amcl3d_test_b200 import Grid3d, ParticleFilter, synth

extent, resol = (100.0, 100.0, 20.0), 0.05
sm = synth.make_map(extent, resol)
g = Grid3d(0); g.computePointCloud(sm.points, resol); g.computeGrid(sm.points, 0.05, sub=2)
print(g.info())
cloud = synth.make_cloud(sm, 10000, 30.0)
pf = ParticleFilter(g)
for kind in ("T", "G"):
    P = synth.make_particles(sm, 100000, kind)
    pf.p_ = P
    for start in (0, 5000):
        for n in (128, 256, 512, 1024, 2048, 10000):
            c = cloud[start:start + n] if n < 10000 else cloud
            pf.set_cloud(c)
            n_in = pf.update(count=True)
            best = min((pf.update(), pf.last_kernel_ms())[1] for _ in range(4))
            print(f"{kind} start {start} chunk {c.shape[0]:6d}: {best:8.3f} ms  {100000*c.shape[0]/best/1e6:7.1f} G evals/s  inb {n_in/(100000*c.shape[0]):.2f}")
