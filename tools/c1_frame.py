"""C1 frame (600 particles x 2000 points) timing, per stage: python tools/c1_frame.py"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from amcl3d_test_b200 import Grid3d, ParticleFilter, synth  # noqa: E402

for n_p, n_pts in ((600, 2000), (2000, 2000), (600, 10000)):
    sm = synth.make_map((20.0, 20.0, 5.0), 0.1)
    cloud = synth.make_cloud(sm, n_pts, 12.0)
    P = synth.make_particles(sm, n_p, "T")
    g = Grid3d(0)
    g.computePointCloud(sm.points, 0.1)
    g.computeGrid(sm.points, 0.05, sub=2)
    pf = ParticleFilter(g)
    pf.p_ = P
    pf.update(cloud)
    ref = pf.p_.copy()
    t = {k: 0.0 for k in ("set", "update", "update_kernels_ms", "norm", "resample", "read", "mean")}
    reps = 300
    for _ in range(20 + reps):
        a = time.perf_counter(); pf.p_ = P
        b = time.perf_counter(); pf.update(cloud); uk = pf.last_kernel_ms()
        c = time.perf_counter(); pf.particleNormalize()
        d = time.perf_counter(); pf.resample(0.37)
        e = time.perf_counter(); out = pf.p_
        f = time.perf_counter(); pf.getMean()
        h = time.perf_counter()
        if _ >= 20:
            for k, v in zip(t, (b - a, c - b, uk * 1e-3, d - c, e - d, f - e, h - f)):
                t[k] += v
    print(f"{n_p} x {n_pts}: frame {sum(v for k, v in t.items() if k != 'update_kernels_ms') / reps * 1e3:.3f} ms; " +
          ", ".join(f"{k} {v / reps * 1e3:.3f}" for k, v in t.items()), flush=True)
