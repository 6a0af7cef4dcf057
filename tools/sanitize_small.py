"""Small pass over the round-2 kernels for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from amcl3d_test_b200 import Grid3d, ParticleFilter, synth  # noqa: E402

sm = synth.make_map((6.3, 5.1, 3.3), 0.1)  # odd sizes: partial tiles in every pass of the fused build
g = Grid3d(0)
g.computePointCloud(sm.points, 0.1)
g.computeGrid(sm.points, 0.05, sub=2)                 # xpass8 / ypass_dpx / zpass_dpx / sparse_patch
g.computeGrid(sm.points, 0.05, sub=2, keep_edt=True)  # exact envelope passes + coder
g.computeGrid(sm.points, 0.05, sub=2)
cloud = synth.make_cloud(sm, 700, 2.5)
P = synth.make_particles(sm, 4099, "T")               # ragged last tile, large-set path
pf = ParticleFilter(g)
pf.p_ = P
pf.update(cloud)                                      # pose + sweep + ordered sum
pf.particleNormalize()
pf.resample(0.37)
print("mean exact", pf.getMean(exact=True))           # TMA ring + product warps
sh = ParticleFilter(g)
sh.shard_init(P.shape[0], 0, 1)
sh.p_ = P
sh.update(cloud)
sh.shard_normalize()
sh.shard_resample(0.37, normalize=False)
sh.p_ = P
sh.update(cloud)
idx = sh.shard_resample_fast(0.37, want_idx=True)     # publish_scalar / combine_scalar / shard_push
print("fast idx ok", bool(np.all(np.diff(idx) >= 0)), sh.shard_mean())
r = g.clone(0)
q = ParticleFilter(r)
q.p_ = P[:300]
q.update(cloud)                                       # single-kernel path on a replica
print("sanitize_small done")
