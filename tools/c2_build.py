"""Config-2 grid build on its own (for ncu launch lists / full captures): python tools/c2_build.py [fast=1] [reps=2]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from amcl3d_test_b200 import Grid3d, synth  # noqa: E402

fast = int(sys.argv[1]) if len(sys.argv) > 1 else 1
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sm = synth.make_map((100.0, 100.0, 20.0), 0.05)
g = Grid3d(0)
g.computePointCloud(sm.points, 0.05)
g.set_edt_fast(bool(fast))
for _ in range(reps):
    t = time.time()
    g.computeGrid(sm.points, 0.05, sub=2)
    print(f"fast={fast} wall {time.time() - t:.3f} s, kernels {g.last_build_ms():.2f} ms, lut_n {g.info()['lut_n']}", flush=True)
