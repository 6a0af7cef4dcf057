"""One workload's update on its own, for ncu captures: python tools/update_only.py [workload=C3] [kind=T] [reps=3]
Builds the workload's grid, sets the particle set and runs update() `reps` times (the first one also counts in-bounds)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from amcl3d_test_b200 import Grid3d, ParticleFilter, synth  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "C3"
kind = sys.argv[2] if len(sys.argv) > 2 else "T"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = synth.CONFIGS[workload]
sm = synth.make_map(cfg["extent"], cfg["resol"])
g = Grid3d(0)
g.computePointCloud(sm.points, cfg["resol"])
g.computeGrid(sm.points, 0.05, sub=2)
cloud = synth.make_cloud(sm, cfg["points"], cfg["radius"])
n = min(cfg["particles"], int(sys.argv[4]) if len(sys.argv) > 4 else cfg["particles"])
pf = ParticleFilter(g)
pf.p_ = synth.make_particles(sm, n, kind)
pf.set_cloud(cloud)
n_in = pf.update(count=True)
for _ in range(reps):
    pf.update()
    print(f"{workload} {kind} {n}x{cloud.shape[0]}: update {pf.last_kernel_ms():.3f} ms, phases {pf.last_phase_ms()}, in-bounds {n_in}", flush=True)
