"""Development micro-benchmarks (not the contract bench): gather roofline probes + update timing."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from amcl3d_test_b200 import Grid3d, ParticleFilter, device_info, synth  # noqa: E402


def main():
    extent = tuple(float(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (50, 50, 10)))
    resol = float(sys.argv[4]) if len(sys.argv) > 4 else 0.05
    n_p = int(sys.argv[5]) if len(sys.argv) > 5 else 100_000
    n_pts = int(sys.argv[6]) if len(sys.argv) > 6 else 10_000
    print("device", device_info(0))
    t = time.time()
    sm = synth.make_map(extent, resol)
    print(f"map {sm.size} = {sm.n_voxels/1e6:.1f} Mvox, {sm.points.shape[0]/1e6:.2f} M points, host gen {time.time()-t:.1f}s")
    g = Grid3d(0)
    g.computePointCloud(sm.points, resol)
    t = time.time()
    g.computeGrid(sm.points, 0.05, sub=2)
    print(f"grid build (incl. H2D of points) {time.time()-t:.3f}s")
    out = {}
    import ctypes as C
    from amcl3d_test_b200 import capi
    L = capi.load()
    cur = C.c_int(0); L.amcl3d_b200_get_l2_fetch_granularity(0, C.byref(cur)); print("default L2 fetch granularity", cur.value)
    for gran in (128, 64, 32):
        print("set granularity", gran, "rc", L.amcl3d_b200_set_l2_fetch_granularity(0, gran))
        ms, loads = g.bench_gather(window_cells=0, blocks=148 * 8, iters=256, coherent=False)
        print(f"  random gather over whole grid: {loads/ms/1e6:.1f} G loads/s")
        ms, loads = g.bench_gather(window_cells=0, blocks=148 * 8, iters=256, coherent=True)
        print(f"  coherent gather over whole grid: {loads/ms/1e6:.1f} G loads/s")
    for win_mb, k in ((64, 1), (512, 16), (1024, 32), (2048, 64), (4096, 128), (0, 16), (0, 32), (0, 64)):
        cells = win_mb * (1 << 20) // 4
        ms, loads = g.bench_gather(window_cells=cells, blocks=148 * 8, iters=256, coherent=(max(k, 0 if k == 1 else 16) << 1) if k > 1 else 0)
        print(f"sparse probe window {win_mb or 'all'} MB, 1 line in {k}: {loads/ms/1e6:.1f} G loads/s")
    for flavor, name in enumerate(["ldg.nc", "ld.ca", "ld.cg", "ld.cv", "nc.noalloc", "ld.noalloc", "evict_first", "ld.cs"]):
        ms, loads = g.bench_gather(window_cells=0, blocks=148 * 8, iters=256, coherent=flavor << 1)
        ms2, loads2 = g.bench_gather(window_cells=0, blocks=148 * 8, iters=256, coherent=(flavor << 1) | 1)
        print(f"flavor {name}: random {loads/ms/1e6:.1f} G/s, coherent {loads2/ms2/1e6:.1f} G/s")
    for win_mb in (8, 0):
        cells = win_mb * (1 << 20) // 4
        for coh in (0, 1):
            ms, loads = g.bench_gather(window_cells=cells, blocks=148 * 8, iters=256, coherent=bool(coh))
            rate = loads / (ms * 1e-3)
            key = f"gather_win{win_mb or 'all'}MB_{'coherent' if coh else 'random'}"
            out[key] = rate
            print(f"{key}: {rate/1e9:.1f} G loads/s = {rate*32/1e9:.0f} GB/s sector traffic")
    cloud = synth.make_cloud(sm, n_pts, 30.0 if min(extent[:2]) >= 60 else min(extent[:2]) / 2.2)
    pf = ParticleFilter(g)
    for kind in ("T", "G"):
        P = synth.make_particles(sm, n_p, kind)
        pf.p_ = P
        pf.set_cloud(cloud)
        n_in = pf.update(count=True)
        best = 1e9
        for _ in range(5):
            pf.update()
            best = min(best, pf.last_kernel_ms())
        evals = n_p * n_pts
        print(f"update {kind}: {n_p} x {n_pts}: {best:.3f} ms -> {evals/best/1e6:.1f} G evals/s, in-bounds {n_in/evals:.3f}, "
              f"sector GB/s {n_in*32/best/1e6:.0f}")
        out[f"update_{kind}_ms"] = best
        out[f"update_{kind}_inb"] = n_in / evals
        pf.particleNormalize()
        t_norm = pf.last_kernel_ms()
        pf.resample(0.37)
        t_res = pf.last_kernel_ms()
        pf.p_ = P
        pf.update()
        pf.uniformSample(n_p)
        t_us = pf.last_kernel_ms()
        m = pf.getMean(exact=True)
        t_me = pf.last_kernel_ms()
        pf.getMean(exact=False)
        t_mf = pf.last_kernel_ms()
        print(f"  normalize {t_norm:.3f} ms, resample {t_res:.3f} ms, uniformSample {t_us:.3f} ms, mean exact {t_me:.3f} / fast {t_mf:.3f} ms")
    Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / "microbench.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
