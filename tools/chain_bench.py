"""Times the sequential-f32 chain (inclusive prefix of the weights) in its three evaluation modes.
   python tools/chain_bench.py [n ...]"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from amcl3d_test_b200 import api, capi, synth  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [1 << 20, 1 << 23]
    sm = synth.make_map((8.0, 8.0, 4.0), 0.1)
    g = api.Grid3d(0)
    g.computePointCloud(sm.points, 0.1)
    g.computeGrid(sm.points, 0.05, mode=api.GRID_SPLAT)
    L = capi.load()
    for n in sizes:
        rng = np.random.RandomState(1)
        P = np.zeros((n, 7), np.float32)
        P[:, :3] = np.array([4, 4, 1.0], np.float32)
        for name, w in (("gamma", rng.gamma(2.0, 1.0, n)), ("equal", np.full(n, 1.0 / n)),
                        ("half_zero", np.where(rng.rand(n) < 0.5, 0, rng.rand(n)))):
            P[:, 6] = w.astype(np.float32)
            pf = api.ParticleFilter(g)
            pf.p_ = P
            ref = None
            for mode in (0, 2, 1):
                pf.set_chain_mode(mode)
                ms = []
                for it in range(6):
                    dp = C.c_void_p()
                    api.check(L.amcl3d_b200_pf_prefix(pf._h, C.byref(dp), None, 0))
                    ms.append(pf.last_kernel_ms())
                out = pf.prefix()
                if ref is None:
                    ref = out
                same = np.array_equal(out.view(np.uint32), ref.view(np.uint32))
                print(f"n={n:>9} {name:<10} mode={mode} prefix {np.median(ms[1:]) * 1e3:8.1f} us  identical={same}", flush=True)
            pf.set_chain_mode(0)


if __name__ == "__main__":
    main()
