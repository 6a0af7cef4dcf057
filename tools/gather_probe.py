"""Does any L2 fetch granularity or load flavour make B200 fetch less than a 128-byte line per missing 4-byte gather?
Runs the library's gather probe (random addresses over the whole C3-small float grid: every load misses L2) at
cudaLimitMaxL2FetchGranularity 32 / 64 / 128 and with the load flavours of gather_bench_kernel; run it under
  ncu --metrics dram__sectors_read.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,lts__t_sectors_srcunit_tex.sum,gpu__time_duration.sum -k regex:gather_bench
and read sectors per miss from the launch list (order of launches = order of the lines printed here)."""
import ctypes as C
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from amcl3d_test_b200 import Grid3d, capi, synth  # noqa: E402

L = capi.load()
sm = synth.make_map((50.0, 50.0, 10.0), 0.05)  # 1000 x 1000 x 200 voxels = 800 MB of float cells >> L2
g = Grid3d(0)
g.computePointCloud(sm.points, 0.05)
g.computeGrid(sm.points, 0.05, sub=2)
out = []
names = ["ldg.nc", "ld.ca", "ld.cg", "ld.cv", "nc.L1::no_allocate", "ld.L1::no_allocate", "L1::evict_first", "ld.cs"]
for gran in (128, 64, 32):
    rc = L.amcl3d_b200_set_l2_fetch_granularity(0, gran)
    cur = C.c_int(0)
    L.amcl3d_b200_get_l2_fetch_granularity(0, C.byref(cur))
    for flavor in ((0, 2, 4) if gran != 32 else range(8)):
        ms, loads = g.bench_gather(window_cells=0, blocks=148 * 8, iters=64, coherent=flavor << 1)
        rec = {"l2_fetch_granularity_set": gran, "reported": cur.value, "flavour": names[flavor], "loads": loads,
               "ms": ms, "loads_per_sec": loads / (ms * 1e-3)}
        out.append(rec)
        print(json.dumps(rec), flush=True)
# L2-resident window (8 MB) for contrast: sectors per load when everything hits
for coh in (0, 1):
    ms, loads = g.bench_gather(window_cells=8 << 18, blocks=148 * 8, iters=64, coherent=coh)
    print(json.dumps({"window": "8 MB (L2-resident)", "coherent": coh, "loads": loads, "ms": ms, "loads_per_sec": loads / (ms * 1e-3)}), flush=True)
for coh in (0, 1):
    ms, loads = g.bench_gather(window_cells=0, blocks=148 * 8, iters=64, coherent=coh)
    print(json.dumps({"window": "whole grid", "coherent": coh, "loads": loads, "ms": ms, "loads_per_sec": loads / (ms * 1e-3)}), flush=True)
