"""The L2 prefetch-size qualifiers of ld.global (.L2::64B / .L2::128B / .L2::256B) on a gather that always misses L2: do they
change the DRAM bytes fetched per missing 4-byte load?  Same probe as tools/gather_probe.py (flavours 8-11 of
gather_bench_kernel); run plain for the rates and under
  ncu --metrics dram__sectors_read.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,gpu__time_duration.sum -k regex:gather_bench
for the bytes (launch order = order of the lines printed here, one warm-up + one timed launch each)."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from amcl3d_test_b200 import Grid3d, synth  # noqa: E402

sm = synth.make_map((50.0, 50.0, 10.0), 0.05)  # 1000 x 1000 x 200 voxels = 800 MB of float cells >> L2
g = Grid3d(0)
g.computePointCloud(sm.points, 0.05)
g.computeGrid(sm.points, 0.05, sub=2)
names = {0: "ldg.nc (default)", 8: "ld.global.L2::64B", 9: "ld.global.L2::128B", 10: "ld.global.L2::256B",
         11: "ld.global.nc.L1::no_allocate.L2::64B"}
for flavor, name in names.items():
    ms, loads = g.bench_gather(window_cells=0, blocks=148 * 8, iters=64, coherent=flavor << 1)
    print(json.dumps({"flavour": name, "loads": loads, "ms": ms, "loads_per_sec": loads / (ms * 1e-3)}), flush=True)
