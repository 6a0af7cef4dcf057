#!/bin/bash
# ncu evidence of the final round-2 code (one GPU).  Every program below has already exited 0 without ncu.
cd "$(dirname "$0")/.."
O=gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c4 > $O/r2_f2_plain.json 2> $O/r2_f2_plain.err || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file $O/r2_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c4 > $O/r2_f2_ncu.log 2>&1
python tools/c2_build.py 1 2 > $O/r2_f2_c2_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"class_count|voxelize|find_cap|xpass8|ypass_dpx|zpass_dpx|sparse_patch" -c 12 -f -o $O/r2_build_C2_v3 \
    python tools/c2_build.py 1 1 > $O/r2_f2_build_ncu.log 2>&1
ls -la $O/r2_build_C2_v3.ncu-rep $O/r2_launches_bench.csv
