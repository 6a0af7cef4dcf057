#!/bin/bash
# Round-2 ncu evidence (one GPU).  Every program below has already exited 0 without ncu.
set -x
cd "$(dirname "$0")/.."
O=gpurun_out
python tools/gather_probe.py > $O/r2_gather_probe_plain.log 2>&1
ncu --metrics dram__sectors_read.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,lts__t_sectors_srcunit_tex.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum \
    --clock-control none -k regex:gather_bench --csv --log-file $O/r2_gather_probe_ncu.csv python tools/gather_probe.py > $O/r2_gather_probe_ncu.log 2>&1
python tools/update_only.py C3 T 3 > $O/r2_update_T_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pose_kernel|sweep_kernel|ordered_sum_kernel" --launch-skip 3 -c 6 -f -o $O/r2_update_T \
    python tools/update_only.py C3 T 3 > $O/r2_update_T_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pose_kernel|sweep_kernel|ordered_sum_kernel" --launch-skip 3 -c 6 -f -o $O/r2_update_G \
    python tools/update_only.py C3 G 3 > $O/r2_update_G_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"class_count|voxelize|find_cap|xpass8|ypass_dpx|zpass_dpx|sparse_patch" -c 12 -f -o $O/r2_build_C2 \
    python tools/c2_build.py 1 1 > $O/r2_build_ncu.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_b5_plain.json 2> $O/r2_b5_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2_b5_ncu.log 2>&1
ls -la $O/*.ncu-rep
