#!/bin/bash
# Round-2 final single-GPU evidence: tests, bench lines, launch list, build capture.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2_final_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2_final_tests.log
python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_final_ref.json 2> $O/r2_final_ref.err
python bench.py --steps 20 --warmup 5 > $O/r2_final_n1.json 2> $O/r2_final_n1.err
python bench.py --workload C2 > $O/r2_final_c2.json 2> $O/r2_final_c2.err
python bench.py --workload C4 --steps 10 --no-cpu-baseline > $O/r2_final_c4.json 2> $O/r2_final_c4.err
python bench.py --workload C5s --steps 10 --no-cpu-baseline > $O/r2_final_c5s.json 2> $O/r2_final_c5s.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c4 > $O/r2_final_plain.json 2> $O/r2_final_plain.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file $O/r2_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c4 > $O/r2_final_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"class_count|voxelize|find_cap|xpass8|ypass_dpx|zpass_dpx|sparse_patch" -c 12 -f -o $O/r2_build_C2_v2 \
    python tools/c2_build.py 1 1 > $O/r2_build_ncu2.log 2>&1
tail -3 $O/r2_final_tests.log
