#!/bin/bash
# Validation of a kernel change: GPU tests, the small odd-size script (compute-sanitizer is closed on this pool: plain run),
# the C2 and default bench lines.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r2_v_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2_v_tests.log
timeout 600 python tools/sanitize_small.py > $O/r2_v_san.log 2>&1; echo "small-case rc=$?" >> $O/r2_v_san.log
python bench.py --workload C2 > $O/r2_v_c2.json 2> $O/r2_v_c2.err
( time python bench.py ) > $O/r2_v_n1.json 2> $O/r2_v_n1.err
tail -3 $O/r2_v_tests.log; tail -3 $O/r2_v_san.log
