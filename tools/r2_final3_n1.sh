#!/bin/bash
# Last single-GPU check of the round: GPU tests, smoke, the default bench line.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2_f3_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2_f3_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_f3_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2_f3_smoke.log
python bench.py --steps 20 --warmup 5 > $O/r2_f3_n1.json 2> $O/r2_f3_n1.err; echo "bench rc=$?" >> $O/r2_f3_n1.err
tail -3 $O/r2_f3_tests.log; tail -2 $O/r2_f3_smoke.log; tail -2 $O/r2_f3_n1.err
