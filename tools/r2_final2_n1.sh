#!/bin/bash
# Final single-GPU evidence of round 2 (after the capacity shards / event-ordered group calls): tests, both bench arms, C2.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2_f2_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2_f2_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_f2_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2_f2_smoke.log
python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_f2_ref.json 2> $O/r2_f2_ref.err
( time python bench.py --steps 20 --warmup 5 ) > $O/r2_f2_n1.json 2> $O/r2_f2_n1.err
python bench.py --workload C2 > $O/r2_f2_c2.json 2> $O/r2_f2_c2.err
tail -3 $O/r2_f2_tests.log; tail -2 $O/r2_f2_smoke.log; tail -4 $O/r2_f2_n1.err
