#!/bin/bash
# run ON the GPU box: tools/sweep_defs.sh "<nvcc -D flags>" ... ; rebuilds update.cu and runs bench (C3) T and G
cd "$(dirname "$0")/.."
for defs in "$@"; do
  export AMCL3D_NVCC_DEFS="$defs"
  touch amcl3d_test_b200/csrc/update.cu
  python amcl3d_test_b200/build.py -v 2>&1 | grep -E "update_kernelILi1" -A2 | grep -E "Used" | head -1
  for kind in T G; do
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --particles $kind ${BENCH_ARGS:-} 2>/dev/null > /tmp/sw.json
    python -c "
import json
d=json.load(open('/tmp/sw.json')); print('[$defs] $kind update_ms %.3f  step_ms %.3f  value %.1f G evals/s' % (d['update_ms'], d['ms_per_step'], d['value']/1e9))"
  done
done
