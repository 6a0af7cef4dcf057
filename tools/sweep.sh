#!/bin/bash
# tuning sweep, run ON the GPU box: rebuilds update.cu with different knobs and times the update kernel
# usage: tools/sweep.sh "THREADS TILE UNROLL MINBLOCKS" ...
cd "$(dirname "$0")/.."
for cfg in "$@"; do
  set -- $cfg
  export AMCL3D_NVCC_DEFS="-DAMCL3D_THREADS=$1 -DAMCL3D_TILE=$2 -DAMCL3D_UNROLL=$3 -DAMCL3D_MINBLOCKS=$4"
  touch amcl3d_test_b200/csrc/update.cu
  python amcl3d_test_b200/build.py -v 2>&1 | grep -E "update_kernelILi1" -A2 | grep -E "Used|spill" | head -2
  echo "== cfg threads=$1 tile=$2 unroll=$3 minblocks=$4"
  python tools/microbench.py ${SWEEP_ARGS:-} 2>&1 | grep -E "^update"
done
