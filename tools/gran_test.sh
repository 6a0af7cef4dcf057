#!/bin/bash
cd "$(dirname "$0")/.."
for g in 32 64 128; do
  AMCL3D_B200_L2_FETCH=$g python bench.py --steps 5 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} 2>/dev/null > /tmp/g_$g.json
  python -c "
import json
d=json.load(open('/tmp/g_$g.json')); print('gran $g update_ms', d['update_ms'], 'value G', d['value']/1e9)"
done
