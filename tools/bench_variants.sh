#!/bin/bash
# tools/bench_variants.sh NAME...: kernel time of the bench workload for each build/exp/NAME.so (tuning aid)
cd "$(dirname "$0")/.."
for n in "$@"; do
  CLB_LIBRARY_PATH=$PWD/build/exp/$n.so python bench.py --steps 1500 --no-cpu-baseline --e2e-steps 3 ${BENCH_ARGS} 2>&1 | python -c "
import json,sys
t=sys.stdin.read()
try:
    d=json.loads(t.strip().splitlines()[-1]); print('$n', d['config']['kernel'][:50], '%.2f us' % (1e3*d['ms_per_step']))
except Exception as e: print('$n', 'FAILED', t[-400:])"
done
