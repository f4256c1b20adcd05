#!/bin/bash
# sweeps the K1 skeleton switches (B200I_WS_OPTS) over a list of variants
for o in ${OPTS:-0 2 6}; do
  echo "== B200I_WS_OPTS=$o"
  B200I_WS_OPTS=$o VARIANTS=${VARIANTS:-10,20,13,22} python scripts/sweep_k1.py 1000000 5 2>&1 | grep variant | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('  v%d %.3f ms %.0f GB/s' % (d['variant'], d['ms'], d['gbs']))"
done
