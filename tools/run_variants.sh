#!/bin/bash
# usage: tools/run_variants.sh <args to bench_configs.py...>   runs every variants/libfemb_*.so through the same config
for so in variants/libfemb_*.so; do
  echo "== $so"
  FEMB_LIB=$PWD/$so python tools/bench_configs.py "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print({k: r[k] for k in ('ms_per_step', 'cells_per_s', 'check_max_abs_colsum', 'check_sum_b') if k in r})
"
done
