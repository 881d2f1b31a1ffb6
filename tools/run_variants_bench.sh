#!/bin/bash
# usage: tools/run_variants_bench.sh <args to bench.py...>   every variants/libfemb_*.so through bench.py (device-resident value + kernel time only)
for so in variants/libfemb_*.so; do
  echo "== $so"
  FEMB_LIB=$PWD/$so python bench.py --no-cpu --no-legs "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print({'ms': round(r['ms_per_step'],4), 'kernel_ms': r['roofline']['kernel_ms'], 'frac': r['roofline']['frac'], 'dev_vol': r['value_cellvol_device'] and r['value_cellvol_device']['kernel_ms'], 'checks': r['checks']})
"
done
