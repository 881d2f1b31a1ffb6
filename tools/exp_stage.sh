#!/bin/bash
for st in 1 0; do
  for cv in given device; do
  echo -n "STAGE=$st cellvol=$cv: "
  FEMB_P1_STAGE=$st python bench.py --steps 10 --warmup 3 --no-cpu --cellvol $cv 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value']/1e9, d['roofline']['kernel_ms'], d['roofline']['frac'])"
done; done
