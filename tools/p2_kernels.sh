#!/bin/bash
# usage: tools/p2_kernels.sh <tag> [env...]: per-kernel durations (ncu, gpu__time_duration) of the C3 path at m = 96
TAG=$1; shift
env "$@" ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_p2|k_cell|k_h1_gather" -s 6 -c 6 --csv --log-file gpurun_out/l_$TAG.csv python tools/bench_configs.py --which c3 --m3 96 --steps 1 --warmup 1 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/l_$TAG.csv")) if len(r)>10 and r[0].isdigit()]
out={}
for r in rows: out.setdefault((r[0], r[4].split('(')[0][-28:]), {})[r[-3].split('.')[0].replace('__','_')[-22:]]=r[-1]
for k,v in out.items(): print("$TAG", k, v)
PY
