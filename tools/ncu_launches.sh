#!/bin/bash
# usage: tools/ncu_launches.sh <out-name> <skip> <count> -- <command...>   (per-launch device times, cold-cache, serialised)
set -e
OUT=$1; SKIP=$2; COUNT=$3; shift 4
mkdir -p gpurun_out
"$@" > gpurun_out/plain_$OUT.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $COUNT --csv --log-file gpurun_out/$OUT.csv "$@" > gpurun_out/ncu_$OUT.log 2>&1
tail -2 gpurun_out/ncu_$OUT.log
