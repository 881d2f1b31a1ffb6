#!/bin/bash
# usage: tools/ncu_full.sh <kernel-regex> <out-name> <skip> <count> -- <command...>
# Runs the command plainly first (must exit 0), then once under `ncu --set full` for the named kernel.
set -e
KREGEX=$1; OUT=$2; SKIP=$3; COUNT=$4; shift 5
mkdir -p gpurun_out
"$@" > gpurun_out/plain_$OUT.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $SKIP -c $COUNT -f -o gpurun_out/$OUT "$@" > gpurun_out/ncu_$OUT.log 2>&1
tail -3 gpurun_out/ncu_$OUT.log
