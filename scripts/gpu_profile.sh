#!/bin/bash
# Profiles the hot path on one B200 (run under gpurun).  Usage: scripts/gpu_profile.sh <tag> <kernel-regex>
# 1. plain run must exit 0;  2. ncu launch list (gpu__time_duration) of the same command;
# 3. plain run again;        4. ncu --set full of the top kernel (-k regex, 3 launches).
set -u
TAG=${1:-r1}
KREGEX=${2:-sgemm_kernel}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --perms 10 --no-cpu-baseline"
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
$CMD > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 20 -c 3 -f -o $OUT/${TAG}_top $CMD > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT
