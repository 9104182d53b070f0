#!/bin/bash
# ncu --set full of one kernel regex with chosen skip/count.  Usage: gpu_profile_one.sh <tag> <regex> <skip> <count> [bench args]
set -u
TAG=$1; RX=$2; SKIP=$3; CNT=$4; shift 4
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --perms 10 --no-cpu-baseline $*"
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c $CNT -f -o $OUT/${TAG} $CMD > $OUT/${TAG}_ncu.log 2>&1
ls -la $OUT | grep ${TAG}
