#!/bin/bash
# ncu --set full of selected kernels of the DGCNN step (run under gpurun, one GPU).
# Usage: scripts/gpu_profile_kernels.sh <tag> <regex1> [<regex2> ...]
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --perms 10 --no-cpu-baseline"
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
for RX in "$@"; do
  NAME=$(echo $RX | tr -c 'a-zA-Z0-9_' '_')
  $CMD > $OUT/${TAG}_plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$RX -s 30 -c 2 -f -o $OUT/${TAG}_${NAME} $CMD > $OUT/${TAG}_ncu_${NAME}.log 2>&1
done
ls -la $OUT | grep $TAG
