#!/bin/bash
# Round profile: (1) launch list with per-launch device time, (2) one --set full capture of one forward chunk
# (every kernel family once).  Usage: gpu_profile_round.sh <tag>     (run under gpurun; outputs in gpurun_out/)
set -u
TAG=$1
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --perms 20 --no-cpu-baseline"
$CMD > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 330 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
$CMD > $OUT/${TAG}_plain2.log 2>&1 || { echo "plain run 2 failed"; exit 1; }
# launches 330.. are the second warm-up step: mask, then knn_xyz .. conv5 of the first 148-cloud chunk
timeout 600 ncu --set full --clock-control none -s 330 -c 24 -f -o $OUT/${TAG}_full $CMD > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT | grep ${TAG}
