#!/bin/bash
# round 2 ncu evidence (one GPU, under gpurun):  scripts/r2_profile.sh <tag>
#  1. launch list (gpu__time_duration) of one default bench step       -> <tag>_launches.csv
#  2. dram bytes of every launch of that step (traffic per kernel)      -> <tag>_traffic.csv
#  3. --set full of every kernel family once (all five models + coalition kernels) -> <tag>_all.ncu-rep
set -u
TAG=$1
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline"
$CMD > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/${TAG}_traffic.csv $CMD > $OUT/${TAG}_ncu_traffic.log 2>&1; echo "traffic rc=$?"
python scripts/profile_all_kernels.py > $OUT/${TAG}_all_plain.log 2>&1 || { echo "profile_all plain run failed"; tail -5 $OUT/${TAG}_all_plain.log; exit 1; }
timeout 2400 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $OUT/${TAG}_all python scripts/profile_all_kernels.py > $OUT/${TAG}_ncu_all.log 2>&1; echo "full capture rc=$?"
ls -la $OUT | grep ${TAG}
