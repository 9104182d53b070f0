#!/bin/bash
# round 2 ncu evidence (one GPU, under gpurun):  scripts/r2_profile.sh <tag>
#  1. launch list (gpu__time_duration) + DRAM bytes of every launch of ONE headline step  -> <tag>_step.csv
#  2. --set full of the hot path: coalition kernels, the collapse kernels, one DGCNN chunk  -> <tag>_hot_raw.csv
#  3. the other four models, the sections the summary table needs                          -> <tag>_models_raw.csv
# Reports are converted to CSV on the box and deleted: gpurun only brings back 64 MiB.
set -u
TAG=$1
OUT=gpurun_out; mkdir -p $OUT
python scripts/profile_step.py > $OUT/${TAG}_step_plain.log 2>&1 || { echo "plain step failed"; tail -5 $OUT/${TAG}_step_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_step.csv python scripts/profile_step.py > $OUT/${TAG}_ncu_step.log 2>&1; echo "step list rc=$?"
python scripts/profile_all_kernels.py > $OUT/${TAG}_all_plain.log 2>&1 || { echo "profile_all plain run failed"; tail -5 $OUT/${TAG}_all_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/${TAG}_hot python scripts/profile_all_kernels.py coalition dgcnn > $OUT/${TAG}_ncu_hot.log 2>&1; echo "hot capture rc=$?"
ncu -i /tmp/${TAG}_hot.ncu-rep --page raw --csv > $OUT/${TAG}_hot_raw.csv 2>/dev/null; ls -la /tmp/${TAG}_hot.ncu-rep
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis --clock-control none --profile-from-start off -f -o /tmp/${TAG}_models python scripts/profile_all_kernels.py gcnn pointnet pointnet2 pointconv > $OUT/${TAG}_ncu_models.log 2>&1; echo "models capture rc=$?"
ncu -i /tmp/${TAG}_models.ncu-rep --page raw --csv > $OUT/${TAG}_models_raw.csv 2>/dev/null; ls -la /tmp/${TAG}_models.ncu-rep
du -sh $OUT; ls -la $OUT | grep ${TAG}
