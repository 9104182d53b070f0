#!/bin/bash
# source-level ncu capture of two kernels of the DGCNN chunk (gather_max, knn_rerank), one launch each
set -u
OUT=gpurun_out; mkdir -p $OUT
python scripts/profile_all_kernels.py dgcnn > /dev/null 2>&1
for RX in gather_max_smem_kernel knn_rerank_mask_kernel; do
  timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:$RX -s 2 -c 1 -f -o $OUT/src_$RX python scripts/profile_all_kernels.py dgcnn > $OUT/src_$RX.log 2>&1; echo "$RX rc=$?"
done
ls -la $OUT/*.ncu-rep; du -sh $OUT
