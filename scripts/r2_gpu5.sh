#!/bin/bash
# chain kernel unit tests, one process per configuration (a device fault is sticky per process)
set -u
OUT=gpurun_out; mkdir -p $OUT
for ID in "32-32-64-16" "64-64-128-32" "64-96-128-128" "128-128-256-64" "128-128-256-128" "32-32-64-128" "128-128-256-16"; do
  timeout 120 python -m pytest tests/test_gpu_tc_gemm.py -q -s -k "chained_grouped_mlp_kernel and $ID" > $OUT/r2_5_chain_$ID.log 2>&1; echo "chain $ID rc=$? $(grep -E 'passed|failed|error|assert' $OUT/r2_5_chain_$ID.log | tail -2 | tr '\n' ' ')"
done
timeout 300 python -m pytest tests/test_gpu_tc_gemm.py -q -s -k "pointnet2_chain" > $OUT/r2_5_chain_model.log 2>&1; echo "model chain rc=$?"; grep -E "chain vs|passed|failed" $OUT/r2_5_chain_model.log | tail -3
