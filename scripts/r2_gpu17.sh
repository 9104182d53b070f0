#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_edge_cases.py -q -s > $OUT/r2_17_edge.log 2>&1; echo "edge tests rc=$?"; grep -E "passed|failed|N=" $OUT/r2_17_edge.log | tail -8; grep -E "^FAILED" $OUT/r2_17_edge.log | head
