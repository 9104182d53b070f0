#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q -s > $OUT/r2_10_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; grep -E "passed|failed" $OUT/r2_10_gpu_tests.log | tail -3; grep -E "^FAILED|^ERROR" $OUT/r2_10_gpu_tests.log | head -20
grep -E "max\|I\||float64 audit|cloud .*ours-ref|worst logits|vs reference|chain vs|collapsed vs|further than|phi err" $OUT/r2_10_gpu_tests.log > $OUT/r2_10_parity_numbers.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/r2_10_bench.json 2> $OUT/r2_10_bench.err; echo "bench rc=$?"
bash scripts/r2_profile.sh r2p2
