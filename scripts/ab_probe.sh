#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_knn_tc.py tests/test_gpu_models.py tests/test_gpu_collapse.py -q > $OUT/ab_tests.log 2>&1; echo "tests rc=$?"; tail -2 $OUT/ab_tests.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value %.0f ms %.2f gate %s' % (d['value'], d['ms_per_step'], d['parity_gate']['logits_rel_err'])); 
for k,v in list(d['breakdown']['by_kernel'].items())[:8]: print('   %-20s %7.3f' % (k, v['ms']))"
done
