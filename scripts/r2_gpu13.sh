#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tc_gemm.py -q -k "store_epilogue" > $OUT/r2_21_store.log 2>&1; echo "store tests rc=$?"; tail -3 $OUT/r2_21_store.log
timeout 900 python -m pytest tests -m gpu -q -k "pointconv" > $OUT/r2_21_models.log 2>&1; echo "pointconv/pointnet2 tests rc=$?"; tail -4 $OUT/r2_21_models.log
timeout 300 python bench.py --config C5 --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $OUT/r2_21_C5.json 2> $OUT/r2_21_C5.err; echo "bench C5 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_21_C5.json").read().strip().splitlines()[-1])
    print("C5", "value %.0f ms %.2f" % (d["value"], d["ms_per_step"]))
    for k, v in list(d["breakdown"]["by_kernel"].items())[:14]:
        print("   %-22s %8.3f ms %4d launches  %.3f" % (k, v["ms"], v["launches"], v["share"]))
except Exception as e:
    print("unreadable:", e)
PY
