#!/bin/bash
# round 2, GPU call 1: collapse parity + the full GPU suite + bench A/B (collapsed vs IQ_NO_COLLAPSE)
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_collapse.py -x -q -s > $OUT/r2_1_collapse_tests.log 2>&1; echo "collapse tests rc=$?"; tail -15 $OUT/r2_1_collapse_tests.log
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_collapse.py > $OUT/r2_1_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; tail -8 $OUT/r2_1_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r2_1_bench_collapse.json 2> $OUT/r2_1_bench_collapse.err; echo "bench rc=$?"
IQ_NO_COLLAPSE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/r2_1_bench_plain.json 2> $OUT/r2_1_bench_plain.err; echo "bench plain rc=$?"
python - <<'PY'
import json
for f in ("r2_1_bench_collapse", "r2_1_bench_plain"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.0f e2e %.0f ms %.2f gate %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["parity_gate"]))
        for k, v in list(d["breakdown"]["by_kernel"].items())[:14]:
            print("   %-22s %8.3f ms %4d launches  %.3f" % (k, v["ms"], v["launches"], v["share"]))
    except Exception as e:
        print(f, "unreadable:", e)
PY
