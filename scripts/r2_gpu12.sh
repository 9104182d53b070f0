#!/bin/bash
# breakdowns of the configs that are not the headline (one GPU)
set -u
OUT=gpurun_out; mkdir -p $OUT
for C in C5 C1 C4g; do
  timeout 300 python bench.py --config $C --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $OUT/r2_12_$C.json 2> $OUT/r2_12_$C.err; echo "bench $C rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_12_$C.json").read().strip().splitlines()[-1])
    print("$C", "value %.0f ms %.2f" % (d["value"], d["ms_per_step"]))
    for k, v in list(d["breakdown"]["by_kernel"].items())[:16]:
        print("   %-22s %8.3f ms %4d launches  %.3f" % (k, v["ms"], v["launches"], v["share"]))
except Exception as e:
    print("$C unreadable:", e)
PY
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
