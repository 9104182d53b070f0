#!/bin/bash
# round 2, GPU call 2: full GPU suite (minus goldens still being generated) + the new bench line with all legs
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q -s -k "not c4_interactions and not interaction_path and not pointnet2_100 and not whole_chain" > $OUT/r2_2_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; grep -E "passed|failed|error" $OUT/r2_2_gpu_tests.log | tail -5; grep -E "^FAILED|^ERROR" $OUT/r2_2_gpu_tests.log | head -20
grep -E "collapsed vs plain|dgcnn_headline|dgcnn_2048_4" $OUT/r2_2_gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/r2_2_bench.json 2> $OUT/r2_2_bench.err; echo "bench rc=$?"; tail -3 $OUT/r2_2_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_2_bench.json").read().strip().splitlines()[-1])
    print("value %.0f e2e %.0f ms %.2f rows %.3f gate %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["rows_evaluated_fraction"], d["parity_gate"]))
    print("tf32", d["tf32_peak"]); print("cpu", d["cpu_baseline"]); print("strong", d["strong"])
    for k, v in (d["configs"] or {}).items():
        print("  ", k, {kk: (round(vv, 1) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk not in ("note",)})
    for k in d["breakdown"]["kernels"]:
        print("   %-18s %-6s frac %.3f  ach %9.1f %s  launches %3d share %.3f exec_frac %s" % (k["kernel"], k["bound"], k["frac"], k["achieved"], k["unit"], k["launches_per_step"], k["share_of_step"], k.get("executed_frac_of_tf32_peak")))
    print(d["breakdown"]["evaluated_clouds_by_points"])
except Exception as e:
    print("bench unreadable:", e)
PY
