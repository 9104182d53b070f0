#!/bin/bash
# final single-GPU pass: full GPU suite, smoke, both bench arms
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q -s > $OUT/final_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; grep -E "passed|failed" $OUT/final_gpu_tests.log | tail -2; grep -E "^FAILED|^ERROR" $OUT/final_gpu_tests.log | head -20
grep -E "max\|I\||float64 audit|cloud .*ours-ref|worst logits|vs reference|chain vs|collapsed vs|further than|phi err|N=" $OUT/final_gpu_tests.log > $OUT/final_parity_numbers.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/final_bench_reference.json 2> $OUT/final_bench_reference.err; echo "reference arm rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/final_bench.json 2> $OUT/final_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    r = json.loads(open("gpurun_out/final_bench_reference.json").read().strip().splitlines()[-1])
    print("reference arm: %.1f forwards/s, kind %s, cores %s" % (r["value"], r["cpu_baseline"]["kind"], r["cpu_baseline"]["cores"]))
    d = json.loads(open("gpurun_out/final_bench.json").read().strip().splitlines()[-1])
    print("value %.0f e2e %.0f ms %.2f launches %d clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"]))
    rf = d["roofline"]; print({k: rf[k] for k in ("kernel", "frac", "achieved", "traffic", "algorithmic_per_launch", "executed_frac_of_tf32_peak", "executed_frac_of_tf32_sustained")})
    print("strong", round(d["strong"]["value"]))
    for k, v in d["configs"].items(): print("  ", k, round(v.get("value", 0)), v.get("error", ""))
except Exception as e:
    print("unreadable:", e)
PY
