#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
FAIL=0
for ID in "32-32-64-16" "64-64-128-32" "64-96-128-128" "128-128-256-64" "128-128-256-128" "32-32-64-128" "128-128-256-16"; do
  timeout 120 python -m pytest tests/test_gpu_tc_gemm.py -q -s -k "chained_grouped_mlp_kernel and $ID" > $OUT/r2_11_chain_$ID.log 2>&1; RC=$?; [ $RC -ne 0 ] && FAIL=1
  echo "chain $ID rc=$RC $(grep -E 'passed|failed|assert' $OUT/r2_11_chain_$ID.log | tail -2 | tr '\n' ' ')"
done
[ $FAIL -ne 0 ] && { echo "chain unit tests failed: stopping"; exit 0; }
timeout 300 python -m pytest tests/test_gpu_tc_gemm.py -q -s -k "pointnet2_chain" > $OUT/r2_11_chain_model.log 2>&1; echo "model chain rc=$?"; grep -E "chain vs|passed|failed" $OUT/r2_11_chain_model.log | tail -3
for MODE in chain nochain; do
  if [ $MODE = nochain ]; then export IQ_TC_NO_CHAIN=1; else unset IQ_TC_NO_CHAIN; fi
  timeout 600 python bench.py --config C2 --perms 100 --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $OUT/r2_11_c2_$MODE.json 2> $OUT/r2_11_c2_$MODE.err; echo "bench C2 $MODE rc=$?"
done
unset IQ_TC_NO_CHAIN
python - <<'PY'
import json
for f in ("chain", "nochain"):
    try:
        d = json.loads(open("gpurun_out/r2_11_c2_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.0f ms %.2f gate %s" % (d["value"], d["ms_per_step"], d["parity_gate"]))
        for k, v in list(d["breakdown"]["by_kernel"].items())[:8]:
            print("   %-22s %8.3f ms %4d launches  %.3f" % (k, v["ms"], v["launches"], v["share"]))
        r = d["roofline"]; print("   roofline:", r["kernel"], "frac %.3f ach %.1f exec_frac %s" % (r["frac"], r["achieved"], r.get("executed_frac_of_tf32_peak")))
    except Exception as e:
        print(f, "unreadable:", e)
PY
timeout 2400 python -m pytest tests -m gpu -q -s -k "(c4_interactions or pointnet2) and not chained_grouped_mlp_kernel" > $OUT/r2_11_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; grep -E "passed|failed" $OUT/r2_11_gpu_tests.log | tail -3; grep -E "^FAILED|^ERROR" $OUT/r2_11_gpu_tests.log | head -20
grep -E "max\|I\||re-evaluated|cloud .*ours-ref|worst logits|logits err vs|chain vs|collapsed vs" $OUT/r2_11_gpu_tests.log > $OUT/r2_11_parity_numbers.txt; wc -l $OUT/r2_11_parity_numbers.txt
