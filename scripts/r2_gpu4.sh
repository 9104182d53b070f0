#!/bin/bash
# round 2, GPU call 4: chained grouped MLP (PointNet++) -- parity first, then A/B timing, then the whole GPU suite
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_tc_gemm.py -q -s -k "chain" > $OUT/r2_4_chain_test.log 2>&1; echo "chain test rc=$?"; tail -4 $OUT/r2_4_chain_test.log
timeout 600 python -m pytest tests/test_gpu_models.py -q -s -k "pointnet2" > $OUT/r2_4_pn2_tests.log 2>&1; echo "pn2 tests rc=$?"; tail -4 $OUT/r2_4_pn2_tests.log
for MODE in chain nochain; do
  if [ $MODE = nochain ]; then export IQ_TC_NO_CHAIN=1; else unset IQ_TC_NO_CHAIN; fi
  timeout 600 python bench.py --config C2 --perms 100 --scaling weak --steps 3 --warmup 3 --no-extras --no-cpu-baseline > $OUT/r2_4_c2_$MODE.json 2> $OUT/r2_4_c2_$MODE.err; echo "bench C2 $MODE rc=$?"
done
unset IQ_TC_NO_CHAIN
python - <<'PY'
import json
for f in ("chain", "nochain"):
    try:
        d = json.loads(open("gpurun_out/r2_4_c2_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.0f ms %.2f gate %s" % (d["value"], d["ms_per_step"], d["parity_gate"]))
        for k, v in list(d["breakdown"]["by_kernel"].items())[:8]:
            print("   %-22s %8.3f ms %4d launches  %.3f" % (k, v["ms"], v["launches"], v["share"]))
    except Exception as e:
        print(f, "unreadable:", e)
PY
# headline A/B: chunk size and lanes with the collapsed buckets
for CFG in "--chunk 0" "--chunk 222" "--chunk 296" "--chunk 444"; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline $CFG 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$CFG', 'value %.0f ms %.2f' % (d['value'], d['ms_per_step']))"
done
for L in 1 3 4; do
  IQ_LANES=$L timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lanes $L', 'value %.0f ms %.2f' % (d['value'], d['ms_per_step']))"
done
timeout 2400 python -m pytest tests -m gpu -q -s > $OUT/r2_4_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; grep -E "passed|failed" $OUT/r2_4_gpu_tests.log | tail -3; grep -E "^FAILED|^ERROR" $OUT/r2_4_gpu_tests.log | head -20
grep -E "max\|I\||re-evaluated|cloud .*ours-ref|worst logits|logits err vs|chain vs" $OUT/r2_4_gpu_tests.log > $OUT/r2_4_parity_numbers.txt; wc -l $OUT/r2_4_parity_numbers.txt
