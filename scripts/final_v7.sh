#!/bin/bash
# last single-GPU pass of round 2: threshold of the small-cloud gather form (same box A/B), then the GPU suite, smoke and
# the driver-style bench line at the best threshold, and a narrow ncu --set full capture of the kernels that changed format.
set -u
OUT=gpurun_out; mkdir -p $OUT
for t in 512 768 1024; do
    IQ_GM_SMALL=$t timeout 200 python bench.py --no-extras --no-cpu-baseline --steps 10 --warmup 3 > $OUT/v7_bench_gm$t.json 2> $OUT/v7_bench_gm$t.err; echo "bench IQ_GM_SMALL=$t rc=$?"
done
BEST=$(python - <<'PY'
import json
vals = {}
for t in (512, 768, 1024):
    try:
        vals[t] = json.loads(open("gpurun_out/v7_bench_gm%d.json" % t).read().strip().splitlines()[-1])["value"]
    except Exception:
        pass
best = 512
for t, v in vals.items():                     # leave 512 (validated earlier) only for a gain beyond the run-to-run noise
    if v > vals.get(best, 0.0) * (1.005 if best == 512 else 1.0):
        best = t
print(best)
PY
)
python - <<'PY'
import json
for t in (512, 768, 1024):
    try:
        d = json.loads(open("gpurun_out/v7_bench_gm%d.json" % t).read().strip().splitlines()[-1])
        print("IQ_GM_SMALL=%d value %.0f ms %.2f gather_max %.2f" % (t, d["value"], d["ms_per_step"], d["breakdown"]["by_kernel"]["gather_max"]["ms"]))
    except Exception as e:
        print(t, "unreadable", e)
PY
echo "best threshold: $BEST"; export IQ_GM_SMALL=$BEST
timeout 900 python -m pytest tests -m gpu -q -s > $OUT/v7_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; grep -E "passed|failed" $OUT/v7_gpu_tests.log | tail -2; grep -E "^FAILED|^ERROR" $OUT/v7_gpu_tests.log | head -20
grep -E "max\|I\||float64 audit|cloud .*ours-ref|worst logits|vs reference|chain vs|collapsed vs|further than|phi err|N=|f16 paths|f16x2" $OUT/v7_gpu_tests.log > $OUT/v7_parity_numbers.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/v7_bench.json 2> $OUT/v7_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/v7_bench.json").read().strip().splitlines()[-1])
    ks = {k: round(v["ms"], 2) for k, v in d["breakdown"]["by_kernel"].items() if v["ms"] > 0.3}
    print("value %.0f e2e %.0f ms %.2f launches %d clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"]))
    print("    ", ks)
    print("     strong", round(d["strong"]["value"]), {k: round(v.get("value", 0)) for k, v in d["configs"].items()})
except Exception as e:
    print("unreadable:", e)
PY
timeout 200 ncu --set full --clock-control none --profile-from-start off -k regex:'gemm_tc_kernel|gram_knn_kernel|gather_max_smem' -f -o /tmp/v7_f16 python scripts/profile_all_kernels.py dgcnn > $OUT/v7_ncu_f16.log 2>&1; echo "ncu capture rc=$?"
ncu -i /tmp/v7_f16.ncu-rep --page raw --csv > $OUT/v7_f16_raw.csv 2>/dev/null; ls -la /tmp/v7_f16.ncu-rep $OUT/v7_f16_raw.csv
