#!/bin/bash
# final single-GPU pass of round 2 after the kind::f16 paths: GPU suite, smoke, the driver-style bench line, two tuning
# A/Bs (chunk lanes, small-cloud gather form) and a narrow ncu --set full capture of the kernels that changed format.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -s > $OUT/v6_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; grep -E "passed|failed" $OUT/v6_gpu_tests.log | tail -2; grep -E "^FAILED|^ERROR" $OUT/v6_gpu_tests.log | head -20
grep -E "max\|I\||float64 audit|cloud .*ours-ref|worst logits|vs reference|chain vs|collapsed vs|further than|phi err|N=|f16 paths|f16x2" $OUT/v6_gpu_tests.log > $OUT/v6_parity_numbers.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/v6_bench.json 2> $OUT/v6_bench.err; echo "bench rc=$?"
IQ_RERANK_F2=1 timeout 200 python -m pytest tests/test_gpu_knn_tc.py -q > $OUT/v6_knn_f2.log 2>&1; echo "knn unit (packed re-rank) rc=$?"; tail -1 $OUT/v6_knn_f2.log
IQ_SGEMM_F2=1 timeout 200 python -m pytest tests/test_gpu_tc_gemm.py -q -k "store_epilogue or pool_epilogue" > $OUT/v6_sgemm_f2.log 2>&1; echo "sgemm unit (packed fma) rc=$?"; tail -1 $OUT/v6_sgemm_f2.log
for cfg in "IQ_LANES=3" "IQ_GM_SMALL=512" "IQ_SGEMM_F2=1" "IQ_RERANK_F2=1"; do
    env $cfg timeout 200 python bench.py --no-extras --no-cpu-baseline --steps 10 --warmup 3 > $OUT/v6_bench_$cfg.json 2> $OUT/v6_bench_$cfg.err; echo "bench $cfg rc=$?"
done
IQ_SGEMM_F2=1 IQ_RERANK_F2=1 timeout 200 python bench.py --no-extras --no-cpu-baseline --steps 10 --warmup 3 > $OUT/v6_bench_IQ_F2_BOTH.json 2> $OUT/v6_bench_IQ_F2_BOTH.err; echo "bench both F2 rc=$?"
python - <<'PY'
import json, glob
def show(fn):
    try:
        d = json.loads(open(fn).read().strip().splitlines()[-1])
        ks = {k: round(v["ms"], 2) for k, v in d["breakdown"]["by_kernel"].items() if v["ms"] > 0.3}
        print(fn, "value %.0f e2e %.0f ms %.2f launches %d clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"]))
        print("    ", ks)
        rf = d["roofline"]; print("    ", {k: rf.get(k) for k in ("kernel", "frac", "achieved", "traffic", "mma_kind")})
        for k in d["breakdown"]["kernels"]:
            if k["bound"] == "tensor": print("      ", k["kernel"], "frac %.3f" % k["frac"], k.get("mma_kind"), "executed/bf16 %s" % k.get("executed_frac_of_bf16_sustained"))
        if d.get("strong"): print("     strong", round(d["strong"]["value"]))
        for k, v in (d.get("configs") or {}).items(): print("       ", k, round(v.get("value", 0)), v.get("error", ""))
    except Exception as e:
        print(fn, "unreadable:", e)
show("gpurun_out/v6_bench.json")
for fn in sorted(glob.glob("gpurun_out/v6_bench_IQ*.json")): show(fn)
PY
