#!/bin/bash
# A/B of the kind::f16 operand paths (IQ_F16_CONV5 / IQ_F16_STORE / IQ_F16_GRAM) on one GPU, under gpurun:
#   unit tests of the fp16 GEMM forms, the kNN unit tests with the fp16 Gram, the parity suites with every path on,
#   and the headline bench with the paths switched on one by one (same box, same build).
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_tc_gemm.py -q -s -k "f16" > $OUT/f16_unit.log 2>&1; echo "f16 unit rc=$?"
grep -E "passed|failed|f16 paths|^FAILED|^ERROR|rror:" $OUT/f16_unit.log | tail -24
IQ_F16_CONV5=1 IQ_F16_GRAM=1 timeout 300 python -m pytest tests/test_gpu_knn_tc.py -q -s > $OUT/f16_knn.log 2>&1; echo "knn unit (fp16 Gram) rc=$?"; tail -2 $OUT/f16_knn.log
for cfg in "0 0 0" "1 0 0" "1 1 0" "1 1 1"; do
    set -- $cfg
    IQ_F16_CONV5=$1 IQ_F16_STORE=$2 IQ_F16_GRAM=$3 timeout 240 python bench.py --no-extras --no-cpu-baseline --steps 10 --warmup 3 \
        > $OUT/f16_bench_$1$2$3.json 2> $OUT/f16_bench_$1$2$3.err; echo "bench $1$2$3 rc=$?"
done
python - <<'PY'
import json
for tag in ("000", "100", "110", "111"):
    try:
        d = json.loads(open("gpurun_out/f16_bench_%s.json" % tag).read().strip().splitlines()[-1])
        ks = {k: round(v["ms"], 2) for k, v in d["breakdown"]["by_kernel"].items() if v["ms"] > 0.3}
        print(tag, "value %.0f e2e %.0f ms %.2f gate %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["parity_gate"]))
        print("    ", ks)
        rf = d["roofline"]; print("    ", {k: rf.get(k) for k in ("kernel", "frac", "mma_kind", "executed_frac_of_bf16_sustained", "executed_frac_of_tf32_peak")})
    except Exception as e:
        print(tag, "unreadable:", e)
PY
IQ_F16_CONV5=1 IQ_F16_STORE=1 IQ_F16_GRAM=1 timeout 400 python -m pytest tests/test_gpu_wide_parity.py tests/test_gpu_models.py tests/test_gpu_collapse.py tests/test_gpu_edge_cases.py -q -s \
    > $OUT/f16_parity.log 2>&1; echo "parity suites (all fp16 paths) rc=$?"
grep -E "passed|failed|^FAILED|^ERROR" $OUT/f16_parity.log | tail -12
grep -E "float64 audit|further than|worst logits|collapsed vs|phi err" $OUT/f16_parity.log | head -30
