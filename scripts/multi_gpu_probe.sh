#!/bin/bash
# round 2, GPU call 3 (2 GPUs): NCCL paths of the new bench (perm-shard weak + strong, pair-shard C4, replicas C5) and the
# pose-sharded runner
set -u
OUT=gpurun_out; mkdir -p $OUT
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/final_bench_${N}gpu.json 2> $OUT/final_bench_${N}gpu.err; echo "bench $N rc=$?"; tail -3 $OUT/final_bench_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/pose_shard_probe.py > $OUT/final_pose_shard_${N}gpu.log 2>&1; echo "pose shard rc=$?"; tail -6 $OUT/final_pose_shard_${N}gpu.log
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/final_bench_${N}gpu.json").read().strip().splitlines()[-1])
    print("n_gpus", d["n_gpus"], "value %.0f e2e %.0f ms %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
    print("strong", d["strong"])
    for k, v in (d["configs"] or {}).items():
        print("  ", k, {kk: (round(vv, 1) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk not in ("note","workload")})
except Exception as e:
    print("bench unreadable:", e)
PY
