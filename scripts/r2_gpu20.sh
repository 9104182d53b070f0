#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python bench.py --perms 1000 --scaling strong --steps 5 --warmup 3 --no-extras --no-cpu-baseline > $OUT/r2_20_strong1000_1gpu.json 2> $OUT/r2_20_strong1000_1gpu.err; echo "strong-1000 1 GPU rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r2_20_strong1000_1gpu.json').read().strip().splitlines()[-1]); print('1000 perms: value %.0f e2e %.0f ms %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
bash scripts/r2_profile.sh r2p3
