#!/bin/bash
# bench.py per model with one and two chunk lanes (IQ_LANES overrides the per-model default)
mkdir -p gpurun_out
: > gpurun_out/lanes_probe.log
for m in ${MODELS_LIST:-pointnet2 pointconv pointnet gcnn dgcnn}; do
  for l in ${LANES_LIST:-1 2}; do
    IQ_LANES=$l timeout 200 python bench.py --model $m --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | \
      python -c "import sys,json; j=json.loads([x for x in sys.stdin if x.startswith('{')][-1]); print('$m lanes=$l value %.0f e2e %.0f ms %.2f' % (j['value'], j['e2e']['value'], j['ms_per_step']))" >> gpurun_out/lanes_probe.log
  done
done
cat gpurun_out/lanes_probe.log
