set -u
OUT=gpurun_out; mkdir -p $OUT
IQ_TC_ALL=1 timeout 600 python -m pytest tests/test_gpu_wide_parity.py -q -s -k "dgcnn_headline" > $OUT/tcall_parity.log 2>&1; echo "rc=$?"; grep -E "dgcnn_headline|cloud " $OUT/tcall_parity.log | cut -c1-200 | tail -30
IQ_TC_ALL=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('IQ_TC_ALL value %.0f ms %.2f gate %s' % (d['value'], d['ms_per_step'], d['parity_gate']))"
