#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12
for v in 1 0; do
FG_USE_CQ=$v timeout 900 python bench.py --config talos --steps 3 --warmup 3 --no-cpu-baseline ${EXTRA} 2>gpurun_out/ab.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('talos cq=$v', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],3), d.get('parity_check'))"
tail -3 gpurun_out/ab.err
done
