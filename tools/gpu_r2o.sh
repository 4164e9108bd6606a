#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
for c in ur10 tiago; do
FG_RQ_CALIBRATE=2 timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check 2>gpurun_out/ab.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['ms_per_step'],3))"
grep "\[fg\] rq calibration:" gpurun_out/ab.err | cut -c1-300
done
