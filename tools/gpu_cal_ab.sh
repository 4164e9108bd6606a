#!/bin/bash
# calibration modes A/B: 1 = timed producer count, model apportionment of the row groups; 3 = measured fold cycles too
for rep in 1 2; do
for c in ur10 tiago; do
for m in 2 3; do
  FG_RQ_CALIBRATE=$m timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity-check 2>gpurun_out/ab.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c calibrate=$m', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4))"
  grep "rq calibration:" gpurun_out/ab.err | tail -1 | cut -c1-200
done; done; done
