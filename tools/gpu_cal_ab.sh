#!/bin/bash
# calibration robustness: the same bench five times (what does the timed calibration pick, what does the kernel take)
for rep in 1 2 3 4 5; do
for c in ur10 tiago; do
  FG_RQ_CALIBRATE=2 timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity-check 2>gpurun_out/ab.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4))"
  grep "rq calibration:" gpurun_out/ab.err | tail -1 | cut -c1-200
done; done
