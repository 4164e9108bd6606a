#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for c in ur10 tiago; do
for f in figaroh-plus_b200/lib/libfigaroh_b200.so figaroh-plus_b200/lib/variants/lib_rq_*.so; do
  FG_B200_LIB=$PWD/$f FG_RQ_CALIBRATE=0 timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity-check 2>gpurun_out/ab.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$f $c', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4))"
done; done; done
