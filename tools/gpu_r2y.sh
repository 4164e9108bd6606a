#!/bin/bash
for rep in 1 2; do
for cfgs in "1 17" "1 18" "4 17" "4 18" "2 17" "2 18" "1 19"; do
set -- $cfgs
FG_H2D_STREAMS=$1 FG_H2D_CAP=$2 FG_RQ_CALIBRATE=2 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check 2>gpurun_out/ab.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ur10 streams $1 cap 2^$2', 'step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), round(d['e2e']['value']/1e6,1), 'h2d', round(d['e2e']['h2d_gbs_per_rank'],1))"
done; done
grep "rq calibration (" gpurun_out/ab.err | cut -c1-120
FG_RQ_CALIBRATE=2 timeout 900 python bench.py --config tiago --steps 5 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e 2>&1 >/dev/null | grep "rq calibration (" | cut -c1-120
