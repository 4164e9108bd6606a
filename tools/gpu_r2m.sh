#!/bin/bash
# ring kernel: calibration / back-off A/B
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity-check $EXTRA 2>gpurun_out/ab_$name.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4))"
  grep "\[fg\]" gpurun_out/ab_$name.err | cut -c1-600
}
for rep in 1 2; do
run cal FG_RQ_CALIBRATE=2
run nocal FG_RQ_CALIBRATE=0
run nocal_pc22 FG_RQ_CALIBRATE=0 FG_RQ_NPCTA=22
run nocal_pc18 FG_RQ_CALIBRATE=0 FG_RQ_NPCTA=18
done
EXTRA="--config tiago"
run tiago_cal FG_RQ_CALIBRATE=2
run tiago_nocal FG_RQ_CALIBRATE=0
run tiago_nocal_s20 FG_RQ_CALIBRATE=0 FG_RQ_SLEEP=20
EXTRA=""
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
