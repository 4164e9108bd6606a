#!/bin/bash
# time the fused kernel under a few launch configurations (run under gpurun)
mkdir -p gpurun_out
for cfg in ${CFGS:-"8 4" "8 8" "4 4" "4 2" "3 1" "2 1" "1 1"}; do
  set -- ${cfg/:/ }
  echo "== warps=$1 wpr=$2" 
  FG_BQ_WARPS=$1 FG_BQ_WPR=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --samples ${SAMPLES:-1000000} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('step_ms', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'])
    elif 'rror' in l: print(l.strip())
"
done
