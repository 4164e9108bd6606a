#!/bin/bash
for n in 100000 200000 500000 1000000; do
timeout 900 python bench.py --config talos --samples $n --steps 3 --warmup 2 --no-cpu-baseline --no-parity-check --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('talos $n', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4))"
done
