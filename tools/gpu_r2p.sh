#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 900 python bench.py --config talos --steps 3 --warmup 3 --no-cpu-baseline ${EXTRA} 2>gpurun_out/ab.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('talos', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],3), d.get('parity_check'))"
tail -3 gpurun_out/ab.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_talos.csv python bench.py --config talos --samples 200000 --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check --no-e2e > /dev/null 2>&1; echo "rc=$?"
