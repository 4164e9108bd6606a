#!/bin/bash
# the bench lines of the shipped build (1 GPU)
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_r02_ur10.json 2> gpurun_out/bench_r02_ur10.err; echo "ur10 rc=$?"
timeout 600 python bench.py --config tiago > gpurun_out/bench_r02_tiago.json 2> gpurun_out/bench_r02_tiago.err; echo "tiago rc=$?"
python - <<'PY'
import json,glob
for f in ('gpurun_out/bench_r02_ur10.json','gpurun_out/bench_r02_tiago.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, round(d['value']/1e6,3),'M/s', 'step',round(d['ms_per_step'],3), 'kernel', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['ms_per_step'], d['parity_check'].get('torque_norm_err'), (d.get('cpu_baseline') or {}).get('value'))
PY
