#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== pytest gpu full"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/probe_models.py talos 60000 2>&1 | tail -1
FG_WIDE_OCC2=1 timeout 300 python tools/probe_models.py talos 60000 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_ur10.json 2> gpurun_out/bench_ur10.err; echo "ur10 rc=$?"
timeout 600 python bench.py --config tiago --steps 5 --no-cpu-baseline > gpurun_out/bench_tiago.json 2> gpurun_out/bench_tiago.err; echo "tiago rc=$?"
for p in 20 23 27 30; do FG_RQ_NPCTA=$p timeout 600 python bench.py --config tiago --steps 5 --no-cpu-baseline --no-e2e --no-parity-check > gpurun_out/bench_tiago_c$p.json 2> gpurun_out/bench_tiago_c$p.err; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,2),'M/s', 'step',round(d['ms_per_step'],3),'kern',round(d['roofline']['kernel_ms'],3),'frac',round(d['roofline']['frac'],3),'e2e', d.get('e2e',{}).get('value'), d.get('parity_check'))
    except Exception as e:
        print(f,'ERR',e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_accumulate_samples -s 1 -c 1 -f -o gpurun_out/prof_wide3 \
    python tools/probe_models.py talos 30000 > gpurun_out/prof_wide3.log 2>&1; echo "ncu wide rc=$?"
