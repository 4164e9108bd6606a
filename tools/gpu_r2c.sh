#!/bin/bash
# Round-2 call C: wide kernel bring-up (Talos / TIAGo-24 parity), ring kernel variants.
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== pytest gpu"
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log
echo "== bench A/B"
for nw in 16 20; do for npcta in 20 22 24 26 28; do
  FG_RQ_NW=$nw FG_RQ_NPCTA=$npcta timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_v_w${nw}_c${npcta}.json 2> gpurun_out/bench_v_w${nw}_c${npcta}.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_v_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,1),'M/s', 'step',round(d['ms_per_step'],3),'kern',round(d['roofline']['kernel_ms'],3),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']/1e6,1), d['config'].get('eliminated'), d['config'].get('base_parameters'))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-400:])
PY
echo "== model probes"
timeout 300 python tools/probe_models.py tiago 400000 2>&1 | tail -1
timeout 300 python tools/probe_models.py talos 100000 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_accumulate_samples -s 1 -c 1 -f -o gpurun_out/prof_wide \
    python tools/probe_models.py talos 30000 > gpurun_out/prof_wide.log 2>&1; echo "ncu wide rc=$?"
ls gpurun_out | head -60
