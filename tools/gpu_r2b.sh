#!/bin/bash
# Round-2 call B: SM-specialised ring kernel A/B, wide kernel bring-up (Talos / TIAGo-24), profiles.
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== pytest gpu"
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log
echo "== bench A/B (producer CTAs, ring)"
for npcta in 0 18 22 26 30 36; do for ring in 1024; do
  if [ $npcta = 0 ]; then unset FG_RQ_NPCTA; else export FG_RQ_NPCTA=$npcta; fi
  FG_RQ_RING=$ring timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rq_c${npcta}_r${ring}.json 2> gpurun_out/bench_rq_c${npcta}_r${ring}.err; echo "rq npcta=$npcta ring=$ring rc=$?"
done; done
unset FG_RQ_NPCTA
FG_RQ_RING=512 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rq_c0_r512.json 2> gpurun_out/bench_rq_c0_r512.err
FG_RQ_RING=2048 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rq_c0_r2048.json 2> gpurun_out/bench_rq_c0_r2048.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_rq_c*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,1),'M/s', 'step',round(d['ms_per_step'],3),'kern',round(d['roofline']['kernel_ms'],3),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']/1e6,1), d['config'].get('eliminated'), d['config'].get('base_parameters'))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-400:])
PY
echo "== model probes"
timeout 300 python tools/probe_models.py tiago 400000 2>&1 | tail -1
timeout 300 python tools/probe_models.py talos 100000 2>&1 | tail -1
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full (1M launch)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rq_accumulate -s 3 -c 1 -f -o gpurun_out/prof_rq \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_rq_bench.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_accumulate_samples -s 2 -c 1 -f -o gpurun_out/prof_wide \
    python tools/probe_models.py talos 30000 > gpurun_out/prof_wide.log 2>&1; echo "ncu wide rc=$?"
ls gpurun_out | head -60
