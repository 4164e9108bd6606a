#!/bin/bash
# Round-2 call D: ring kernel variants (again, with the fixed bench), wide kernel v2, bench configs.
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== pytest gpu (quick subset: wide + ring parity)"
timeout 1200 python -m pytest tests/test_gpu_parity_oracle.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
echo "== UR10 bench A/B"
for nw in 16 20; do for npcta in 22 24 26 28; do
  FG_RQ_NW=$nw FG_RQ_NPCTA=$npcta timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e > gpurun_out/bench_v_w${nw}_c${npcta}.json 2> gpurun_out/bench_v_w${nw}_c${npcta}.err
done; done
FG_USE_RQ=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e > gpurun_out/bench_v_gq.json 2> gpurun_out/bench_v_gq.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_v_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,1),'M/s', 'step',round(d['ms_per_step'],3),'kern',round(d['roofline']['kernel_ms'],3),'frac',round(d['roofline']['frac'],3), d['config'].get('eliminated'), d['config'].get('base_parameters'))
    except Exception as e:
        print(f,'ERR',e, open(f.replace('.json','.err')).read()[-600:])
PY
echo "== full bench lines"
timeout 600 python bench.py > gpurun_out/bench_ur10.json 2> gpurun_out/bench_ur10.err; echo "ur10 rc=$?"; cat gpurun_out/bench_ur10.json; tail -3 gpurun_out/bench_ur10.err
timeout 600 python bench.py --config tiago --steps 5 > gpurun_out/bench_tiago.json 2> gpurun_out/bench_tiago.err; echo "tiago rc=$?"; cat gpurun_out/bench_tiago.json; tail -3 gpurun_out/bench_tiago.err
timeout 900 python bench.py --config talos --samples 200000 --steps 3 > gpurun_out/bench_talos.json 2> gpurun_out/bench_talos.err; echo "talos rc=$?"; cat gpurun_out/bench_talos.json; tail -3 gpurun_out/bench_talos.err
FG_WIDE_OCC2=1 timeout 900 python bench.py --config talos --samples 200000 --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_talos_occ2.json 2> gpurun_out/bench_talos_occ2.err; echo "talos occ2 rc=$?"; cat gpurun_out/bench_talos_occ2.json; tail -3 gpurun_out/bench_talos_occ2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
ls gpurun_out | wc -l
