#!/bin/bash
# Round-2 evidence call (1 GPU): suites, bench lines of the three robots, ncu evidence, sanitizers, sweep.
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench lines"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_ref.json 2>/dev/null
timeout 900 python bench.py > gpurun_out/bench_r02_ur10.json 2> gpurun_out/bench_r02_ur10.err; echo "ur10 rc=$?"
timeout 900 python bench.py --config tiago > gpurun_out/bench_r02_tiago.json 2> gpurun_out/bench_r02_tiago.err; echo "tiago rc=$?"
timeout 1500 python bench.py --config talos --steps 5 > gpurun_out/bench_r02_talos.json 2> gpurun_out/bench_r02_talos.err; echo "talos rc=$?"
timeout 900 python bench.py --pageable --no-cpu-baseline --no-parity-check > gpurun_out/bench_r02_ur10_pageable.json 2> gpurun_out/bench_r02_ur10_pageable.err; echo "pageable rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r02_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,3),'M/s', 'step',round(d['ms_per_step'],3), 'frac', d.get('roofline',{}).get('frac'), 'e2e', (d.get('e2e') or {}).get('value'), d.get('parity_check'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(f,'ERR',e)
PY
timeout 300 python tools/probe_materialize.py 1000000 > gpurun_out/materialize_r02.json 2>/dev/null; cat gpurun_out/materialize_r02.json
echo "== ncu launch lists"
for c in ur10 tiago; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$c.csv python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e > /dev/null 2>&1; echo "rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_talos.csv python bench.py --config talos --samples 200000 --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e > /dev/null 2>&1; echo "rc=$?"
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rq_accumulate -s 7 -c 1 -f -o gpurun_out/prof_rq_ur10 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e > /dev/null 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rq_accumulate -s 7 -c 1 -f -o gpurun_out/prof_rq_tiago python bench.py --config tiago --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e > /dev/null 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cq_accumulate -s 3 -c 1 -f -o gpurun_out/prof_cq_talos python bench.py --config talos --samples 100000 --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check --no-e2e > /dev/null 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gq_merge -s 3 -c 1 -f -o gpurun_out/prof_merge python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e > /dev/null 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matz_kernel -s 3 -c 1 -f -o gpurun_out/prof_matz python tools/probe_materialize.py 1000000 > /dev/null 2>&1; echo "rc=$?"
echo "== sanitizers"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity_oracle.py tests/test_gpu_kernels_r2.py -m gpu -q -x -k "snapped or talos or tiago_active or cooperative_tsqr_matches_lapack or materialise" > gpurun_out/memcheck_r02.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_r02.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity_oracle.py tests/test_gpu_kernels_r2.py -m gpu -q -x -k "snapped or talos_external_wrench_vs_oracle or 70-200 or ur10-77" > gpurun_out/racecheck_r02.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck_r02.log
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_parity_oracle.py tests/test_gpu_kernels_r2.py -m gpu -q -x -k "snapped or talos_external_wrench_vs_oracle or 70-200 or ur10-77" > gpurun_out/synccheck_r02.log 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/synccheck_r02.log
echo "== sweep (1 GPU)"
timeout 2400 python tools/sweep.py --gpus 1 --cpu --out gpurun_out/sweep_n1.jsonl
ls gpurun_out | wc -l
