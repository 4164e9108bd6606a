#!/bin/bash
# Final validation of the shipped build (1 GPU): build, smoke, GPU suite, bench lines, materialise probe + capture, sanitizers
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke_final.log 2>&1; echo "build+smoke rc=$?"; tail -1 gpurun_out/smoke_final.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_ref.json 2>/dev/null
timeout 900 python bench.py > gpurun_out/bench_r02_ur10.json 2> gpurun_out/bench_r02_ur10.err; echo "ur10 rc=$?"
timeout 900 python bench.py --config tiago > gpurun_out/bench_r02_tiago.json 2> gpurun_out/bench_r02_tiago.err; echo "tiago rc=$?"
timeout 1500 python bench.py --config talos --steps 5 > gpurun_out/bench_r02_talos.json 2> gpurun_out/bench_r02_talos.err; echo "talos rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r02_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,3),'M/s', 'step',round(d['ms_per_step'],3), 'kernel', (d.get('roofline') or {}).get('kernel_ms'), 'frac', (d.get('roofline') or {}).get('frac'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'), (d.get('parity_check') or {}).get('phi_b_max_rel_err'), (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(f,'ERR',e)
PY
timeout 300 python tools/probe_materialize.py 1000000 > gpurun_out/materialize_r02.json 2>/dev/null; cat gpurun_out/materialize_r02.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matz_kernel -s 3 -c 1 -f -o gpurun_out/prof_matz python tools/probe_materialize.py 1000000 > /dev/null 2>&1; echo "rc=$?"
echo "== sanitizers"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity_oracle.py tests/test_gpu_kernels_r2.py tests/test_gpu_multi.py -m gpu -q -x -k "snapped or talos or tiago_active or cooperative_tsqr_matches_lapack or materialise or simulated_shards" > gpurun_out/memcheck_final_r02.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck_final_r02.log
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_parity_oracle.py tests/test_gpu_kernels_r2.py tests/test_gpu_multi.py -m gpu -q -x -k "snapped or talos_external_wrench_vs_oracle or 70-200 or ur10-77 or simulated_shards" > gpurun_out/synccheck_final_r02.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/synccheck_final_r02.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity_oracle.py tests/test_gpu_kernels_r2.py -m gpu -q -x -k "snapped or 70-200 or ur10-77" > gpurun_out/racecheck_final_r02.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/racecheck_final_r02.log
