#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
K="talos_external_wrench_vs_oracle or 70-200 or 1000-330 or wide_drop_in"
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity_oracle.py tests/test_gpu_kernels_r2.py -m gpu -q -x -k "$K" > gpurun_out/racecheck_cq_r02.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/racecheck_cq_r02.log
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_parity_oracle.py tests/test_gpu_kernels_r2.py -m gpu -q -x -k "$K" > gpurun_out/synccheck_cq_r02.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/synccheck_cq_r02.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity_oracle.py tests/test_gpu_kernels_r2.py -m gpu -q -x -k "$K" > gpurun_out/memcheck_cq_r02.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck_cq_r02.log
for rep in 1 2; do
for f in figaroh-plus_b200/lib/libfigaroh_b200.so figaroh-plus_b200/lib/variants/lib_rq_narrow.so; do
for c in ur10 tiago; do
  FG_B200_LIB=$PWD/$f timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity-check 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$f $c', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4))"
done; done; done
timeout 900 python bench.py --config talos --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_talos.json 2> gpurun_out/bench_r02_talos.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_talos.json').read().strip().splitlines()[-1]); print('talos', d['roofline']['kernel_ms'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['parity_check'])"
