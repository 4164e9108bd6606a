#!/bin/bash
# One GPU round: parity tests, bench, ncu launch list and a full capture of the fused kernel.
# Run under gpurun; everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gq_accumulate -s 3 -c 1 -f -o gpurun_out/prof_gq \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --samples ${PROF_SAMPLES:-300000} > gpurun_out/prof_gq_bench.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
