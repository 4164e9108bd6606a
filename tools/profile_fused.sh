#!/bin/bash
# ncu passes for the bench command (run under gpurun).  Outputs in gpurun_out/.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --samples ${SAMPLES:-1000000} > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:accumulate_samples -s 3 -c 1 -f -o gpurun_out/prof_fused \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --samples ${SAMPLES:-1000000} > gpurun_out/prof_bench.log 2>&1
ls -la gpurun_out
