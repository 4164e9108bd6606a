#!/bin/bash
# ncu launch list (durations) of the bench command.  Outputs gpurun_out/launches.csv
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; echo "ncu list rc=$?"
