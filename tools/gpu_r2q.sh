#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cq_accumulate -s 1 -c 1 -f -o gpurun_out/prof_cq_talos python bench.py --config talos --samples 100000 --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check --no-e2e > /dev/null 2>&1; echo "rc=$?"
