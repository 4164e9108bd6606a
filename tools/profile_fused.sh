#!/bin/bash
# ncu capture of the fused kernel for the bench command (run under gpurun).  Outputs in gpurun_out/.
# usage: tools/profile_fused.sh <kernel-regex> <out-name> [samples]
set -x
mkdir -p gpurun_out
K=${1:-bq_accumulate}; O=${2:-prof_bq}; S=${3:-300000}
ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/$O \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --samples $S > gpurun_out/${O}_bench.log 2>&1
ls -la gpurun_out | tail -5
