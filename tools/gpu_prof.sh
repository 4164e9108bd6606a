#!/bin/bash
# ncu full capture of one fused-kernel launch of the bench command.  usage: tools/gpu_prof.sh <kernel-regex> <out-name> [samples]
mkdir -p gpurun_out
K=${1:-gq_accumulate}; O=${2:-prof_gq}; S=${3:-300000}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/$O \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --samples $S > gpurun_out/${O}_bench.log 2>&1; echo "ncu full rc=$?"
