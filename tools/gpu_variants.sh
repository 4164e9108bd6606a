#!/bin/bash
# A/B bench of the kernel variants built into figaroh-plus_b200/lib/variants against the in-tree library
# (development aid; alternates the libraries to expose run-to-run drift)
mkdir -p gpurun_out
for rep in 1 2; do
for f in figaroh-plus_b200/lib/libfigaroh_b200.so figaroh-plus_b200/lib/variants/*.so; do
  FG_B200_LIB=$PWD/$f timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$f', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), d['clocks'])"
done; done | tee gpurun_out/variants.txt
