#!/bin/bash
for rep in 1 2; do
for cap in 15 16 17; do
FG_H2D_CAP=$cap timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ur10 cap 2^$cap', 'step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3), round(d['e2e']['value']/1e6,1), 'h2d', round(d['e2e']['h2d_gbs_per_rank'],1))"
done; done
