#!/bin/bash
# UR10: ring size A/B (slots of 62 KB)
for ring in 256 288 320 352 384 416 384 320; do
FG_RQ_RING=$ring timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity-check 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ring $ring', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],4))"
done
