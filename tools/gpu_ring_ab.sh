#!/bin/bash
timeout 120 python tools/probe_h2d.py
# TIAGo: ring size A/B (the ring + factors exceed L2 at 512 slots: 8.4 GB of DRAM traffic per 1 M samples)
for rep in 1 2; do
for ring in 192 256 320 384 512; do
  FG_RQ_RING=$ring FG_RQ_CALIBRATE=2 timeout 300 python bench.py --config tiago --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-parity-check 2>gpurun_out/ab.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('tiago ring $ring', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4))"
  grep "rq calibration:" gpurun_out/ab.err
done; done
RX='regex:rq_|gq_|cq_|matz|fused_solve|merge_pairs|small_qr|gather_cols|upper_solve|tri_to|wide_|pivot|dense_to'
for c in ur10 tiago; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$RX" -c 3000 --csv --log-file gpurun_out/launches_$c.csv python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e > /dev/null 2>&1; echo "rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$RX" -c 3000 --csv --log-file gpurun_out/launches_talos.csv python bench.py --config talos --samples 200000 --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e > /dev/null 2>&1; echo "rc=$?"
wc -l gpurun_out/launches_*.csv
