#!/bin/bash
# launch lists of the three robots (our kernels only) + full captures of the timed launches (calibration off, the
# calibrated producer count fixed, so that launch number 4 is the step's)
mkdir -p gpurun_out
for c in ur10 tiago; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fg::|cq_|tri_to' -c 3000 --csv --log-file gpurun_out/launches_$c.csv python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e > /dev/null 2>&1; echo "rc=$?"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'fg::|cq_|tri_to' -c 3000 --csv --log-file gpurun_out/launches_talos.csv python bench.py --config talos --samples 200000 --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e > /dev/null 2>&1; echo "rc=$?"
FG_RQ_CALIBRATE=0 FG_RQ_NPCTA=22 timeout 900 ncu --set full --clock-control none --import-source on -k regex:rq_accumulate -s 3 -c 1 -f -o gpurun_out/prof_rq_ur10 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e > /dev/null 2>&1; echo "rc=$?"
FG_RQ_CALIBRATE=0 FG_RQ_NPCTA=21 timeout 900 ncu --set full --clock-control none --import-source on -k regex:rq_accumulate -s 3 -c 1 -f -o gpurun_out/prof_rq_tiago python bench.py --config tiago --steps 1 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e > /dev/null 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cq_accumulate -s 4 -c 2 -f -o gpurun_out/prof_cq_talos python bench.py --config talos --samples 100000 --steps 1 --warmup 1 --no-cpu-baseline --no-parity-check --no-e2e > /dev/null 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
