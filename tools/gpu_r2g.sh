#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/probe_models.py talos 60000 2>&1 | tail -1
timeout 300 python tools/probe_materialize.py 1000000 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_talos.csv \
    python tools/probe_models.py talos 60000 > gpurun_out/launches_talos.log 2>&1; echo "ncu list rc=$?"
