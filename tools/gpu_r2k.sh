#!/bin/bash
# Round-2 check call (1 GPU): GPU suite, materialise bandwidth (new kernel vs old), UR10 bench line.
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== pytest gpu"; timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== materialise"
timeout 300 python tools/probe_materialize.py 1000000 > gpurun_out/materialize_r02_matz.json 2>gpurun_out/materialize_r02_matz.err; cat gpurun_out/materialize_r02_matz.json
FG_USE_MATZ=0 timeout 300 python tools/probe_materialize.py 1000000 > gpurun_out/materialize_r02_old.json 2>/dev/null; cat gpurun_out/materialize_r02_old.json
echo "== bench"
timeout 900 python bench.py > gpurun_out/bench_r02_ur10.json 2> gpurun_out/bench_r02_ur10.err; echo "ur10 rc=$?"; cat gpurun_out/bench_r02_ur10.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matz_kernel -s 3 -c 1 -f -o gpurun_out/prof_matz python tools/probe_materialize.py 1000000 > /dev/null 2>&1; echo "rc=$?"
