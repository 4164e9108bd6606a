#!/bin/bash
# materialise kernels: parity (golden regressor entries, edge cases) + bandwidth, chain kernel on / off
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels_r2.py tests/test_gpu_dropin.py -m gpu -q -k "golden or materialise or structural_script or tx40 or torque_rand" 2>&1 | tail -4
for v in 1 0; do
for rep in 1 2; do FG_USE_MATZG=$v timeout 300 python tools/probe_materialize.py 1000000 2>&1 | tail -1 | cut -c1-200; done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:matzg_kernel -s 3 -c 1 -f -o gpurun_out/prof_matzg python tools/probe_materialize.py 1000000 > /dev/null 2>&1; echo "rc=$?"
