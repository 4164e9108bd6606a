#!/bin/bash
# materialise kernel A/B: parity (golden regressor entries) + bandwidth per library variant
mkdir -p gpurun_out
for f in figaroh-plus_b200/lib/libfigaroh_b200.so figaroh-plus_b200/lib/variants/lib_matz_*.so; do
  echo "== $f"
  FG_B200_LIB=$PWD/$f timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k golden 2>&1 | tail -1
  for rep in 1 2; do FG_B200_LIB=$PWD/$f timeout 300 python tools/probe_materialize.py 1000000 2>&1 | tail -1 | cut -c1-200; done
done
