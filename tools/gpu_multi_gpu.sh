#!/bin/bash
# Round-2 multi-GPU call: usage tools/gpu_multi_gpu.sh N   (run under gpurun --gpus N)
# weak scaling (1 M samples per GPU), the north-star strong-scaling run (10 M UR10 samples over N GPUs), Talos weak,
# the 2-rank NCCL parity test.
N=${1:-2}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
run() { # name, bench args
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N "$@" > gpurun_out/${name}_n$N.json 2> gpurun_out/${name}_n$N.err; echo "$name N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${name}_n$N.json').read().strip().splitlines()[-1])
    print('${name}', 'value', round(d['value']/1e6,2), 'M/s step', round(d['ms_per_step'],3), 'kernel', round(d['roofline']['kernel_ms'],3), 'e2e', d.get('e2e') and round(d['e2e']['value']/1e6,2), d.get('e2e') and round(d['e2e']['ms_per_step'],3), d.get('parity_check'))
except Exception as e: print('${name} ERR', e)
PY
}
run weak_ur10 --steps 10 --warmup 3
run strong10m_ur10 --total-samples 10000000 --steps 10 --warmup 3
run weak_talos --config talos --steps 3 --warmup 3 --no-parity-check
run weak_tiago --config tiago --steps 10 --warmup 3 --no-parity-check
if [ "$N" = "1" ]; then exit 0; fi
timeout 600 python -m pytest tests/test_gpu_parity_oracle.py -m gpu -q -k "nccl" 2>&1 | tail -3
