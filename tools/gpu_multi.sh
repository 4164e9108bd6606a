#!/bin/bash
# N-GPU run of the bench (torchrun, NCCL) + the NCCL parity test.  usage: tools/gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
cat gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
