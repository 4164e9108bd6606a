#!/bin/bash
# two GPUs: full suite (incl. the NCCL parity test), Talos probe, UR10 weak / strong scaling at N=2
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python tools/probe_models.py talos 60000 2>&1 | tail -1
timeout 300 python tools/probe_materialize.py 1000000 2>&1 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_ur10_n2.json 2> gpurun_out/bench_ur10_n2.err; echo "n2 rc=$?"; cut -c1-400 gpurun_out/bench_ur10_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu-baseline --total-samples 10000000 > gpurun_out/bench_ur10_strong_n2.json 2> gpurun_out/bench_ur10_strong_n2.err; echo "strong n2 rc=$?"; cut -c1-400 gpurun_out/bench_ur10_strong_n2.json
timeout 600 python bench.py --no-cpu-baseline --total-samples 10000000 > gpurun_out/bench_ur10_strong_n1.json 2> gpurun_out/bench_ur10_strong_n1.err; echo "strong n1 rc=$?"; cut -c1-400 gpurun_out/bench_ur10_strong_n1.json
