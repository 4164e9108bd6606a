#!/bin/bash
# cost-model sweep of the ring kernel (UR10): consumer warps per row group, producer SMs
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
run() { # name env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity-check --no-e2e > gpurun_out/bench_s_$name.json 2> gpurun_out/bench_s_$name.err
}
run base FG_X=0
for t in 60 80 130 160 200; do run trail$t FG_RQ_COST=250,1070,$t; done
for p in 900 1000 1150 1250; do run panel$p FG_RQ_COST=250,$p,100; done
for b in 0 500 900; do run base$b FG_RQ_COST=$b,1070,100; done
for c in 24 25 27 28; do run pc$c FG_RQ_NPCTA=$c; done
run ring512 FG_RQ_RING=512
run ring768 FG_RQ_RING=768
python - <<'PY'
import json,glob
rows=[]
for f in sorted(glob.glob('gpurun_out/bench_s_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        rows.append((d['roofline']['kernel_ms'], f, d['ms_per_step']))
    except Exception as e:
        print(f,'ERR',e)
for k,f,s in sorted(rows): print(round(k,3), round(s,3), f)
PY
