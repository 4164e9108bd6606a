#!/bin/bash
# Round-2 call A: generation-5 (ring) kernel bring-up: parity suites, A/B against generation 4, first profiles.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv
python __graft_entry__.py > gpurun_out/build.log 2>&1
echo "== smoke (ring kernel)"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest gpu (ring kernel)"
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity_oracle.py::test_fused_structural_index_sets_wide_models > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
echo "== new oracle tests on generation 4"
FG_USE_RQ=0 timeout 900 python -m pytest tests/test_gpu_parity_oracle.py -m gpu -q --deselect tests/test_gpu_parity_oracle.py::test_fused_structural_index_sets_wide_models > gpurun_out/pytest_oracle_gq.log 2>&1; echo "rc=$?"
tail -8 gpurun_out/pytest_oracle_gq.log
echo "== bench A/B"
FG_USE_RQ=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gq.json 2> gpurun_out/bench_gq.err; echo "gq rc=$?"
for npc in 2 3; do for ring in 256 512 1024; do
  FG_RQ_NPC=$npc FG_RQ_RING=$ring timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rq_p${npc}_r${ring}.json 2> gpurun_out/bench_rq_p${npc}_r${ring}.err; echo "rq npc=$npc ring=$ring rc=$?"
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']/1e6,1),'M/s', 'step',round(d['ms_per_step'],3),'kern',round(d['roofline']['kernel_ms'],3),'frac',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value']/1e6,1), d['config'].get('eliminated'), d['config'].get('base_parameters'))
    except Exception as e:
        print(f,'ERR',e)
PY
echo "== model probes (current path)"
timeout 300 python tools/probe_models.py tiago 200000 2>&1 | tail -1
timeout 300 python tools/probe_models.py talos 50000 2>&1 | tail -1
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full (1M launch)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rq_accumulate -s 3 -c 1 -f -o gpurun_out/prof_rq \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_rq_bench.log 2>&1; echo "ncu full rc=$?"
echo "== sanitizer"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest "tests/test_gpu_parity_oracle.py::test_fused_flags_on_snapped_models_vs_oracle" -m gpu -q -x > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest "tests/test_gpu_parity.py::test_fused_structural_index_sets" -m gpu -q -x > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"
tail -5 gpurun_out/racecheck.log
ls -la gpurun_out | head -50
