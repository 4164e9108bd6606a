#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
for c in ur10 tiago talos; do
st=10; [ $c = talos ] && st=4
FG_RQ_CALIBRATE=2 timeout 900 python bench.py --config $c --steps $st --warmup 3 --no-cpu-baseline 2>gpurun_out/ab.err | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'step', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['ms_per_step'],3), round(d['e2e']['value']/1e6,1), 'h2d', round(d['e2e']['h2d_gbs_per_rank'],1), d['parity_check'].get('phi_b_max_rel_err'))"
grep "rq calibration:" gpurun_out/ab.err
done
# e2e only process (calibration from a short first chunk -> synthetic inputs)
timeout 300 python - <<'PY'
import sys, os, time
sys.path.insert(0, 'figaroh-plus_b200'); sys.path.insert(0, 'tests')
import numpy as np, torch
from conftest import load_flat, make_param, rand_inputs
from figaroh_b200 import engine as eng
fm = load_flat("ur10"); dm = eng.DeviceModel(fm)
cfg = eng.RegressorConfig.from_param(make_param(True, True, True))
N = 600000
q, v, a = rand_inputs(fm, N, 3, ((-3, 3), (-3, 3), (-6, 6)))
_, cols = dm.shape(cfg)
tau = np.random.default_rng(1).normal(size=(6, N))
qh, vh, ah, th = (torch.from_numpy(x).pin_memory() for x in (q, v, a, tau))
st_h = eng.TsqrState(cols).accumulate_samples_host(dm, cfg, qh, vh, ah, th)
st_d = eng.TsqrState(cols).accumulate_samples(dm, cfg, q, v, a, tau)
c1, c0 = st_h.colsq(), st_d.colsq()
print("host-only path vs device path colsq rel diff", np.abs(c1 - c0).max() / c0.max(), st_h.rows, st_d.rows)
PY
