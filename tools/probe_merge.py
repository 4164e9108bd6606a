"""Time of the N-independent multi-GPU tail on one GPU: export, merge_packed of G gathered factors, solve (development probe)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "figaroh-plus_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from conftest import load_flat, make_param, rand_inputs
from figaroh_b200 import engine as eng

fm = load_flat("ur10"); dm = eng.DeviceModel(fm)
cfg0 = eng.RegressorConfig.from_param(make_param(True, True, True))
sel = eng.structural_columns(dm, cfg0); cfg = cfg0.with_columns(sel)
N = 20000
q, v, a = rand_inputs(fm, N, 3, ((-3, 3), (-3, 3), (-6, 6)))
tau = np.random.default_rng(1).normal(size=(6, N))
st = eng.TsqrState(len(sel)).accumulate_samples(dm, cfg, q, v, a, tau)
names = [f"c{i}" for i in range(dm.shape(cfg0)[1])]
def timeit(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = {"export_ms": timeit(lambda: st.export())}
for G in (2, 4, 8):
    packs = torch.cat([st.export() for _ in range(G)])
    merged = eng.TsqrState(len(sel))
    def f():
        merged.reset(); merged.merge_packed(packs, G)
    res[f"reset+merge_packed_{G}_ms"] = timeit(f)
res["solve_ms"] = timeit(lambda: eng.solve_state(st, names, tol_e=1e-6, col_map=sel, p_full=len(names)))
print(json.dumps(res))
