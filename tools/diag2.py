"""Development diagnostic: new materialise kernel against the first one, fused accumulate against the dense path."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "figaroh-plus_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from conftest import load_flat, make_param, rand_inputs
from figaroh_b200 import engine as eng

fm = load_flat("ur10")
dm = eng.DeviceModel(fm)
cfg = eng.RegressorConfig.from_param(make_param(True, True, True))
for N in (20, 64, 100, 1000, 4001):
    q, v, a = rand_inputs(fm, N, 3, ((-3, 3), (-3, 3), (-6, 6)))
    os.environ["FG_USE_MATZ"] = "1"
    W1 = dm.materialize(cfg, q, v, a).cpu().numpy()
    os.environ["FG_USE_MATZ"] = "0"
    W0 = dm.materialize(cfg, q, v, a).cpu().numpy()
    os.environ["FG_USE_MATZ"] = "1"
    d = np.abs(W1 - W0)
    print("N", N, "max diff", d.max(), "nan", np.isnan(W1).sum(), "bad rows", np.unique(np.nonzero(d > 1e-9)[0] % N)[:10], "bad cols", np.unique(np.nonzero(d > 1e-9)[1])[:20])
    _, cols = dm.shape(cfg)
    tau = np.random.default_rng(1).normal(size=(6, N))
    st = eng.TsqrState(cols).accumulate_samples(dm, cfg, q, v, a, tau)
    c1 = st.colsq()
    c0 = np.concatenate([(W0 ** 2).sum(0), [(tau ** 2).sum()]])[: len(c1)]
    print("   fused colsq rel err", np.abs(c1 - c0[: len(c1)]).max() / c0.max(), "rows", st.rows)
