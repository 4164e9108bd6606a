import sys, json, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, "."); sys.path.insert(0, "figaroh-plus_b200")
from conftest import *
from oracle import ref_path
from figaroh_b200 import engine as eng
cases = {c["name"]: c for c in json.load(open("tests/golden/structural.json"))}
c = cases["ur10_fr_off"]
fm = load_flat("ur10"); om = load_oracle_model("ur10")
dm = eng.DeviceModel(fm)
cfg = eng.RegressorConfig.from_param(c["param"])
for N in (1000, 999, 64, 4, 8, 28, 2000):
    q, v, a = rand_inputs(fm, N, c["seed"], tuple(tuple(x) for x in c["lo_hi"]))
    W = ref_path.build_regressor_basic(om, q, v, a, c["param"], per_sample=False)
    sq = (W*W).sum(0); keep = [i for i in range(84) if sq[i] >= 1e-6]
    R = np.linalg.qr(W[:, keep], mode="r"); d = np.abs(np.diag(R))
    st = eng.TsqrState(84).accumulate_samples(dm, cfg, q, v, a, None)
    sqg = st.colsq()[:84]
    dg = st.reduce(keep)[:len(keep)]
    Wg = dm.materialize(cfg, q, v, a)
    st2 = eng.TsqrState(84).accumulate_dense(Wg, None)
    dg2 = st2.reduce(keep)[:len(keep)]
    k = min(len(d), len(dg))
    rel = np.abs(dg[:k]/d[:k]-1) * (d[:k]>1e-8); rel2 = np.abs(dg2[:k]/d[:k]-1) * (d[:k]>1e-8)
    worst = np.argsort(-rel)[:4]
    print(N, "worst fused", [(int(i), keep[i], float(d[i]), float(dg[i]), float(dg2[i])) for i in worst])
    print(N, "colsq relerr", np.abs(sqg - sq).max() / sq.max(), "diag fused", np.abs(dg[:k]/d[:k]-1)[d[:k]>1e-8].max(),
          "diag dense", np.abs(dg2[:k]/d[:k]-1)[d[:k]>1e-8].max(), "argmax", np.argmax(np.abs(dg[:k]/d[:k]-1) * (d[:k]>1e-8)))
