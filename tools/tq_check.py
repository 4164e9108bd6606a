"""GPU diagnostic of the tile-QR fused kernel: Gram identity R^T R = [W|tau]^T [W|tau] against the
materialised regressor, and a timing comparison with the previous kernel (FG_USE_TQ=0)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "figaroh-plus_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from figaroh_b200 import engine as eng  # noqa: E402
from figaroh_b200.model import FlatModel  # noqa: E402


def load(model):
    return FlatModel.from_json(open(os.path.join(ROOT, "tests", "golden", "models", model + ".json")).read())


def gram_check(model, flags, N, coupling=False, seed=0):
    fm = load(model)
    dm = eng.DeviceModel(fm)
    param = dict(is_joint_torques=True, is_external_wrench=False, force_torque=None, has_friction=flags[0],
                 has_actuator_inertia=flags[1], has_joint_offset=flags[2])
    cfg = eng.RegressorConfig.from_param(param, coupling=coupling)
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = (torch.rand((N, fm.nq), generator=g, device="cuda", dtype=torch.float64) - 0.5) * 6
    v = (torch.rand((N, fm.nv), generator=g, device="cuda", dtype=torch.float64) - 0.5) * 6
    a = (torch.rand((N, fm.nv), generator=g, device="cuda", dtype=torch.float64) - 0.5) * 12
    tau = torch.randn((fm.nv, N), generator=g, device="cuda", dtype=torch.float64)
    W = dm.materialize(cfg, q, v, a)
    cols = eng.structural_columns(dm, cfg)
    cfg2 = cfg.with_columns(cols)
    Wc = W[:, torch.as_tensor(cols.astype(np.int64), device="cuda")]
    M = torch.cat([Wc, tau.reshape(-1, 1)], dim=1)
    G = (M.T @ M).cpu().numpy()
    out = {}
    for name, env in (("tq", "1"), ("old", "0")):
        os.environ["FG_USE_TQ"] = env
        st = eng.TsqrState(len(cols))
        st.accumulate_samples(dm, cfg2, q, v, a, tau)
        torch.cuda.synchronize()
        p1 = st.p + 1
        pk = st.export().cpu().numpy()
        R = pk[:p1 * p1].reshape(p1, p1)
        perm = np.concatenate([st.get_perm(), [st.p]])
        G2 = R.T @ R
        Gp = G[np.ix_(perm, perm)]
        # columns that are pure rounding noise (gravity through a 1e-17 rotation entry) only have to stay noise
        dg = np.maximum(np.diag(Gp), 1e-24 * np.diag(Gp).max())
        scale = np.sqrt(np.outer(dg, dg))
        err = np.abs(G2 - Gp) / scale
        csq = st.colsq()
        cerr = np.abs(csq - np.diag(G)) / np.maximum(np.diag(G), 1e-24 * np.diag(G).max())
        out[name] = (err.max(), cerr.max())
        print(f"{model} flags={flags} N={N} {name}: max rel Gram err {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"
              f", colsq err {cerr.max():.3e}, tril max {np.abs(np.tril(R, -1)).max():.1e}, perm identity={np.array_equal(perm, np.arange(p1))}")
        if err.max() > 1e-9 and name == "tq" and N <= 4000:
            Mp = M.cpu().numpy()[:, perm]
            Rref = np.linalg.qr(Mp, mode="r")
            dR = np.abs(np.abs(R) - np.abs(Rref)) / (np.abs(Rref).max())
            rows_bad = np.flatnonzero(dR.max(axis=1) > 1e-9)
            cols_bad = np.flatnonzero(dR.max(axis=0) > 1e-9)
            print("  |R| vs numpy: first bad rows", rows_bad[:12].tolist(), "first bad cols", cols_bad[:12].tolist())
            if len(rows_bad):
                r = rows_bad[0]
                print("  row", r, "ours", np.round(R[r, r:r + 10], 6).tolist())
                print("  row", r, "ref ", np.round(Rref[r, r:r + 10], 6).tolist())
        if err.max() > 1e-9 and name == "tq":
            bad = np.argwhere(err > 1e-9)
            print("  bad entries (state order):", bad[:20].tolist(), "count", len(bad))
            print("  perm:", perm.tolist())
    os.environ["FG_USE_TQ"] = "1"
    return out


def timing(model, N, reps=5):
    fm = load(model)
    dm = eng.DeviceModel(fm)
    param = dict(is_joint_torques=True, is_external_wrench=False, force_torque=None, has_friction=True,
                 has_actuator_inertia=True, has_joint_offset=True)
    cfg = eng.RegressorConfig.from_param(param)
    q = (torch.rand((N, fm.nq), device="cuda", dtype=torch.float64) - 0.5) * 6
    v = (torch.rand((N, fm.nv), device="cuda", dtype=torch.float64) - 0.5) * 6
    a = (torch.rand((N, fm.nv), device="cuda", dtype=torch.float64) - 0.5) * 12
    tau = torch.randn((fm.nv, N), device="cuda", dtype=torch.float64)
    cols = eng.structural_columns(dm, cfg)
    cfg2 = cfg.with_columns(cols)
    for name, env in (("tq", "1"), ("old", "0")):
        os.environ["FG_USE_TQ"] = env
        st = eng.TsqrState(len(cols))
        st.set_timing(True)
        for _ in range(2):
            st.reset()
            st.accumulate_samples(dm, cfg2, q, v, a, tau)
        torch.cuda.synchronize()
        st.set_timing(True)
        t0 = time.perf_counter()
        for _ in range(reps):
            st.reset()
            st.accumulate_samples(dm, cfg2, q, v, a, tau)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / reps * 1e3
        ms, n = st.get_timing()
        print(f"timing {model} N={N} {name}: fold kernel {ms / n:.3f} ms, whole accumulate {wall:.3f} ms")
    os.environ["FG_USE_TQ"] = "1"


if __name__ == "__main__":
    torch.cuda.set_device(0)
    if len(sys.argv) > 1 and sys.argv[1] == "prof":
        timing("ur10", int(sys.argv[2]) if len(sys.argv) > 2 else 300000, reps=1)
        sys.exit(0)
    gram_check("ur10", (True, True, True), 64)
    gram_check("ur10", (True, True, True), 1003)
    gram_check("ur10", (False, False, False), 517)
    gram_check("tx40", (True, False, True), 2000, coupling=True)
    gram_check("ur10", (True, True, True), 200000)
    timing("ur10", 1000000)
