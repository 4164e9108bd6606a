"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): regressor entries and base parameters within
1e-9 relative; eliminated / base index sets bit-exact.
"""
import numpy as np
import pytest

from conftest import (TX40_PARAM, golden_path, load_flat, load_oracle_model, make_param, rand_inputs)

pytestmark = pytest.mark.gpu

REL = 1e-9


def _eng():
    from figaroh_b200 import engine
    return engine


def _relerr(a, b):
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / scale


@pytest.mark.parametrize("case,model,coupling", [
    ("ur10_on", "ur10", False), ("tx40_cpl", "tx40", True), ("tiago_on", "tiago", False),
    ("talos_ew", "talos_ff", False), ("talos_ew2", "talos_ff", False)])
def test_materialize_golden_cases(case, model, coupling):
    """Kernel (1) against regressor entries produced by the unmodified reference (tests/golden)."""
    eng = _eng()
    z = np.load(golden_path("regressor_cases.npz"))
    q, v, a, Wref = z[case + "_q"], z[case + "_v"], z[case + "_a"], z[case + "_W"]
    if case == "ur10_on":
        param = make_param(True, True, True)
    elif case == "tx40_cpl":
        param = dict(TX40_PARAM)
    elif case == "tiago_on":
        param = make_param(True, True, True)
    elif case == "talos_ew":
        param = make_param(False, False, False, is_joint_torques=False, is_external_wrench=True, force_torque=["All"])
    else:
        param = make_param(True, False, True, is_joint_torques=False, is_external_wrench=True, force_torque=["Fz", "Mx"])
    dm = eng.DeviceModel(load_flat(model))
    cfg = eng.RegressorConfig.from_param(param, coupling=coupling)
    W = dm.materialize(cfg, q, v, a).cpu().numpy()
    assert W.shape == Wref.shape
    # structural zeros must be exact zeros
    assert np.array_equal(W == 0.0, Wref == 0.0) or np.abs(W[Wref == 0.0]).max() < 1e-12
    err = np.abs(W - Wref).max() / np.abs(Wref).max()
    assert err < REL, err
    _assert_columns_close(W, Wref)


def _assert_columns_close(W, Wref):
    """Per-column relative error <= 1e-9; columns that are pure rounding noise in the reference
    (|entries| < 1e-12 of the matrix scale, e.g. gravity leaking through a 1e-17 rotation entry)
    only have to be equally negligible."""
    scale = np.abs(Wref).max()
    colscale = np.abs(Wref).max(axis=0)
    errs = np.abs(W - Wref).max(axis=0)
    noise = colscale < 1e-12 * scale
    assert (errs[~noise] <= REL * colscale[~noise]).all(), (errs[~noise] / colscale[~noise]).max()
    assert (np.abs(W[:, noise]).max(axis=0, initial=0.0) < 1e-12 * scale).all()


@pytest.mark.parametrize("model,N,param_kw", [
    ("ur10", 1003, dict(fr=True, ia=True, jo=True)),
    ("tx40", 517, dict(fr=True, ia=False, jo=True)),
    ("tiago", 130, dict(fr=True, ia=True, jo=True)),
])
def test_materialize_vs_oracle_random(model, N, param_kw):
    from oracle import ref_path
    eng = _eng()
    fm = load_flat(model)
    om = load_oracle_model(model)
    q, v, a = rand_inputs(fm, N, 5, ((-3, 3), (-3, 3), (-6, 6)))
    param = make_param(param_kw["fr"], param_kw["ia"], param_kw["jo"])
    Wref = ref_path.build_regressor_basic(om, q, v, a, param, per_sample=False)
    dm = eng.DeviceModel(fm)
    W = dm.materialize(eng.RegressorConfig.from_param(param), q, v, a).cpu().numpy()
    assert W.shape == Wref.shape
    _assert_columns_close(W, Wref)


@pytest.mark.parametrize("rows,p", [(1, 5), (7, 3), (300, 31), (1000, 32), (2075, 75), (5000, 96), (4096, 130), (900, 250),
                                    (700, 340)])
def test_dense_tsqr_matches_lapack(rows, p):
    """R^T R = W^T W and |diag R| equals LAPACK's for a generic full-rank matrix."""
    eng = _eng()
    rng = np.random.default_rng(rows + p)
    W = rng.normal(size=(rows, p)) * rng.uniform(0.1, 10, size=p)
    tau = rng.normal(size=rows)
    st = eng.TsqrState(p)
    st.accumulate_dense(W, tau)
    assert st.rows == rows
    sq = st.colsq()
    np.testing.assert_allclose(sq[:p], (W * W).sum(0), rtol=1e-12)
    np.testing.assert_allclose(sq[p], tau @ tau, rtol=1e-12)
    diag = st.reduce(np.arange(p))
    Rref = np.linalg.qr(np.c_[W, tau], mode="r")
    k = min(rows, p + 1)
    np.testing.assert_allclose(diag[:k], np.abs(np.diag(Rref))[:k], rtol=1e-9, atol=1e-12 * np.abs(Rref).max())


def test_dense_tsqr_incremental_and_ragged():
    """Folding in chunks (ragged sizes, an empty one) gives the same factor as one shot."""
    eng = _eng()
    rng = np.random.default_rng(3)
    W = rng.normal(size=(3001, 40))
    tau = rng.normal(size=3001)
    a = eng.TsqrState(40).accumulate_dense(W, tau)
    b = eng.TsqrState(40)
    for lo, hi in [(0, 1), (1, 1), (1, 777), (777, 3001)]:
        if hi > lo:
            b.accumulate_dense(W[lo:hi], tau[lo:hi])
    assert a.rows == b.rows == 3001
    np.testing.assert_allclose(a.reduce(np.arange(40)), b.reduce(np.arange(40)), rtol=1e-10)


def _check_solution(sol, ref_idx_e, ref_out, names, tol=REL):
    W_b, bp, base_names, phi_b = ref_out[:4]
    assert sol.idx_e == list(ref_idx_e)
    assert sol.params_base == base_names
    np.testing.assert_allclose(sol.phi_b, phi_b, rtol=0, atol=2e-6)


@pytest.mark.parametrize("name", ["ur10_off", "ur10_on", "ur10_fr_off", "tx40_on_coupled"])
def test_fused_structural_index_sets(name, structural_cases):
    """Fused kernel (1)+(2)+(3): eliminated and base index sets identical to the reference's."""
    eng = _eng()
    c = structural_cases[name]
    model = "tx40" if name.startswith("tx40") else "ur10"
    fm = load_flat(model)
    q, v, a = rand_inputs(fm, c["N"], c["seed"], tuple(tuple(x) for x in c["lo_hi"]))
    dm = eng.DeviceModel(fm)
    cfg = eng.RegressorConfig.from_param(c["param"], coupling=c["coupling"])
    _, cols = dm.shape(cfg)
    names = [f"c{i}" for i in range(cols)]
    st = eng.TsqrState(cols).accumulate_samples(dm, cfg, q, v, a, None)
    assert st.rows == c["N"] * fm.nv
    sq = st.colsq()[:cols]
    idx_e = [i for i in range(cols) if sq[i] < c["tol_e"]]
    assert idx_e == c["idx_e"]
    keep = [i for i in range(cols) if i not in set(idx_e)]
    diag = st.reduce(keep)
    idx_base = [i for i in range(len(keep)) if diag[i] > 1e-8]
    assert idx_base == c["idx_base"]
    d = diag[:len(keep)]
    # Margins: once a dependent column has been passed, the direction of its reflector is rounding
    # noise, so later |R_ii| are only reproducible to O(#dependent / rows) (see DESIGN.md "R is not
    # unique"); the classification is what must match, with the same order of magnitude of margin.
    # (the fused kernel eliminates the friction / inertia / offset columns first, so its noise directions
    # differ from LAPACK's: 10 % covers the O(#dependent / rows) spread at these small N)
    assert abs(d[d > 1e-8].min() / c["min_base_diag"] - 1) < 0.10
    assert d[d <= 1e-8].max() < 1e-8


def test_fused_matches_dense_and_oracle_ur10():
    """Fused path == materialise + dense path == oracle double_QR (phi_b within 1e-9 relative)."""
    from oracle import ref_path
    eng = _eng()
    fm, om = load_flat("ur10"), load_oracle_model("ur10")
    N = 2000
    q, v, a = rand_inputs(fm, N, 77, ((-3, 3), (-3, 3), (-6, 6)))
    # friction + offsets on, actuator inertia off: well conditioned
    param = make_param(True, False, True, fv=TX40_PARAM["fv"], fs=TX40_PARAM["fs"], off=TX40_PARAM["off"])
    std = ref_path.standard_parameters(om, param)
    names = list(std.keys())
    Wref = ref_path.build_regressor_basic(om, q, v, a, param, per_sample=False)
    rng = np.random.default_rng(0)
    tau = Wref @ np.array(list(std.values()), dtype=float)
    tau = tau + 0.01 * np.sqrt(np.mean(tau ** 2)) * rng.normal(size=tau.shape)
    idx_e, params_r = ref_path.get_index_eliminate(Wref, std, 1e-6)
    W_e = ref_path.build_regressor_reduced(Wref, idx_e)
    ref = ref_path.double_QR(tau, W_e, params_r, std)

    dm = eng.DeviceModel(fm)
    cfg = eng.RegressorConfig.from_param(param)
    sol = eng.identify(dm, cfg, q, v, a, tau, names, tol_e=1e-6, params_std=std)
    assert sol.idx_e == idx_e and sol.params_r == params_r
    assert sol.params_base == ref[2]
    # unrounded comparison against lstsq on the base columns
    Wb = W_e[:, sol.idx_base]
    phi_ls = np.linalg.lstsq(Wb, tau, rcond=None)[0]
    cond = np.linalg.cond(Wb)
    rel = np.abs(sol.phi_b_raw - phi_ls).max() / np.abs(phi_ls).max()
    assert rel < REL, (rel, cond)
    assert np.array_equal(sol.phi_b, ref[3]) or np.abs(sol.phi_b - ref[3]).max() <= 1e-6
    np.testing.assert_allclose(sol.phi_std, ref[4], atol=1e-5)
    # same through the materialised + dense route and through the host-buffer route
    W = dm.materialize(cfg, q, v, a)
    st2 = eng.TsqrState(W.shape[1]).accumulate_dense(W, tau)
    sol2 = eng.solve_state(st2, names, tol_e=1e-6, params_std=std)
    assert sol2.idx_e == sol.idx_e and sol2.params_base == sol.params_base
    np.testing.assert_allclose(sol2.phi_b_raw, sol.phi_b_raw, rtol=1e-9)
    sol3 = eng.identify(dm, cfg, q, v, a, tau, names, tol_e=1e-6, params_std=std, host_inputs=True)
    import torch
    torch.cuda.synchronize()
    np.testing.assert_allclose(sol3.phi_b_raw, sol.phi_b_raw, rtol=1e-12)
    # residual norm and relative standard deviation
    r2 = np.sum((tau - Wb @ phi_ls) ** 2)
    assert abs(sol.resid2 / r2 - 1) < 1e-8
    sd_ref = ref_path.relative_stdev(ref[0], ref[3], tau)
    sd = eng.relative_stdev_from_solution(sol)
    np.testing.assert_allclose(sd, sd_ref, atol=0.011)


def test_tx40_kat_double_qr():
    """Reference golden: the 60 base-parameter expressions of TX40_bp_5.csv from the decimated regressor."""
    import json
    eng = _eng()
    z = np.load(golden_path("tx40_kat.npz"))
    meta = json.load(open(golden_path("tx40_kat.json")))
    W_, tau_ = z["W_"], z["tau_"]
    names = meta["params_std_names"]
    st = eng.TsqrState(W_.shape[1]).accumulate_dense(W_, tau_)
    sol = eng.solve_state(st, names, tol_e=0.001)
    assert sol.params_r == meta["params_r"]
    assert sol.params_base == meta["params_base"]
    assert [g[0] for g in meta["golden_csv"]] == sol.params_base
    np.testing.assert_allclose(sol.phi_b, meta["phi_b"], atol=2e-6)


@pytest.mark.gpu
def test_fused_solve_matches_three_call_sequence():
    """fg_state_solve (one launch) == fg_state_colsq + fg_state_reduce + fg_state_regroup_solve."""
    eng = _eng()
    fm = load_flat("ur10")
    N = 3000
    q, v, a = rand_inputs(fm, N, 5, ((-3, 3), (-3, 3), (-6, 6)))
    param = make_param(True, True, True, fv=TX40_PARAM["fv"], fs=TX40_PARAM["fs"], Ia=TX40_PARAM["Ia"], off=TX40_PARAM["off"])
    dm = eng.DeviceModel(fm)
    cfg = eng.RegressorConfig.from_param(param)
    _, cols = dm.shape(cfg)
    rng = np.random.default_rng(3)
    tau = rng.normal(size=fm.nv * N)
    st = eng.TsqrState(cols).accumulate_samples(dm, cfg, q, v, a, tau)
    for tol_e, drop in ((1e-6, ()), (None, (0, 1, 2, 3, 4, 6, 7, 8, 9, 22, 23, 40))):
        fused = st.solve(tol_e, drop, 1e-8)
        assert fused is not None
        sq = st.colsq()[:cols]
        dropped = set(drop) | (set(i for i in range(cols) if sq[i] < tol_e) if tol_e is not None else set())
        keep = [i for i in range(cols) if i not in dropped]
        assert fused["keep"] == keep
        np.testing.assert_array_equal(fused["colsq"][:cols], sq)
        diag = st.reduce(keep)
        np.testing.assert_allclose(fused["diag"], diag, rtol=1e-12, atol=1e-300)
        base = [i for i in range(len(keep)) if diag[i] > 1e-8]
        regroup = [i for i in range(len(keep)) if not diag[i] > 1e-8]
        assert fused["base"] == base and fused["regroup"] == regroup
        beta, phi, R1, r2 = st.regroup_solve(base, regroup)
        np.testing.assert_allclose(fused["beta"], beta, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(fused["phi"], phi, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(np.abs(fused["R1"]), np.abs(R1), rtol=1e-10, atol=1e-12)
        assert abs(fused["resid2"] / r2 - 1) < 1e-10


def test_host_chunked_continuation_matches_device():
    """Host-buffer path (ramped chunks: first / continued / last launch of the fused kernel on the same
    per-warp factors, one merge at the end) == one device-resident launch."""
    eng = _eng()
    fm = load_flat("ur10")
    N = 120_000  # 32768 + 65536 + 21696 samples: three launches
    q, v, a = rand_inputs(fm, N, 21, ((-3, 3), (-3, 3), (-6, 6)))
    param = make_param(True, False, True, fv=TX40_PARAM["fv"], fs=TX40_PARAM["fs"], off=TX40_PARAM["off"])
    dm = eng.DeviceModel(fm)
    cfg = eng.RegressorConfig.from_param(param)
    _, cols = dm.shape(cfg)
    names = [f"c{i}" for i in range(cols)]
    rng = np.random.default_rng(2)
    tau = rng.normal(size=fm.nv * N)
    sol_d = eng.identify(dm, cfg, q, v, a, tau, names, tol_e=1e-6)
    sol_h = eng.identify(dm, cfg, q, v, a, tau, names, tol_e=1e-6, host_inputs=True)
    assert sol_h.rows == sol_d.rows == N * fm.nv
    assert sol_h.idx_e == sol_d.idx_e and sol_h.idx_base == sol_d.idx_base
    assert sol_h.params_base == sol_d.params_base
    np.testing.assert_allclose(sol_h.phi_b_raw, sol_d.phi_b_raw, rtol=1e-9, atol=1e-12)
    assert abs(sol_h.resid2 / sol_d.resid2 - 1) < 1e-10


def test_fused_tiago_active_joints_tree_matches_dense():
    """BASELINE config 3 shape: TIAGo, rows of the 8 active joints (torso + arm, examples/tiago/identification.py
    :157-166) on the 24-DoF tree (grippers and head branch off the chain: the generic fused kernel takes it).
    Fused path == materialise + row selection + dense TSQR, index sets identical, phi_b within 1e-9."""
    import torch
    eng = _eng()
    fm = load_flat("tiago")
    N = 700
    q, v, a = rand_inputs(fm, N, 9, ((-1.5, 1.5), (-2, 2), (-4, 4)))
    rows = [fm.idx_v[i] for i in range(13, 21)]  # torso_lift + arm_1..7
    active = set(rows)
    for j in range(fm.nv):  # the other joints rest, as in the recorded trajectories
        if j not in active:
            v[:, j] = 0.0
            a[:, j] = 0.0
    for i in range(1, fm.njoints):
        if fm.idx_v[i] not in active:
            iq = fm.idx_q[i]
            if fm.jtype[i] == 2:
                q[:, iq], q[:, iq + 1] = 1.0, 0.0
            else:
                q[:, iq] = 0.1
    param = make_param(True, True, True)
    dm = eng.DeviceModel(fm)
    cfg = eng.RegressorConfig.from_param(param, row_sel=rows)
    rps, cols = dm.shape(cfg)
    assert rps == 8
    names = [f"c{i}" for i in range(cols)]
    rng = np.random.default_rng(4)
    tau = rng.normal(size=rps * N)
    sol = eng.identify(dm, cfg, q, v, a, tau, names, tol_e=1e-3)
    W = dm.materialize(cfg, q, v, a)                       # all 24 row blocks, reference layout
    Wsel = W.view(fm.nv, N, cols)[torch.as_tensor(rows, device=W.device)].reshape(rps * N, cols).contiguous()
    st = eng.TsqrState(cols).accumulate_dense(Wsel, tau)
    ref = eng.solve_state(st, names, tol_e=1e-3)
    assert sol.idx_e == ref.idx_e and sol.idx_base == ref.idx_base and sol.params_base == ref.params_base
    cond = np.linalg.cond(ref.R1)
    np.testing.assert_allclose(sol.phi_b_raw, ref.phi_b_raw, rtol=1e-9 * max(1.0, cond / 1e3), atol=1e-12)


def test_fused_talos_external_wrench_matches_dense():
    """BASELINE config 4 shape: Talos with free-flyer, external-wrench rows ('All'), flags off.  The generic
    fused kernel (6 row groups on the root joint, 330 non-zero columns) == materialise + dense TSQR."""
    eng = _eng()
    fm = load_flat("talos_ff")
    N = 400
    q, v, a = rand_inputs(fm, N, 13, ((-1, 1), (-1, 1), (-1, 1)))
    param = dict(is_joint_torques=False, is_external_wrench=True, force_torque=["All"], has_friction=False,
                 has_actuator_inertia=False, has_joint_offset=False, fv=None, fs=None, Ia=None, off=None)
    dm = eng.DeviceModel(fm)
    cfg = eng.RegressorConfig.from_param(param)
    rps, cols = dm.shape(cfg)
    assert rps == 6 and cols == 14 * (fm.njoints - 1)
    names = [f"c{i}" for i in range(cols)]
    rng = np.random.default_rng(8)
    tau = rng.normal(size=rps * N)
    sol = eng.identify(dm, cfg, q, v, a, tau, names, tol_e=1e-6)
    W = dm.materialize(cfg, q, v, a)
    st = eng.TsqrState(cols).accumulate_dense(W, tau)
    ref = eng.solve_state(st, names, tol_e=1e-6)
    assert sol.idx_e == ref.idx_e and len(sol.idx_e) == 132
    assert sol.idx_base == ref.idx_base and len(sol.idx_base) == 234
    cond = np.linalg.cond(ref.R1)
    np.testing.assert_allclose(sol.phi_b_raw, ref.phi_b_raw, rtol=1e-9 * max(1.0, cond / 1e3), atol=1e-11)
