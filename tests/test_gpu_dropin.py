"""GPU: the figaroh.tools drop-in modules behave like the reference's (same call sequences
as examples/ur10/identification.py:71-91, examples/tiago/identification.py:296-348 and
examples/staubli_TX40/identification.py:89-100,237-240), checked against the oracle."""
import json

import numpy as np
import pytest

from conftest import TX40_PARAM, golden_path, load_flat, load_oracle_model, make_param, rand_inputs

pytestmark = pytest.mark.gpu


class _Robot:
    """What the scripts hold: an object with .model / .data (figaroh.tools.robot.Robot without a URDF)."""

    def __init__(self, name):
        from figaroh.tools.robot import Robot
        from figaroh_b200.model import model_from_flat
        self._r = Robot.__new__(Robot)
        self._r.model = model_from_flat(load_flat(name))
        self._r.data = self._r.model.createData()
        self.model, self.data = self._r.model, self._r.data

    def get_standard_parameters(self, param):
        return self._r.get_standard_parameters(param)


def test_ur10_structural_script_sequence():
    from figaroh.tools.qrdecomposition import get_baseIndex, get_baseParams, build_baseRegressor, cond_num
    from figaroh.tools.regressor import build_regressor_basic, build_regressor_reduced, get_index_eliminate
    from oracle import ref_path
    robot, om = _Robot("ur10"), load_oracle_model("ur10")
    param = make_param(False, False, False)
    q, v, a = rand_inputs(load_flat("ur10"), 1000, 4)
    std = robot.get_standard_parameters(param)
    W = build_regressor_basic(robot, q, v, a, param)
    assert isinstance(W, np.ndarray) and W.dtype == np.float64 and W.flags.c_contiguous and W.shape == (6000, 84)
    Wref = ref_path.build_regressor_basic(om, q, v, a, param, per_sample=False)
    assert np.abs(W - Wref).max() <= 1e-9 * np.abs(Wref).max()
    idx_e, params_r = get_index_eliminate(W, std, 1e-6)
    assert (idx_e, params_r) == ref_path.get_index_eliminate(Wref, std, 1e-6)
    W_e = build_regressor_reduced(W, idx_e)
    assert W_e.shape == (6000, 49)
    W_b, params_base, idx_base = get_baseParams(W_e, params_r, std)
    rb = ref_path.get_baseParams(ref_path.build_regressor_reduced(Wref, idx_e), params_r, std)
    assert params_base == rb[1] and idx_base == rb[2] and len(idx_base) == 36
    assert np.allclose(W_b, rb[0])
    assert get_baseIndex(W_e, params_r) == tuple(idx_base)
    assert np.array_equal(build_baseRegressor(W_e, idx_base), W_e[:, idx_base])
    c_ref = np.linalg.cond(W_b)
    assert abs(cond_num(W_b) / c_ref - 1) < 1e-8
    assert abs(cond_num(W_b, "max_over_min_sigma") / ref_path.cond_num(W_b, "max_over_min_sigma") - 1) < 1e-7
    with pytest.raises(np.linalg.LinAlgError):
        cond_num(W_b, "fro")
    with pytest.raises(AssertionError):
        get_baseIndex(W_e, params_r[:-1])


def test_tx40_coupling_and_eliminate():
    from figaroh.tools.qrdecomposition import double_QR
    from figaroh.tools.regressor import add_coupling_TX40, build_regressor_basic, eliminate_non_dynaffect
    from oracle import ref_path
    robot, om = _Robot("tx40"), load_oracle_model("tx40")
    param = dict(TX40_PARAM)
    N = 800
    q, v, a = rand_inputs(load_flat("tx40"), N, 6, ((-6, 6), (-10, 10), (-30, 30)))
    std = robot.get_standard_parameters(param)
    std.update(Iam6=param["Iam6"], fvm6=param["fvm6"], fsm6=param["fsm6"])
    W = build_regressor_basic(robot, q, v, a, param)
    W = add_coupling_TX40(W, robot.model, robot.data, N, 6, 6, 7, q, v, a)
    Wref = ref_path.add_coupling_TX40(ref_path.build_regressor_basic(om, q, v, a, param, per_sample=False), N, v, a)
    assert W.shape == (6 * N, 87) and np.abs(W - Wref).max() <= 1e-9 * np.abs(Wref).max()
    rng = np.random.default_rng(0)
    tau = Wref @ np.array(list(std.values()), float) + rng.normal(0, 0.05, 6 * N)
    W_e, params_r = eliminate_non_dynaffect(W, std, 1e-6)
    W_e_ref, params_r_ref = ref_path.eliminate_non_dynaffect(Wref, std, 1e-6)
    assert params_r == params_r_ref and W_e.shape == W_e_ref.shape
    out = double_QR(tau, W_e, params_r, std)
    ref = ref_path.double_QR(tau, W_e_ref, params_r_ref, std)
    assert len(out) == 5 and out[2] == ref[2] and len(out[2]) == 61
    assert list(out[1].keys()) == list(ref[1].keys())
    np.testing.assert_allclose(out[3], ref[3], atol=2e-6)
    np.testing.assert_allclose(out[4], ref[4], atol=1e-5)
    out4 = double_QR(tau, W_e, params_r)
    assert len(out4) == 4


def test_device_resident_pipeline_and_errors():
    import torch
    from figaroh.tools.qrdecomposition import double_QR
    from figaroh.tools.regressor import build_regressor_basic, build_regressor_reduced, get_index_eliminate
    from figaroh_b200.engine import DeviceMatrix
    robot = _Robot("ur10")
    param = make_param(True, False, True)
    q, v, a = (torch.from_numpy(x).cuda() for x in rand_inputs(load_flat("ur10"), 640, 8, ((-3, 3), (-3, 3), (-6, 6))))
    std = robot.get_standard_parameters(param)
    W = build_regressor_basic(robot, q, v, a, param)
    assert isinstance(W, DeviceMatrix)                       # stays in HBM for CUDA inputs
    idx_e, params_r = get_index_eliminate(W, std, 1e-6)
    W_e = build_regressor_reduced(W, idx_e)
    tau = torch.randn(W.shape[0], dtype=torch.float64, device="cuda")
    W_b, bp, names, phi_b = double_QR(tau, W_e, params_r)
    assert isinstance(W_b, DeviceMatrix) and W_b.shape == (W.shape[0], len(names))
    phi_ls = np.linalg.lstsq(W_b.numpy(), tau.cpu().numpy(), rcond=None)[0]
    np.testing.assert_allclose(phi_b, np.round(phi_ls, 6), atol=2e-6)
    # reference error behaviour
    ff = _Robot("talos_ff")
    qq, vv, aa = rand_inputs(load_flat("talos_ff"), 3, 0)
    with pytest.raises(ValueError):                           # regressor.py:53 cannot broadcast
        build_regressor_basic(ff, qq, vv, aa, make_param(False, False, False))
    with pytest.raises(ValueError, match="Please enter valid parameters"):
        build_regressor_basic(ff, qq, vv, aa, make_param(False, False, False, is_joint_torques=False,
                                                         is_external_wrench=True, force_torque=["F?"]))
    with pytest.raises(UnboundLocalError):
        build_regressor_basic(ff, qq, vv, aa, make_param(False, False, False, is_joint_torques=False))


def test_tiago_golden_csv_on_gpu():
    """The reference's TIAGo golden (44 expressions, phi_b to 6 decimals) GPU-only from the processed log to the
    base parameters: regressor by kernel (1), elimination on the column norms, the per-column zero-phase
    decimation of examples/tiago/identification.py:142-187 by fg_filtfilt_decimate (active joints' row blocks
    only), double_QR by kernels (2)+(3).  The scipy decimation of the same matrix is the cross-check."""
    import torch
    from figaroh.tools.qrdecomposition import double_QR
    from figaroh.tools.regressor import build_regressor_basic, build_regressor_reduced, get_index_eliminate
    from figaroh_b200 import signal as fsig
    from figaroh_b200.engine import DeviceMatrix
    from test_oracle_golden import tiago_decimate, tiago_inputs
    meta = json.load(open(golden_path("tiago_kat.json")))
    robot = _Robot("tiago")
    param = make_param(False, False, False)
    p, v, a, tau, idxv = tiago_inputs()
    N = len(p)
    std = robot.get_standard_parameters(param)
    pd, vd, ad = (torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (p, v, a))
    W = build_regressor_basic(robot, pd, vd, ad, param)            # DeviceMatrix: stays in HBM
    assert isinstance(W, DeviceMatrix)
    idx_e, params_r = get_index_eliminate(W, std, tol_e=0.001)
    assert idx_e == meta["idx_e"]
    W_e = build_regressor_reduced(W, idx_e)
    rows = torch.cat([torch.arange(int(i) * N, (int(i) + 1) * N) for i in idxv]).cuda()
    W_act = W_e.t.index_select(0, rows)                               # the active joints' row blocks
    W_rf = fsig.decimate_blocks(W_act, len(idxv), q=10)
    tau_rf = fsig.decimate(torch.from_numpy(np.ascontiguousarray(tau)).cuda(), 10).t().reshape(-1)
    W_rf_ref, tau_rf_ref = tiago_decimate(W_e.numpy(), tau, idxv, N)
    assert W_rf.shape == W_rf_ref.shape
    assert np.abs(W_rf.cpu().numpy() - W_rf_ref).max() <= 1e-9 * np.abs(W_rf_ref).max()
    assert np.abs(tau_rf.cpu().numpy() - tau_rf_ref).max() <= 1e-9 * np.abs(tau_rf_ref).max()
    W_b, bp, base, phi_b, phi_std = double_QR(tau_rf, DeviceMatrix(W_rf), params_r, std)
    assert base == [g[0] for g in meta["golden_csv"]]
    np.testing.assert_allclose(phi_b, [g[1] for g in meta["golden_csv"]], atol=1.01e-6)
    assert np.mean(phi_b == np.array([g[1] for g in meta["golden_csv"]])) > 0.9


def test_signal_ops_match_scipy():
    """medfilt / filtfilt / gradient / decimate of figaroh_b200.signal against scipy / numpy (per column)."""
    from scipy import signal
    from figaroh_b200 import signal as fsig
    rng = np.random.default_rng(5)
    N, nc = 5870, 7
    t = np.cumsum(rng.uniform(0.009, 0.011, N))
    x = np.sin(2 * np.pi * 0.3 * t)[:, None] * rng.uniform(0.5, 2, nc) + 0.05 * rng.normal(size=(N, nc))
    x[100:103, 2] += 5.0  # spikes for the median filter
    m = fsig.medfilt(x, 5)
    assert np.array_equal(m, np.stack([signal.medfilt(x[:, j], 5) for j in range(nc)], 1))
    b, a = signal.butter(4, 2 / 50, "low")
    f = fsig.filtfilt(b, a, m, padlen=3 * (max(len(a), len(b)) - 1))
    fref = np.stack([signal.filtfilt(b, a, m[:, j], padtype="odd", padlen=3 * (max(len(a), len(b)) - 1)) for j in range(nc)], 1)
    assert np.abs(f - fref).max() <= 1e-9 * np.abs(fref).max()
    g = fsig.gradient(f, t)
    gref = np.stack([np.gradient(fref[:, j]) / np.gradient(t) for j in range(nc)], 1)
    assert np.abs(g - gref).max() <= 1e-9 * np.abs(gref).max()
    assert np.abs(fsig.gradient(x[:, 0]) - np.gradient(x[:, 0])).max() <= 1e-12
    d = fsig.decimate(x, 10)
    dref = np.stack([signal.decimate(x[:, j], q=10, zero_phase=True) for j in range(nc)], 1)
    assert d.shape == dref.shape == ((N + 9) // 10, nc)
    assert np.abs(d - dref).max() <= 1e-9 * np.abs(dref).max()
    d1 = fsig.decimate(x[:, 3], 10)
    assert d1.shape == (dref.shape[0],) and np.abs(d1 - dref[:, 3]).max() <= 1e-9 * np.abs(dref).max()
    with pytest.raises(Exception):
        fsig.medfilt(x, 4)


def test_qr_pivoting_matches_reference_semantics():
    """QR_pivoting (qrdecomposition.py:24-86): pivot order, rank quirk, expressions and values against the
    oracle's restatement (scipy.linalg.qr(pivoting=True) on the tall matrix)."""
    from figaroh.tools.qrdecomposition import QR_pivoting
    from oracle import ref_path
    rng = np.random.default_rng(11)
    M, p = 4000, 12
    # well separated column norms (no pivoting ties), three exactly dependent columns
    W = rng.normal(size=(M, p)) * np.array([9.0, 1.0, 7.5, 3.0, 12.0, 0.5, 2.0, 5.0, 1.5, 4.0, 6.0, 0.8])
    W[:, 3] = 0.5 * W[:, 0] - 0.25 * W[:, 4]
    W[:, 8] = 2.0 * W[:, 1]
    W[:, 10] = W[:, 2] + 0.5 * W[:, 7]  # (equal weights would make the two residuals an exact pivoting tie)
    names = [f"p{i}" for i in range(p)]
    tau = W @ rng.normal(size=p) + 0.01 * rng.normal(size=M)
    W_b, bp = QR_pivoting(tau, W, names)
    W_b_ref, bp_ref = ref_path.QR_pivoting(tau, W, names)
    def norm(expr):  # the order of the rank-deficient tail of the pivoting is rounding noise: compare term sets
        tok = expr.split(" ")
        return tok[0], sorted(zip(tok[1::2], tok[2::2]))
    assert [norm(k) for k in bp.keys()] == [norm(k) for k in bp_ref.keys()]
    np.testing.assert_allclose(np.array(list(bp.values())), np.array(list(bp_ref.values())), atol=2e-6)
    assert W_b.shape == W_b_ref.shape
    np.testing.assert_allclose(W_b, W_b_ref, rtol=1e-9, atol=1e-9)
    # full column rank: the reference's rank stays 0 (quirk Q5) -> empty result
    Wf = rng.normal(size=(500, 5))
    W_b0, bp0 = QR_pivoting(Wf @ np.ones(5), Wf, names[:5])
    assert W_b0.shape == (500, 0) and bp0 == {}


def test_preprocessing_filter_and_differentiation_match_reference():
    """low_pass_filter_data + calculate_first_second_order_differentiation (identification_tools.py:334-424)
    on the GPU against the oracle's restatement (scipy.signal.filtfilt, np.gradient): <= 1e-9 relative."""
    import torch
    from figaroh.identification.identification_tools import (calculate_first_second_order_differentiation,
                                                             low_pass_filter_data)
    from oracle import pre
    rng = np.random.default_rng(17)
    for N, ncols, ts, fc, nb in ((5000, 6, 1e-3, 100.0, 5), (777, 3, 1e-2, 10.0, 4), (40, 2, 1e-2, 20.0, 2)):
        t = np.arange(N)[:, None] * ts
        x = np.sin(2 * np.pi * (0.7 + np.arange(ncols)) * t) + 0.05 * rng.normal(size=(N, ncols)) + 3.0
        param = dict(ts=ts, cut_off_frequency_butterworth=fc, is_joint_torques=True, is_external_wrench=False)
        ref = pre.low_pass_filter_data(x, param, nb)
        out = low_pass_filter_data(x, param, nb)
        assert out.shape == ref.shape and out.dtype == np.float64
        assert np.abs(out - ref).max() <= 1e-9 * np.abs(ref).max()
        out_d = low_pass_filter_data(torch.from_numpy(x).cuda(), param, nb)      # device in -> device out
        assert out_d.is_cuda and np.array_equal(out_d.cpu().numpy(), out)
        out1 = low_pass_filter_data(x[:, 0], param, nb)                          # 1-D signal
        assert out1.shape == (ref.shape[0],) and np.abs(out1 - ref[:, 0]).max() <= 1e-9 * np.abs(ref).max()

        class M:
            nq = nv = ncols
        qf = ref
        q_r, dq_r, ddq_r = pre.calculate_first_second_order_differentiation(ncols, qf, param)
        q_o, dq_o, ddq_o = calculate_first_second_order_differentiation(M, qf, param)
        assert q_o.shape == q_r.shape and dq_o.shape == dq_r.shape and ddq_o.shape == ddq_r.shape
        assert np.array_equal(q_o, q_r)
        assert np.abs(dq_o - dq_r).max() <= 1e-9 * np.abs(dq_r).max()
        assert np.abs(ddq_o - ddq_r).max() <= 1e-9 * np.abs(ddq_r).max()
        assert np.all(ddq_o[:, -1] == 0.0)                                       # the reference's range(nq - 1)
        dt = ts * (1.0 + 0.1 * rng.uniform(size=qf.shape[0] - 1))                # variable step
        _, dq_r2, ddq_r2 = pre.calculate_first_second_order_differentiation(ncols, qf, param, dt)
        _, dq_o2, ddq_o2 = calculate_first_second_order_differentiation(M, qf, param, dt)
        assert np.abs(dq_o2 - dq_r2).max() <= 1e-9 * np.abs(dq_r2).max()
        assert np.abs(ddq_o2 - ddq_r2).max() <= 1e-9 * np.abs(ddq_r2).max()


def test_get_torque_rand_matches_rnea_oracle():
    """figaroh.tools.randomdata.get_torque_rand (randomdata.py:93-147): regressor identity on the GPU ==
    the oracle's RNEA + the reference's friction / inertia / offset / coupled-wrist terms."""
    from figaroh.tools.randomdata import get_torque_rand
    from oracle import kernels
    robot, om = _Robot("tx40"), load_oracle_model("tx40")
    param = dict(TX40_PARAM)
    N = 300
    q, v, a = rand_inputs(load_flat("tx40"), N, 31, ((-3, 3), (-5, 5), (-10, 10)))
    v[3, 2] = 0.0  # sign(0) = 0
    tau = get_torque_rand(N, robot, q, v, a, param)
    nv = 6
    ref = np.zeros(nv * N)
    for i in range(N):
        t = kernels.rnea(om, q[i], v[i], a[i])
        for j in range(nv):
            ref[j * N + i] = t[j] + v[i, j] * param["fv"][j] + np.sign(v[i, j]) * param["fs"][j] \
                + param["Ia"][j] * a[i, j] + param["off"][j]
            if j == nv - 2:
                ref[j * N + i] += param["Iam6"] * v[i, nv - 1] + param["fvm6"] * v[i, nv - 1] \
                    + param["fsm6"] * np.sign(v[i, nv - 2] + v[i, nv - 1])
            if j == nv - 1:
                ref[j * N + i] += param["Iam6"] * v[i, nv - 2] + param["fvm6"] * v[i, nv - 2] \
                    + param["fsm6"] * np.sign(v[i, nv - 2] + v[i, nv - 1])
    assert tau.shape == ref.shape
    assert np.abs(tau - ref).max() <= 1e-9 * np.abs(ref).max()


def test_statistics_on_the_factor_match_reference():
    """relative_stdev and weigthed_least_squares (identification_tools.py:204-234, 291-331) from one TSQR pass
    on the GPU against the numpy restatement of the reference."""
    from figaroh.identification.identification_tools import relative_stdev, weigthed_least_squares
    from oracle import pre
    rng = np.random.default_rng(23)
    nq, n, r = 3, 400, 7
    W_b = rng.normal(size=(nq * n, r)) * np.array([1.0, 5.0, 0.3, 2.0, 8.0, 0.7, 1.5])
    phi_true = rng.normal(size=r) * 3
    noise = np.repeat([0.05, 0.5, 0.01], n)
    tau = W_b @ phi_true + noise * rng.normal(size=nq * n)
    phi_ls = np.round(np.linalg.lstsq(W_b, tau, rcond=None)[0], 6)
    np.testing.assert_allclose(relative_stdev(W_b, phi_ls, tau), pre.relative_stdev(W_b, phi_ls, tau), atol=0.011)
    param = dict(idx_tau_stop=[n, 2 * n, 3 * n])

    class R:
        class model:
            pass
    R.model.nq = nq
    tau_est = W_b @ phi_ls
    out = weigthed_least_squares(R, phi_ls, W_b, tau, tau_est, param)
    ref = pre.weigthed_least_squares(nq, phi_ls, W_b, tau, tau_est, param)
    np.testing.assert_allclose(out, ref, atol=2e-6)
    # rank-deficient weighted system: the reference's pinv gives the minimum-norm solution
    W_d = np.c_[W_b, 2.0 * W_b[:, 0]]
    phi_d = np.r_[phi_ls, 0.0]
    out_d = weigthed_least_squares(R, phi_d, W_d, tau, tau_est, param)
    ref_d = pre.weigthed_least_squares(nq, phi_d, W_d, tau, tau_est, param)
    np.testing.assert_allclose(out_d, ref_d, atol=5e-6)


def test_human_free_flyer_560_columns_drop_in():
    """examples/human/identification.py calls get_index_eliminate / get_baseParams on the full 14 x 40 = 560-column
    external-wrench regressor of the 41-joint human model: wider than one TSQR pass used to take."""
    from figaroh.tools.qrdecomposition import get_baseIndex
    from figaroh.tools.regressor import build_regressor_basic, build_regressor_reduced, get_index_eliminate
    from oracle import ref_path
    robot, om = _Robot("human_ff"), load_oracle_model("human_ff")
    fm = load_flat("human_ff")
    param = make_param(False, False, False, is_joint_torques=False, is_external_wrench=True, force_torque=["All"])
    N = 300
    q, v, a = rand_inputs(fm, N, 3, ((-1, 1), (-1, 1), (-2, 2)))
    std = robot.get_standard_parameters(param)
    W = build_regressor_basic(robot, q, v, a, param)
    Wref = ref_path.build_regressor_basic(om, q, v, a, param, per_sample=False)
    assert W.shape == Wref.shape == (6 * N, 14 * (fm.njoints - 1)) and W.shape[1] == 560
    assert np.abs(W - Wref).max() <= 1e-9 * np.abs(Wref).max()
    idx_e, params_r = get_index_eliminate(W, std, 1e-6)
    assert (idx_e, params_r) == ref_path.get_index_eliminate(Wref, std, 1e-6)
    W_e = build_regressor_reduced(W, idx_e)
    assert get_baseIndex(W_e, params_r) == ref_path.get_baseIndex(ref_path.build_regressor_reduced(Wref, idx_e), params_r)


def test_total_least_squares_payload_regressors():
    """build_total_regressor_current / _wrench (regressor.py:296-500): W_tot identical to the oracle's, the TLS
    vector (last right singular vector, from the R factor instead of the tall matrix) and the residue to 1e-9."""
    from figaroh.tools.regressor import build_total_regressor_current, build_total_regressor_wrench
    from oracle import ref_path
    rng = np.random.default_rng(4)
    ns, nj = 400, 6
    for flags in ((True, False), (False, True), (False, False)):
        width = 12 if flags[0] else (14 if flags[1] else 10)
        W_l = rng.normal(size=(nj * ns, width * nj))
        W_l[:, 5 * width + 2] = 0.0  # an eliminated payload column
        W_b_u, W_b_l = rng.normal(size=(nj * ns, 20)), rng.normal(size=(nj * ns, 20))
        I_u, I_l = rng.normal(size=nj * ns), rng.normal(size=nj * ns)
        std = {f"p{i}": 0.0 for i in range(width * nj)}
        param = dict(nb_samples=ns, has_friction=flags[0], has_actuator_inertia=flags[1], which_body_loaded=5, mass_load=3.0)
        W, V, r = build_total_regressor_current(W_b_u, W_b_l, W_l, I_u, I_l, std, param)
        W0, V0, r0 = ref_path.build_total_regressor_current(W_b_u, W_b_l, W_l, I_u, I_l, std, param)
        assert np.array_equal(W, W0)
        assert np.abs(V - V0).max() <= 1e-9 * np.abs(V0).max() and np.abs(r - r0).max() <= 1e-9 * np.abs(r0).max()
    W_l = rng.normal(size=(6 * ns, 10 * nj))
    W_b_u, W_b_l = rng.normal(size=(6 * ns, 25)), rng.normal(size=(6 * ns, 25))
    tu, tl = rng.normal(size=6 * ns), rng.normal(size=6 * ns)
    param = dict(which_body_loaded=5, mass_load=2.5)
    W, V, r = build_total_regressor_wrench(W_b_u, W_b_l, W_l, tu, tl, {}, param)
    W0, V0, r0 = ref_path.build_total_regressor_wrench(W_b_u, W_b_l, W_l, tu, tl, {}, param)
    assert np.array_equal(W, W0)
    assert np.abs(V - V0).max() <= 1e-9 * np.abs(V0).max() and np.abs(r - r0).max() <= 1e-9 * np.abs(r0).max()
