"""CPU, container only: the numpy/C restatement against the UNMODIFIED reference files,
imported by path from /root/reference with the pinocchio shim (oracle/pinshim.py)."""
import numpy as np
import pytest

from conftest import TX40_PARAM, load_flat, make_param, rand_inputs
from oracle import pinshim, ref_path

pytestmark = pytest.mark.skipif(not pinshim.reference_available(), reason="/root/reference is not mounted")

URDF = {"ur10": "examples/ur10/data/robot.urdf", "tx40": "models/staubli_tx40_description/urdf/tx40_mdh_modified.urdf",
        "tiago": "examples/tiago/urdf/tiago_48_schunk.urdf", "talos_ff": "examples/talos/data/talos_full_v2.urdf"}


def _robot(name):
    rob = pinshim.load_reference("robot")
    return rob.Robot(pinshim.REF_ROOT + "/" + URDF[name], package_dirs=None, isFext=name.endswith("_ff"))


@pytest.mark.parametrize("name", ["ur10", "tx40", "tiago", "talos_ff"])
def test_urdf_loader_matches_committed_model(name):
    """The committed flattened models are what the URDF loader produces today."""
    from figaroh_b200.model import flatten_model
    fm, ref = flatten_model(_robot(name).model), load_flat(name)
    assert fm.names == ref.names
    for k in ("parent", "jtype", "idx_q", "idx_v", "axis", "plc_R", "plc_p", "mass", "lever", "inertia", "gravity"):
        assert np.array_equal(getattr(fm, k), getattr(ref, k)), k


@pytest.mark.parametrize("name,param,N,coupling", [
    ("ur10", make_param(True, True, True), 300, False),
    ("tx40", dict(TX40_PARAM), 300, True),
    ("tiago", make_param(True, False, True), 40, False),
    ("talos_ff", make_param(True, True, True, is_joint_torques=False, is_external_wrench=True,
                            force_torque=["Fx", "Mz"]), 150, False),
])
def test_pipeline_functions_equal_reference(name, param, N, coupling):
    reg, qrd = pinshim.load_reference("regressor"), pinshim.load_reference("qrdecomposition")
    robot = _robot(name)
    om = pinshim._oracle_model(robot.model)
    q, v, a = rand_inputs(load_flat(name), N, 9, ((-2, 2), (-3, 3), (-5, 5)))
    v[1, :] = 0.0
    W0 = reg.build_regressor_basic(robot, q, v, a, param)
    assert np.array_equal(W0, ref_path.build_regressor_basic(om, q, v, a, param, per_sample=True))
    assert np.array_equal(W0, ref_path.build_regressor_basic(om, q, v, a, param, per_sample=False))
    std0, std1 = robot.get_standard_parameters(param), ref_path.standard_parameters(om, param)
    assert list(std0.keys()) == list(std1.keys())
    assert np.array_equal(np.array(list(std0.values()), float), np.array(list(std1.values()), float))
    if coupling:
        W0 = reg.add_coupling_TX40(W0, robot.model, robot.data, N, robot.model.nq, robot.model.nv, robot.model.njoints, q, v, a)
        assert np.array_equal(W0, ref_path.add_coupling_TX40(W0[:, :-3], N, v, a))
        for k in ("Iam6", "fvm6", "fsm6"):
            std0[k] = std1[k] = param[k]
    ie0, pr0 = reg.get_index_eliminate(W0, std0, 1e-6)
    assert (ie0, pr0) == ref_path.get_index_eliminate(W0, std1, 1e-6)
    We = reg.build_regressor_reduced(W0, ie0)
    assert np.array_equal(We, ref_path.build_regressor_reduced(W0, ie0))
    We2, pr2 = reg.eliminate_non_dynaffect(W0, std0, 1e-6)
    We3, pr3 = ref_path.eliminate_non_dynaffect(W0, std1, 1e-6)
    assert np.array_equal(We2, We3) and pr2 == pr3
    rng = np.random.default_rng(1)
    tau = W0 @ np.array(list(std0.values()), float) + rng.normal(0, 0.1, W0.shape[0])
    r0, r1 = qrd.double_QR(tau, We, pr0, std0), ref_path.double_QR(tau, We, pr0, std1)
    assert np.array_equal(r0[0], r1[0]) and r0[2] == r1[2] and np.array_equal(r0[3], r1[3]) and np.array_equal(r0[4], r1[4])
    b0, b1 = qrd.get_baseParams(We, pr0, std0), ref_path.get_baseParams(We, pr0, std1)
    assert b0[1] == b1[1] and b0[2] == b1[2]
    assert qrd.get_baseIndex(We, pr0) == ref_path.get_baseIndex(We, pr0)
    assert np.array_equal(qrd.build_baseRegressor(We, b0[2]), ref_path.build_baseRegressor(We, b1[2]))
    m = min(We.shape[0], 400)
    p0, p1 = qrd.QR_pivoting(tau[:m], We[:m], pr0), ref_path.QR_pivoting(tau[:m], We[:m], pr0)
    assert list(p0[1].keys()) == list(p1[1].keys())
    for nt in (None, "max_over_min_sigma"):
        assert qrd.cond_num(r0[0], nt) == ref_path.cond_num(r0[0], nt)
    for f in (qrd.cond_num, ref_path.cond_num):  # numpy refuses the Frobenius condition number of a tall matrix
        with pytest.raises(np.linalg.LinAlgError):
            f(r0[0], "fro")


def test_free_flyer_in_joint_torque_mode_raises_like_reference():
    reg = pinshim.load_reference("regressor")
    robot = _robot("talos_ff")
    q, v, a = rand_inputs(load_flat("talos_ff"), 2, 0)
    with pytest.raises(ValueError):
        reg.build_regressor_basic(robot, q, v, a, make_param(False, False, False))
    with pytest.raises(ValueError):
        ref_path.build_regressor_basic(pinshim._oracle_model(robot.model), q, v, a, make_param(False, False, False))


def test_preprocessing_restatement_equals_reference():
    """oracle/pre.py == the unmodified identification_tools.py (identification_tools.py:334-424)."""
    from oracle import pre
    idt = pinshim.load_reference("identification_tools")
    robot = _robot("ur10")
    rng = np.random.default_rng(3)
    N = 400
    q = np.cumsum(rng.normal(size=(N, robot.model.nq)), axis=0) * 1e-2
    param = dict(ts=1e-2, cut_off_frequency_butterworth=10.0, is_joint_torques=True, is_external_wrench=False)
    f0, f1 = idt.low_pass_filter_data(q, param, 5), pre.low_pass_filter_data(q, param, 5)
    assert np.array_equal(f0, f1)
    for dt in (None, 1e-2 * (1 + 0.1 * rng.uniform(size=f0.shape[0] - 1))):
        r0 = idt.calculate_first_second_order_differentiation(robot.model, f0, param, dt)
        r1 = pre.calculate_first_second_order_differentiation(robot.model.nq, f0, param, dt)
        for x0, x1 in zip(r0, r1):
            assert np.array_equal(x0, x1)


def test_statistics_restatement_equals_reference():
    """oracle/pre.py relative_stdev / weigthed_least_squares == the unmodified identification_tools.py:204-331."""
    from oracle import pre
    idt = pinshim.load_reference("identification_tools")
    rng = np.random.default_rng(1)
    W = rng.normal(size=(300, 4))
    phi = rng.normal(size=4)
    tau = W @ phi + 0.1 * rng.normal(size=300)
    assert np.array_equal(idt.relative_stdev(W, phi, tau), pre.relative_stdev(W, phi, tau))

    class R:
        class model:
            nq = 3
    p = dict(idx_tau_stop=[100, 200, 300])
    assert np.array_equal(idt.weigthed_least_squares(R, phi, W, tau, W @ phi, p),
                          pre.weigthed_least_squares(3, phi, W, tau, W @ phi, p))


def test_total_least_squares_regressors_equal_reference():
    """regressor.py:296-500 (payload identification, total least squares) against the restatement."""
    reg = pinshim.load_reference("regressor")
    rng = np.random.default_rng(4)
    ns, nj = 50, 6
    for flags in ((True, False), (False, True), (False, False)):
        width = 12 if flags[0] else (14 if flags[1] else 10)
        W_l = rng.normal(size=(nj * ns, width * nj))
        W_b_u, W_b_l = rng.normal(size=(nj * ns, 20)), rng.normal(size=(nj * ns, 20))
        I_u, I_l = rng.normal(size=nj * ns), rng.normal(size=nj * ns)
        std = {f"p{i}": 0.0 for i in range(width * nj)}
        param = dict(nb_samples=ns, has_friction=flags[0], has_actuator_inertia=flags[1], which_body_loaded=5, mass_load=3.0)
        a = reg.build_total_regressor_current(W_b_u, W_b_l, W_l, I_u, I_l, std, param)
        b = ref_path.build_total_regressor_current(W_b_u, W_b_l, W_l, I_u, I_l, std, param)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    W_l = rng.normal(size=(6 * ns, 10 * nj))
    W_b_u, W_b_l = rng.normal(size=(6 * ns, 25)), rng.normal(size=(6 * ns, 25))
    tu, tl = rng.normal(size=6 * ns), rng.normal(size=6 * ns)
    param = dict(which_body_loaded=5, mass_load=2.5)
    a = reg.build_total_regressor_wrench(W_b_u, W_b_l, W_l, tu, tl, {}, param)
    b = ref_path.build_total_regressor_wrench(W_b_u, W_b_l, W_l, tu, tl, {}, param)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
