"""CPU: host-side logic - C-ABI surface, flattening loader, config mapping, expression strings."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, TX40_PARAM, golden_path, load_flat, make_param


def test_library_exports_every_declared_symbol():
    """libfigaroh_b200.so loads without a GPU and exports every function include/*.h declares."""
    from figaroh_b200 import _lib
    header = open(os.path.join(ROOT, "include", "figaroh_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(fg_[a-z0-9_]+)\s*\(", body))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(L, sym), sym
    assert _lib.lib().fg_abi_version() == 1
    assert _lib.lib().fg_state_packed_doubles(84) == 85 * 85 + 2 * 85


def test_no_cpu_fallback_without_gpu():
    import torch
    from figaroh_b200 import engine
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        engine.DeviceModel(load_flat("ur10"))
    with pytest.raises(RuntimeError):
        engine.TsqrState(10)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "figaroh-plus_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle[./]", src, flags=re.M), os.path.join(dp, f)


def test_flat_model_json_roundtrip_and_rebuild():
    from figaroh_b200.model import FlatModel, flatten_model, model_from_flat
    fm = load_flat("talos_ff")
    fm2 = FlatModel.from_json(fm.to_json())
    m = model_from_flat(fm2)
    fm3 = flatten_model(m)
    for k in ("parent", "jtype", "idx_q", "idx_v", "axis", "plc_R", "plc_p", "mass", "lever", "inertia", "gravity"):
        assert np.array_equal(getattr(fm, k), getattr(fm3, k)), k
    assert (fm3.njoints, fm3.nq, fm3.nv) == (34, 39, 38)
    assert m.joints[1].shortname() == "JointModelFreeFlyer" and m.joints[1].nq == 7 and m.joints[1].nv == 6
    assert m.inertias.tolist()[1].mass == fm.mass[1]


def test_joint_orders_from_survey():
    t = load_flat("tiago")
    assert t.names[13] == "torso_lift_joint" and t.names[14:21] == [f"arm_{i}_joint" for i in range(1, 8)]
    assert t.names[21:23] == ["gripper_finger_joint", "gripper_right_finger_joint"] and t.names[23:] == ["head_1_joint", "head_2_joint"]
    assert (t.njoints, t.nq, t.nv) == (25, 34, 24)
    s = load_flat("talos_ff")
    assert s.names[1] == "root_joint" and s.names[2] == "leg_left_1_joint" and s.names[32:] == ["head_1_joint", "head_2_joint"]


def test_inertia_algebra_matches_dynamic_parameters():
    from figaroh_b200.model import Inertia, SE3
    rng = np.random.default_rng(0)
    A = rng.normal(size=(3, 3))
    Y1 = Inertia(2.0, rng.normal(size=3), A @ A.T)
    B = rng.normal(size=(3, 3))
    Y2 = Inertia(0.7, rng.normal(size=3), B @ B.T)
    # dynamic parameters are additive when both inertias are expressed in the same frame
    np.testing.assert_allclose((Y1 + Y2).toDynamicParameters(), Y1.toDynamicParameters() + Y2.toDynamicParameters(), atol=1e-12)
    Z = Inertia() + Inertia()
    assert Z.mass == 0.0 and np.all(np.isfinite(Z.inertia))
    # moving an inertia with an SE3 keeps mass and rotates the lever
    from figaroh_b200.urdf import rpy_to_matrix
    M = SE3(rpy_to_matrix(0.3, -0.2, 1.1), [0.1, 0.2, 0.3])
    Y3 = M.act_inertia(Y1)
    assert Y3.mass == Y1.mass
    np.testing.assert_allclose(Y3.lever, M.rotation @ Y1.lever + M.translation)


def test_rpy_goes_through_the_quaternion():
    from figaroh_b200.urdf import rpy_to_matrix
    R = rpy_to_matrix(1.570796327, 0, 0)   # the UR10 URDF's rounded pi/2 (examples/ur10/data/robot.urdf:429)
    assert abs(R[1, 1]) < 1e-9 and R[1, 1] != 0.0   # not snapped: the marginal Ia2 column of SURVEY fact 5
    np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-15)
    r, p, y = 0.4, -0.7, 2.0
    cz, sz, cy, sy, cx, sx = np.cos(y), np.sin(y), np.cos(p), np.sin(p), np.cos(r), np.sin(r)
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    np.testing.assert_allclose(rpy_to_matrix(r, p, y), Rz @ Ry @ Rx, atol=1e-15)


def test_regressor_config_from_param():
    from figaroh_b200 import _lib
    from figaroh_b200.engine import RegressorConfig
    c = RegressorConfig.from_param(dict(TX40_PARAM), coupling=True)
    assert c.mode == _lib.FG_MODE_JOINT_TORQUES and c.flags == 15
    ew = make_param(True, False, False, is_joint_torques=False, is_external_wrench=True, force_torque=["Fz", "Mx"])
    c = RegressorConfig.from_param(ew)
    assert c.mode == _lib.FG_MODE_EXTERNAL_WRENCH and c.flags == 1 and c.ft_mask == 0b001100
    assert RegressorConfig.from_param(dict(ew, force_torque=["All"])).ft_mask == 0x3F
    with pytest.raises(ValueError, match="Please enter valid parameters"):      # regressor.py:140
        RegressorConfig.from_param(dict(ew, force_torque=["Fq"]))
    with pytest.raises(UnboundLocalError):                                      # regressor.py:194
        RegressorConfig.from_param(make_param(False, False, False, is_joint_torques=False))
    cs = RegressorConfig(row_sel=[12, 13, 14], col_sel=[0, 5, 9]).cstruct()
    assert cs.n_row_sel == 3 and cs.n_col_sel == 3


def test_expression_strings_match_reference_format():
    from figaroh_b200.engine import expressions
    from oracle import ref_path
    beta = np.array([[1.0, -0.039, 0.0000004], [0.016005, 0.0, -1.0]])
    got = expressions(["m13", "Izz14"], ["m14", "mz15", "Ia2"], beta)
    assert got == ["m13 + 1.0*m14 - 0.039*mz15", "Izz14 + 0.016005*m14 - 1.0*Ia2"]
    assert got == ref_path._expressions(["m13", "Izz14"], ["m14", "mz15", "Ia2"], beta)


def test_standard_parameters_naming_and_order():
    """Robot.get_standard_parameters defines the column order (robot.py:76-155)."""
    from figaroh.tools.robot import Robot
    from figaroh_b200.model import model_from_flat
    from oracle import kernels, ref_path
    r = Robot.__new__(Robot)
    r.model = model_from_flat(load_flat("tx40"))
    std = r.get_standard_parameters(dict(TX40_PARAM))
    keys = list(std.keys())
    assert keys[:14] == ["Ixx1", "Ixy1", "Ixz1", "Iyy1", "Iyz1", "Izz1", "mx1", "my1", "mz1", "m1", "Ia1", "fv1", "fs1", "off1"]
    assert len(keys) == 84 and std["fv3"] == 1.97 and std["off6"] == 0.127 and std["Ia2"] == 0.362
    ref = ref_path.standard_parameters(kernels.OracleModel(golden_path("models", "tx40.json")), dict(TX40_PARAM))
    assert keys == list(ref.keys())
    np.testing.assert_array_equal(np.array(list(std.values()), float), np.array(list(ref.values()), float))
    off = r.get_standard_parameters(make_param(False, False, False))
    assert off["Ia1"] == 0 and off["fs6"] == 0


def test_shard_range_partitions_everything():
    from figaroh_b200.dist import shard_range
    for n, w in [(10, 3), (1_000_000, 8), (7, 8), (0, 2)]:
        parts = [shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1


def test_base_solution_builds_expressions_lazily():
    """The expression strings of qrdecomposition.py:160-183 are built on first access, not in the solve."""
    import numpy as np
    from figaroh_b200 import engine as eng
    s = eng.BaseSolution()
    s.params_r, s.idx_base, s.idx_regroup = ["a", "b", "c"], [0, 2], [1]
    s.beta = np.array([[0.5], [-2e-7]])
    assert s.params_base == ["a + 0.5*b", "c"] and s.rows == 0
