"""CPU: pin the oracle against the fixtures produced by the unmodified reference
(tests/golden/make_golden.py) - including the two golden CSVs the reference ships."""
import json

import numpy as np
import pytest
from scipy import signal

from conftest import TX40_PARAM, golden_path, load_flat, load_oracle_model, make_param, rand_inputs
from oracle import kernels, ref_path


@pytest.mark.parametrize("case,model,param,coupling", [
    ("ur10_on", "ur10", make_param(True, True, True), False),
    ("tx40_cpl", "tx40", dict(TX40_PARAM), True),
    ("tiago_on", "tiago", make_param(True, True, True), False),
    ("talos_ew", "talos_ff", make_param(False, False, False, is_joint_torques=False, is_external_wrench=True,
                                        force_torque=["All"]), False),
    ("talos_ew2", "talos_ff", make_param(True, False, True, is_joint_torques=False, is_external_wrench=True,
                                         force_torque=["Fz", "Mx"]), False),
])
def test_regressor_entries_equal_reference(case, model, param, coupling):
    z = np.load(golden_path("regressor_cases.npz"))
    om = load_oracle_model(model)
    q, v, a = z[case + "_q"], z[case + "_v"], z[case + "_a"]
    for per_sample in (True, False):
        W = ref_path.build_regressor_basic(om, q, v, a, param, per_sample=per_sample)
        if coupling:
            W = ref_path.add_coupling_TX40(W, len(q), v, a)
        assert np.array_equal(W, z[case + "_W"])


@pytest.mark.parametrize("model", ["ur10", "tx40", "tiago", "talos_ff", "human_ff"])
def test_regressor_identity_rnea(model):
    """W_pin(q,v,a) . phi == RNEA(q,v,a) for the model's own inertias."""
    om = load_oracle_model(model)
    fm = load_flat(model)
    q, v, a = rand_inputs(fm, 5, 3, ((-1.5, 1.5), (-2, 2), (-3, 3)))
    phi = om.dynamic_parameters().reshape(-1)
    for i in range(5):
        W = kernels.joint_torque_regressor(om, q[i], v[i], a[i])
        tau = kernels.rnea(om, q[i], v[i], a[i])
        assert np.abs(W @ phi - tau).max() <= 1e-11 * max(1.0, np.abs(tau).max())


def test_structural_index_sets(structural_cases):
    models = {"ur10_off": "ur10", "ur10_on": "ur10", "ur10_fr_off": "ur10", "tx40_on_coupled": "tx40",
              "tiago_on": "tiago", "talos_ew_all": "talos_ff"}
    for name, c in structural_cases.items():
        om, fm = load_oracle_model(models[name]), load_flat(models[name])
        q, v, a = rand_inputs(fm, c["N"], c["seed"], tuple(tuple(x) for x in c["lo_hi"]))
        W = ref_path.build_regressor_basic(om, q, v, a, c["param"], per_sample=False)
        std = ref_path.standard_parameters(om, c["param"])
        if c["coupling"]:
            W = ref_path.add_coupling_TX40(W, c["N"], v, a)
            std.update(Iam6=1.0, fvm6=1.0, fsm6=1.0)
        idx_e, params_r = ref_path.get_index_eliminate(W, std, c["tol_e"])
        assert idx_e == c["idx_e"], name
        _, names, idx_base = ref_path.get_baseParams(ref_path.build_regressor_reduced(W, idx_e), params_r, std)
        assert idx_base == c["idx_base"], name
        assert names == c["params_base"], name


def test_survey_index_vectors(structural_cases):
    """The literal index vectors quoted in SURVEY.md section 8(c)."""
    off = structural_cases["ur10_off"]
    assert off["idx_e"] == [0, 1, 2, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13] + list(range(22, 28)) + list(range(38, 42)) + \
        list(range(52, 56)) + list(range(66, 70)) + list(range(80, 84))
    assert len(off["idx_base"]) == 36
    on = structural_cases["ur10_on"]
    assert on["idx_e"] == [0, 1, 2, 3, 4, 6, 7, 8, 9, 22, 23] and len(on["idx_base"]) == 59
    assert structural_cases["tx40_on_coupled"]["idx_e"] == on["idx_e"]
    assert len(structural_cases["tx40_on_coupled"]["idx_base"]) == 61
    assert len(structural_cases["tiago_on"]["idx_e"]) == 96 and len(structural_cases["tiago_on"]["idx_base"]) == 179
    assert len(structural_cases["talos_ew_all"]["idx_e"]) == 132 and len(structural_cases["talos_ew_all"]["idx_base"]) == 234


def tiago_inputs():
    z = np.load(golden_path("tiago_kat.npz"))
    N = z["q_act"].shape[0]
    fm = load_flat("tiago")
    p = np.tile(z["q0"], (N, 1))
    v = np.zeros((N, fm.nv))
    a = np.zeros((N, fm.nv))
    p[:, z["idxq"]], v[:, z["idxv"]], a[:, z["idxv"]] = z["q_act"], z["dq_act"], z["ddq_act"]
    return p, v, a, z["tau"], z["idxv"]


def tiago_decimate(W_e, tau, idxv, N):
    tau_rf = np.concatenate([signal.decimate(tau[:, i], q=10, zero_phase=True) for i in range(len(idxv))])
    W_rf = np.vstack([np.stack([signal.decimate(W_e[idxv[i] * N:(idxv[i] + 1) * N, j], q=10, zero_phase=True)
                                for j in range(W_e.shape[1])], 1) for i in range(len(idxv))])
    return W_rf, tau_rf


def test_tiago_golden_csv_through_oracle():
    """examples/tiago/.../tiago_bp_19_Oct_2024_2320.csv: 44 expressions and phi_b reproduced exactly."""
    meta = json.load(open(golden_path("tiago_kat.json")))
    om = load_oracle_model("tiago")
    p, v, a, tau, idxv = tiago_inputs()
    N = len(p)
    param = make_param(False, False, False)
    W = ref_path.build_regressor_basic(om, p, v, a, param, per_sample=False)
    std = ref_path.standard_parameters(om, param)
    idx_e, params_r = ref_path.get_index_eliminate(W, std, 0.001)
    assert idx_e == meta["idx_e"]
    W_rf, tau_rf = tiago_decimate(ref_path.build_regressor_reduced(W, idx_e), tau, idxv, N)
    W_b, bp, base, phi_b, phi_std = ref_path.double_QR(tau_rf, W_rf, params_r, std)
    gold = meta["golden_csv"]
    assert [g[0] for g in gold] == base
    assert max(abs(g[1] - x) for g, x in zip(gold, phi_b)) == 0.0
    sd = ref_path.relative_stdev(W_b, phi_b, tau_rf)
    assert max(abs(g[2] - 100 * x) for g, x in zip(gold, sd)) < 1e-6


def test_tx40_golden_csv_through_oracle():
    """examples/staubli_TX40/results/TX40_bp_5.csv: all 60 expression strings reproduced."""
    meta = json.load(open(golden_path("tx40_kat.json")))
    z = np.load(golden_path("tx40_kat.npz"))
    std = {n: 0.0 for n in meta["params_std_names"]}
    W_e, params_r = ref_path.eliminate_non_dynaffect(z["W_"], std, 0.001)
    assert params_r == meta["params_r"]
    W_b, bp, base, phi_b = ref_path.double_QR(z["tau_"], W_e, params_r)
    assert base == [g[0] for g in meta["golden_csv"]]
    assert np.array_equal(phi_b, np.array(meta["phi_b"]))
    # values in the shipped CSV were produced with another scipy (filter coefficients): soft check
    assert max(abs(g[1] - x) for g, x in zip(meta["golden_csv"], meta["phi_ols"])) < 1e-3
