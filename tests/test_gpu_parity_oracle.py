"""GPU parity of the FUSED path (kernels (1)+(2)+(3) through the C ABI) against the CPU oracle
(`oracle/ref_path.py`, which equals the unmodified reference bit for bit, tests/test_oracle_vs_reference.py).

Bars (BASELINE.json north_star): eliminated / base index sets and expression strings identical; phi_b within
1e-9 relative where that is well posed.  The benched workload (UR10, friction + actuator inertia + offsets on
the reference's own URDF) is NOT well posed component-wise: the URDF rounds pi/2 to 1.570796327, which makes
`Ia2` a base column with |R_ii| ~ 1e-9 sqrt(N) and cond(R1) ~ 1e11 - the reference's own two
formulas (inv(R1) Q1^T tau vs lstsq) disagree at 1e-6 there.  Decision, written beside every number:
  (a) on the pi/2-snapped models (tests/golden/models/*_snapped.json; Ia2 exactly dependent, cond(R1) ~ 2e2)
      the plain component-wise 1e-9 bar is asserted, with all flags on;
  (b) on the raw URDF the index sets must be identical and phi_b must agree to 1e-9 in the norm in which a
      least-squares solution IS well posed - the predicted torques, ||R1 (phi_gpu - phi_ref)|| / ||tau|| -
      with cond(R1) asserted to be what makes the component-wise bar meaningless.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, TX40_PARAM, load_flat, load_oracle_model, make_param, rand_inputs

pytestmark = pytest.mark.gpu

REL = 1e-9
FLAGS_ON = dict(fv=TX40_PARAM["fv"], fs=TX40_PARAM["fs"], Ia=TX40_PARAM["Ia"], off=TX40_PARAM["off"])


def _eng():
    from figaroh_b200 import engine
    return engine


def _oracle_path(om, q, v, a, param, tol_e, coupling=False, rows=None, noise_seed=0):
    """The reference's script sequence on the CPU oracle: W -> tau -> idx_e -> W_e -> double_QR."""
    from oracle import ref_path
    N = len(q)
    std = ref_path.standard_parameters(om, param)
    W = ref_path.build_regressor_basic(om, q, v, a, param, per_sample=False)
    if coupling:
        W = ref_path.add_coupling_TX40(W, N, v, a)
        std.update(Iam6=param["Iam6"], fvm6=param["fvm6"], fsm6=param["fsm6"])
    if rows is not None:  # examples/tiago/identification.py:157-166: only the active joints' row blocks
        W = np.concatenate([W[r * N:(r + 1) * N] for r in rows])
    phi = np.array(list(std.values()), dtype=float)[:W.shape[1]]
    rng = np.random.default_rng(noise_seed)
    tau = W @ phi
    tau = tau + 0.01 * np.sqrt(np.mean(tau ** 2)) * rng.normal(size=tau.shape)
    idx_e, params_r = ref_path.get_index_eliminate(W, std, tol_e)
    W_e = ref_path.build_regressor_reduced(W, idx_e)
    ref = ref_path.double_QR(tau, W_e, params_r, std)
    return dict(std=std, names=list(std.keys()), W=W, tau=tau, idx_e=idx_e, params_r=params_r, W_e=W_e, ref=ref)


def _phi_unrounded(o, idx_base):
    """inv(R1) Q1^T tau of qrdecomposition.py:140-147 before its round(., 6)."""
    Wb = o["W_e"][:, idx_base]
    Q1, R1 = np.linalg.qr(Wb)
    return np.linalg.solve(R1, Q1.T @ o["tau"]), R1


@pytest.mark.parametrize("model,coupling,N,seed,lo_hi", [
    ("ur10_snapped", False, 3000, 12, ((-6, 6), (-6, 6), (-6, 6))),
    ("tx40_snapped", True, 2500, 14, ((-6, 6), (-10, 10), (-30, 30))),
])
def test_fused_flags_on_snapped_models_vs_oracle(model, coupling, N, seed, lo_hi):
    """(a): all flags ON, pi/2-snapped model: index sets, expression strings, phi_b (1e-9) and phi_std vs the oracle."""
    eng = _eng()
    fm, om = load_flat(model), load_oracle_model(model)
    q, v, a = rand_inputs(fm, N, seed, lo_hi)
    param = dict(TX40_PARAM) if coupling else make_param(True, True, True, **FLAGS_ON)
    o = _oracle_path(om, q, v, a, param, 1e-6, coupling=coupling)
    dm = eng.DeviceModel(fm)
    cfg = eng.RegressorConfig.from_param(param, coupling=coupling)
    sol = eng.identify(dm, cfg, q, v, a, o["tau"], o["names"], tol_e=1e-6, params_std=o["std"])
    assert sol.idx_e == o["idx_e"] and sol.params_r == o["params_r"]
    assert sol.params_base == o["ref"][2]                      # strings incl. the beta coefficients
    phi_ref, R1 = _phi_unrounded(o, sol.idx_base)
    cond = np.linalg.cond(R1)
    assert cond < 1e4, cond
    rel = np.abs(sol.phi_b_raw - phi_ref).max() / np.abs(phi_ref).max()
    assert rel < REL, (rel, cond)
    np.testing.assert_allclose(sol.phi_b_raw, phi_ref, rtol=1e-9 * cond, atol=1e-9 * np.abs(phi_ref).max())
    assert np.abs(sol.phi_b - o["ref"][3]).max() <= 1.01e-6      # both rounded to 6 decimals
    np.testing.assert_allclose(sol.phi_std, o["ref"][4], atol=1.01e-5)


@pytest.mark.parametrize("N,seed", [(1000, 12), (4000, 31)])
def test_fused_ur10_flags_on_raw_urdf_vs_oracle(N, seed):
    """(b): the BENCHED configuration (raw UR10 URDF, all flags on) against the oracle's double_QR."""
    eng = _eng()
    fm, om = load_flat("ur10"), load_oracle_model("ur10")
    q, v, a = rand_inputs(fm, N, seed)
    param = make_param(True, True, True, **FLAGS_ON)
    o = _oracle_path(om, q, v, a, param, 1e-6)
    dm = eng.DeviceModel(fm)
    cfg = eng.RegressorConfig.from_param(param)
    sol = eng.identify(dm, cfg, q, v, a, o["tau"], o["names"], tol_e=1e-6, params_std=o["std"])
    assert sol.idx_e == o["idx_e"] == [0, 1, 2, 3, 4, 6, 7, 8, 9, 22, 23]
    assert sol.params_r == o["params_r"]
    ref_base = [b.split(" ")[0] for b in o["ref"][2]]
    assert [sol.params_r[i] for i in sol.idx_base] == ref_base and len(sol.idx_base) == 59
    phi_ref, R1 = _phi_unrounded(o, sol.idx_base)
    cond = np.linalg.cond(R1)
    assert cond > 1e9, cond   # the reason the component-wise bar does not apply here (Ia2 is a marginal base column)
    # well-posed norm: the predicted torques W_b phi = Q1 R1 phi
    err = np.linalg.norm(R1 @ (sol.phi_b_raw - phi_ref)) / np.linalg.norm(o["tau"])
    assert err < REL, (err, cond)
    # components outside the near-dependent pair agree to a conditioning-scaled bound; the reference's own
    # inv(R1) Q1^T tau and lstsq differ by more than this on the same data
    phi_ls = np.linalg.lstsq(o["W_e"][:, sol.idx_base], o["tau"], rcond=None)[0]
    own = np.abs(phi_ref - phi_ls).max() / np.abs(phi_ls).max()
    rel = np.abs(sol.phi_b_raw - phi_ref).max() / np.abs(phi_ref).max()
    assert rel < max(100 * own, 2e-16 * cond), (rel, own, cond)
    # the margin of the marginal column is reproduced (|R_ii| of Ia2 ~ 1e-9 sqrt(N), tol_qr = 1e-8)
    d = sol.diag[:len(sol.keep)]
    dref = np.abs(np.diag(np.linalg.qr(o["W_e"], mode="r")))
    assert abs(d[d > 1e-8].min() / dref[dref > 1e-8].min() - 1) < 0.10
    assert d[d <= 1e-8].max() < 1e-9


def test_fused_tiago_active_joints_vs_oracle():
    """BASELINE configs[2] shape (TIAGo torso + arm: 8 active row blocks of the 24-DoF tree, flags on) against
    the oracle: regressor over the full tree, active row blocks, get_index_eliminate(1e-3), double_QR."""
    eng = _eng()
    fm, om = load_flat("tiago"), load_oracle_model("tiago")
    N = 900
    q, v, a = rand_inputs(fm, N, 9, ((-1.5, 1.5), (-2, 2), (-4, 4)))
    rows = [int(fm.idx_v[i]) for i in range(13, 21)]  # torso_lift + arm_1..7
    active = set(rows)
    for j in range(fm.nv):  # the other joints rest, as in the recorded trajectories
        if j not in active:
            v[:, j] = 0.0
            a[:, j] = 0.0
    for i in range(1, fm.njoints):
        if int(fm.idx_v[i]) not in active:
            iq = fm.idx_q[i]
            if fm.jtype[i] == 2:
                q[:, iq], q[:, iq + 1] = 1.0, 0.0
            else:
                q[:, iq] = 0.1
    param = make_param(True, True, True, fv=[0.5] * 24, fs=[0.3] * 24, Ia=[0.1] * 24, off=[0.05] * 24)
    o = _oracle_path(om, q, v, a, param, 1e-3, rows=rows)
    dm = eng.DeviceModel(fm)
    cfg = eng.RegressorConfig.from_param(param, row_sel=rows)
    sol = eng.identify(dm, cfg, q, v, a, o["tau"], o["names"], tol_e=1e-3, params_std=o["std"])
    assert sol.idx_e == o["idx_e"] and sol.params_r == o["params_r"]
    assert sol.params_base == o["ref"][2]
    phi_ref, R1 = _phi_unrounded(o, sol.idx_base)
    cond = np.linalg.cond(R1)
    rel = np.abs(sol.phi_b_raw - phi_ref).max() / np.abs(phi_ref).max()
    assert rel < REL * max(1.0, cond / 1e3), (rel, cond)
    assert np.abs(sol.phi_b - o["ref"][3]).max() <= 1.01e-6


@pytest.mark.parametrize("ft", [["All"], ["Fz", "Mx", "My"]])
def test_fused_talos_external_wrench_vs_oracle(ft):
    """BASELINE configs[3] shape (Talos + free-flyer, external-wrench rows) against the oracle."""
    eng = _eng()
    fm, om = load_flat("talos_ff"), load_oracle_model("talos_ff")
    N = 500
    q, v, a = rand_inputs(fm, N, 13, ((-1, 1), (-1, 1), (-2, 2)))
    param = dict(is_joint_torques=False, is_external_wrench=True, force_torque=ft, has_friction=False,
                 has_actuator_inertia=False, has_joint_offset=False, fv=None, fs=None, Ia=None, off=None)
    o = _oracle_path(om, q, v, a, param, 1e-6)
    dm = eng.DeviceModel(fm)
    cfg = eng.RegressorConfig.from_param(param)
    sol = eng.identify(dm, cfg, q, v, a, o["tau"], o["names"], tol_e=1e-6, params_std=o["std"])
    assert sol.idx_e == o["idx_e"] and sol.params_r == o["params_r"]
    if ft == ["All"]:
        assert len(sol.idx_e) == 132 and len(sol.idx_base) == 234
    ref_base = [b.split(" ")[0] for b in o["ref"][2]]
    assert [sol.params_r[i] for i in sol.idx_base] == ref_base
    assert sol.params_base == o["ref"][2]
    phi_ref, R1 = _phi_unrounded(o, sol.idx_base)
    cond = np.linalg.cond(R1)
    rel = np.abs(sol.phi_b_raw - phi_ref).max() / np.abs(phi_ref).max()
    assert rel < REL * max(1.0, cond / 1e3), (rel, cond)
    assert np.abs(sol.phi_b - o["ref"][3]).max() <= 1.01e-6


@pytest.mark.parametrize("name", ["tiago_on", "talos_ew_all"])
def test_fused_structural_index_sets_wide_models(name, structural_cases):
    """The reference-generated structural vectors of the 24-row-block TIAGo tree (336 columns) and of Talos in
    external-wrench mode (462 columns) through fg_state_accumulate_samples."""
    eng = _eng()
    c = structural_cases[name]
    fm = load_flat("tiago" if name.startswith("tiago") else "talos_ff")
    q, v, a = rand_inputs(fm, c["N"], c["seed"], tuple(tuple(x) for x in c["lo_hi"]))
    dm = eng.DeviceModel(fm)
    cfg = eng.RegressorConfig.from_param(c["param"], coupling=c["coupling"])
    rps, cols = dm.shape(cfg)
    names = [f"c{i}" for i in range(cols)]
    sol = eng.identify(dm, cfg, q, v, a, None, names, tol_e=c["tol_e"])
    assert sol.rows == c["N"] * rps
    assert sol.idx_e == c["idx_e"]
    assert sol.idx_base == c["idx_base"]
    d = sol.diag[:len(sol.keep)]
    assert abs(d[d > 1e-8].min() / c["min_base_diag"] - 1) < 0.10


def test_nccl_ranks_reproduce_single_gpu_solution():
    """Real NCCL: a 2-rank torchrun run of the sharded path (fused kernel per rank, all-gather of the packed
    factors, merge on every rank) must reproduce the single-GPU solution and the oracle's index sets."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    script = os.path.join(ROOT, "tests", "nccl_parity_worker.py")
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), script]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "NCCL_PARITY_OK" in r.stdout, r.stdout[-3000:]
