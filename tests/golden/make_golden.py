"""Generate the committed fixtures under tests/golden/ from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  The reference's Python
files are imported by path through oracle/pinshim.py, whose fake ``pinocchio``
serves the restated C arithmetic (oracle/fg_oracle.c) and the product's URDF loader.
The end-to-end cases are additionally compared with the golden CSVs the reference
ships, so a fixture is only written if the reference's own numbers are reproduced:

  * TIAGo:  examples/tiago/data/identification/dynamic/tiago_bp_19_Oct_2024_2320.csv
            (pipeline restated from examples/tiago/identification.py:41-187,260-348,385-415)
  * TX40:   examples/staubli_TX40/results/TX40_bp_5.csv
            (pipeline restated from examples/staubli_TX40/identification.py:43-240)

Usage:  python tests/golden/make_golden.py
"""
import csv
import json
import os
import sys

import numpy as np
import pandas as pd
import yaml
from scipy import signal

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "figaroh-plus_b200"))

from oracle import pinshim  # noqa: E402
from figaroh_b200.model import JointModelFreeFlyer, flatten_model  # noqa: E402
from figaroh_b200.urdf import build_model_from_urdf  # noqa: E402

REF = "/root/reference/"
reg = pinshim.load_reference("regressor")
qrd = pinshim.load_reference("qrdecomposition")
rob = pinshim.load_reference("robot")
idt = pinshim.load_reference("identification_tools")

URDFS = {
    "ur10": ("examples/ur10/data/robot.urdf", False),
    "tx40": ("models/staubli_tx40_description/urdf/tx40_mdh_modified.urdf", False),
    "tiago": ("examples/tiago/urdf/tiago_48_schunk.urdf", False),
    "talos_ff": ("examples/talos/data/talos_full_v2.urdf", True),
    "human_ff": ("models/human_description/urdf/human.urdf", True),
}

TX40_PARAM = dict(
    is_joint_torques=True, is_external_wrench=False, force_torque=None,
    has_friction=True, has_actuator_inertia=True, has_joint_offset=True, has_coupled_wrist=True,
    fv=[8.05e0, 5.53e0, 1.97e0, 1.11e0, 1.86e0, 6.5e-1], fs=[7.14e0, 8.26e0, 6.34e0, 2.48e0, 3.03e0, 2.82e-1],
    Ia=[3.62e-1, 3.62e-1, 9.88e-2, 3.13e-2, 4.68e-2, 1.05e-2], off=[3.92e-1, 1.37, 3.26e-1, -1.02e-1, -2.88e-2, 1.27e-1],
    Iam6=9.64e-3, fvm6=6.16e-1, fsm6=1.95e0)


def flags(fr, ia, jo, **kw):
    d = dict(is_joint_torques=True, is_external_wrench=False, force_torque=None,
             has_friction=fr, has_actuator_inertia=ia, has_joint_offset=jo,
             fv=None, fs=None, Ia=None, off=None)
    d.update(kw)
    return d


def write_models():
    os.makedirs(os.path.join(HERE, "models"), exist_ok=True)
    for name, (path, ff) in URDFS.items():
        for snap in (False, True):
            m = build_model_from_urdf(REF + path, JointModelFreeFlyer() if ff else None, snap_right_angles=snap)
            fm = flatten_model(m)
            out = os.path.join(HERE, "models", name + ("_snapped" if snap else "") + ".json")
            with open(out, "w") as f:
                f.write(fm.to_json())
            print("model", name, "snap" if snap else "raw", fm.njoints, fm.nq, fm.nv)


def rand_inputs(model, N, seed, lo_hi=((-6, 6), (-6, 6), (-6, 6)), free_flyer=False):
    """Seeded inputs, reproducible in the tests from (seed, N) alone."""
    rng = np.random.default_rng(seed)
    q = rng.uniform(*lo_hi[0], size=(N, model.nq))
    v = rng.uniform(*lo_hi[1], size=(N, model.nv))
    a = rng.uniform(*lo_hi[2], size=(N, model.nv))
    fm = flatten_model(model)
    for i in range(1, fm.njoints):
        iq = fm.idx_q[i]
        if fm.jtype[i] == 2:  # continuous: (cos, sin)
            th = rng.uniform(-np.pi, np.pi, N)
            q[:, iq], q[:, iq + 1] = np.cos(th), np.sin(th)
        elif fm.jtype[i] == 4:
            quat = rng.normal(size=(N, 4))
            q[:, iq + 3:iq + 7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    return q, v, a


def structural_case(name, robot, param, N, seed, tol_e, coupling=False, lo_hi=((-6, 6), (-6, 6), (-6, 6))):
    q, v, a = rand_inputs(robot.model, N, seed, lo_hi)
    W = reg.build_regressor_basic(robot, q, v, a, param)
    std = robot.get_standard_parameters(param)
    if coupling:
        W = reg.add_coupling_TX40(W, robot.model, robot.data, N, robot.model.nq, robot.model.nv, robot.model.njoints, q, v, a)
        std["Iam6"], std["fvm6"], std["fsm6"] = param["Iam6"], param["fvm6"], param["fsm6"]
    idx_e, params_r = reg.get_index_eliminate(W, std, tol_e)
    W_e = reg.build_regressor_reduced(W, idx_e)
    _, params_base, idx_base = qrd.get_baseParams(W_e, params_r, std)
    R = np.linalg.qr(W_e, mode="r")
    d = np.abs(np.diag(R))
    big = d[d > 1e-8].min()
    small = d[d <= 1e-8].max() if (d <= 1e-8).any() else 0.0
    print(f"struct {name}: N={N} p={W.shape[1]} elim={len(idx_e)} p_e={W_e.shape[1]} rank={len(idx_base)} "
          f"min base |Rii|={big:.3e} max dep |Rii|={small:.3e}")
    return dict(name=name, N=N, seed=seed, tol_e=tol_e, lo_hi=lo_hi, coupling=coupling,
                param={k: v_ for k, v_ in param.items()}, idx_e=[int(i) for i in idx_e],
                idx_base=[int(i) for i in idx_base], params_base=params_base,
                min_base_diag=float(big), max_dep_diag=float(small))


def regressor_case(name, robot, param, N, seed, coupling=False):
    q, v, a = rand_inputs(robot.model, N, seed, ((-3, 3), (-3, 3), (-6, 6)))
    v[0, :] = 0.0  # sign(0) = 0 edge case (regressor.py:57)
    W = reg.build_regressor_basic(robot, q, v, a, param)
    if coupling:
        W = reg.add_coupling_TX40(W, robot.model, robot.data, N, robot.model.nq, robot.model.nv, robot.model.njoints, q, v, a)
    return {f"{name}_q": q, f"{name}_v": v, f"{name}_a": a, f"{name}_W": W}


def tiago_kat():
    D = REF + "examples/tiago/"
    robot = rob.Robot(D + "urdf/tiago_48_schunk.urdf", package_dirs=None)
    model = robot.model
    ps = flags(False, False, False)  # the golden was produced with all three flags off
    active = ["torso_lift_joint"] + [f"arm_{i}_joint" for i in range(1, 8)]
    jid = [model.getJointId(n) for n in active]
    idxq = [model.joints[j].idx_q for j in jid]
    idxv = [model.joints[j].idx_v for j in jid]
    pos = pd.read_csv(D + "data/identification/dynamic/tiago_position.csv")
    vel = pd.read_csv(D + "data/identification/dynamic/tiago_velocity.csv")
    eff = pd.read_csv(D + "data/identification/dynamic/tiago_effort.csv")

    def cols(df):
        return [c for jn in active for c in df.columns if jn in c]
    t = pos.iloc[:, [0]].to_numpy()[921:6791]
    q, dq, tau = (df[cols(df)].to_numpy()[921:6791] for df in (pos, vel, eff))
    b1, b2 = signal.butter(4, 2 / 50, "low")
    pad = 3 * (max(len(b1), len(b2)) - 1)
    qf = np.stack([signal.filtfilt(b1, b2, signal.medfilt(q[:, j], 5), padtype="odd", padlen=pad) for j in range(8)], 1)
    dqf = np.stack([signal.filtfilt(b1, b2, signal.medfilt(dq[:, j], 5), padtype="odd", padlen=pad) for j in range(8)], 1)
    ddqf = np.stack([np.gradient(dqf[:, j]) / np.gradient(t[:, 0]) for j in range(8)], 1)
    N = qf.shape[0]
    p = np.tile(robot.q0, (N, 1))
    v = np.tile(robot.v0, (N, 1))
    a = np.tile(robot.v0, (N, 1))
    p[:, idxq], v[:, idxv], a[:, idxv] = qf, dqf, ddqf
    rr = [1, 100, 100, 100, 100, 336, 336, 336]
    km = [1, 0.136, 0.136, -0.087, -0.087, -0.0613, -0.0613, -0.0613]
    pin = pinshim.install()
    pin.computeSubtreeMasses(model, robot.data)
    tau = tau.copy()
    for i, jn in enumerate(active):
        tau[:, i] = rr[i] * km[i] * tau[:, i] + (9.81 * robot.data.mass[jid[i]] if i == 0 else 0.0)
    W = reg.build_regressor_basic(robot, p, v, a, ps)
    std = robot.get_standard_parameters(ps)
    idx_e, par_r = reg.get_index_eliminate(W, std, tol_e=0.001)
    W_e = reg.build_regressor_reduced(W, idx_e)
    tau_rf = np.concatenate([signal.decimate(tau[:, i], q=10, zero_phase=True) for i in range(8)])
    W_rf = np.vstack([np.stack([signal.decimate(W_e[idxv[i] * N:(idxv[i] + 1) * N, j], q=10, zero_phase=True)
                                for j in range(W_e.shape[1])], 1) for i in range(8)])
    W_b, bp, base, phi_b, phi_std = qrd.double_QR(tau_rf, W_rf, par_r, std)
    sd = idt.relative_stdev(W_b, phi_b, tau_rf)
    gold = list(csv.reader(open(D + "data/identification/dynamic/tiago_bp_19_Oct_2024_2320.csv")))
    assert [g[0] for g in gold] == base, "TIAGo expression strings differ from the reference golden"
    assert max(abs(float(g[1]) - x) for g, x in zip(gold, phi_b)) == 0.0
    print(f"tiago KAT: {len(base)}/44 strings, phi_b exact, W{W.shape} -> W_rf{W_rf.shape}")
    np.savez_compressed(os.path.join(HERE, "tiago_kat.npz"), q_act=qf, dq_act=dqf, ddq_act=ddqf, tau=tau,
                        q0=robot.q0, idxq=np.array(idxq), idxv=np.array(idxv))
    with open(os.path.join(HERE, "tiago_kat.json"), "w") as f:
        json.dump(dict(idx_e=[int(i) for i in idx_e], params_base=base, phi_b=[float(x) for x in phi_b],
                       phi_std=[float(x) for x in phi_std], rel_std=[float(x) for x in sd],
                       golden_csv=[[g[0], float(g[1]), float(g[2])] for g in gold],
                       shapes=dict(W=W.shape, W_e=W_e.shape, W_rf=W_rf.shape)), f)


def tx40_kat():
    D = REF + "examples/staubli_TX40/"
    robot = rob.Robot(REF + "models/staubli_tx40_description/urdf/tx40_mdh_modified.urdf", package_dirs=None)
    model = robot.model
    cfg = yaml.safe_load(open(D + "config/T40X_config.yaml"))["identification"]
    ps = idt.get_param_from_yaml(robot, cfg)
    std = robot.get_standard_parameters(ps)
    std["Iam6"], std["fvm6"], std["fsm6"] = ps["Iam6"], ps["fvm6"], ps["fsm6"]
    curr = pd.read_csv(D + "data/curr_data.csv").to_numpy()
    posd = pd.read_csv(D + "data/pos_read_data.csv").to_numpy()
    Nr = ps["N"]
    red_q = np.diag(Nr[:6]).astype(float)
    red_q[5, 4] = Nr[5]
    q_nofilt = np.dot(np.linalg.inv(red_q), posd.T).T
    nbutter = 4
    nbord = 5 * nbutter
    q = np.column_stack([idt.low_pass_filter_data(q_nofilt[:, ii], ps, nbutter) for ii in range(6)])
    q[:, 1] += -np.pi / 2
    q[:, 2] += np.pi / 2
    q, dq, ddq = idt.calculate_first_second_order_differentiation(model, q, ps)
    N = q.shape[0]
    W = reg.build_regressor_basic(robot, q, dq, ddq, ps)
    W = reg.add_coupling_TX40(W, model, robot.data, N, 6, 6, 7, q, dq, ddq)
    red_tau = np.diag(Nr[:6]).astype(float)
    red_tau[4, 5] = Nr[5]
    tau_T = np.dot(red_tau, curr.T)
    tau_T = np.delete(tau_T, np.s_[0:nbord], axis=1)
    tau_T = np.delete(tau_T, np.s_[(tau_T.shape[1] - nbord):tau_T.shape[1]], axis=1)
    tau = np.asarray(tau_T).ravel()
    nj = tau.shape[0] // 6
    tau_l, W_l = [], []
    for i in range(6):
        tt = tau[i * nj:(i + 1) * nj]
        for _ in range(2):
            tt = signal.decimate(tt, q=10, zero_phase=True)
        Wj = np.zeros((tt.shape[0], W.shape[1]))
        for j in range(W.shape[1]):
            c = W[i * nj:(i + 1) * nj, j]
            for _ in range(2):
                c = signal.decimate(c, q=10, zero_phase=True)
            Wj[:, j] = c
        keep = np.abs(Wj[:, i * 14 + 11]) >= ps["dq_lim_def"][i]
        tau_l.append(tt[keep])
        W_l.append(Wj[keep])
    tau_ = np.concatenate(tau_l)
    W_ = np.vstack(W_l)
    W_e, params_r = reg.eliminate_non_dynaffect(W_, std, 0.001)
    W_b, bp, base, phi_b = qrd.double_QR(tau_, W_e, params_r)
    phi_ols = np.around(np.linalg.lstsq(W_b, tau_, rcond=None)[0], 6)
    gold = list(csv.reader(open(D + "results/TX40_bp_5.csv")))
    gold = [g for g in gold if g]
    # first row may be a header
    if not gold[0][1].replace(".", "").replace("-", "").replace("e", "").isdigit():
        try:
            float(gold[0][1])
        except ValueError:
            gold = gold[1:]
    nmatch = sum(g[0] == b for g, b in zip(gold, base))
    dv = max(abs(float(g[1]) - x) for g, x in zip(gold, phi_ols))
    print(f"tx40 KAT: N={N} W{W.shape} -> W_{W_.shape} -> W_e{W_e.shape}; {nmatch}/{len(gold)} strings; max|phi_ols-golden|={dv:.2e}")
    assert nmatch == len(gold) == len(base)
    np.savez_compressed(os.path.join(HERE, "tx40_kat.npz"), W_=W_, tau_=tau_)
    with open(os.path.join(HERE, "tx40_kat.json"), "w") as f:
        json.dump(dict(params_std_names=list(std.keys()), params_r=params_r, params_base=base,
                       phi_b=[float(x) for x in phi_b], phi_ols=[float(x) for x in phi_ols],
                       golden_csv=[[g[0], float(g[1])] for g in gold], N=int(N)), f)


def main():
    write_models()
    cases, arrays = [], {}
    ur10 = rob.Robot(REF + URDFS["ur10"][0], package_dirs=None)
    tx40 = rob.Robot(REF + URDFS["tx40"][0], package_dirs=None)
    tiago = rob.Robot(REF + URDFS["tiago"][0], package_dirs=None)
    talos = rob.Robot(REF + URDFS["talos_ff"][0], package_dirs=None, isFext=True)
    on = flags(True, True, True, fv=TX40_PARAM["fv"], fs=TX40_PARAM["fs"], Ia=TX40_PARAM["Ia"], off=TX40_PARAM["off"])
    cases.append(structural_case("ur10_off", ur10, flags(False, False, False), 1000, 11, 1e-6))
    cases.append(structural_case("ur10_on", ur10, on, 1000, 12, 1e-6))
    cases.append(structural_case("ur10_fr_off", ur10, flags(True, False, True, fv=TX40_PARAM["fv"], fs=TX40_PARAM["fs"],
                                                             off=TX40_PARAM["off"]), 1000, 13, 1e-6))
    cases.append(structural_case("tx40_on_coupled", tx40, dict(TX40_PARAM), 2000, 14, 1e-6, coupling=True,
                                 lo_hi=((-6, 6), (-10, 10), (-30, 30))))
    cases.append(structural_case("tiago_on", tiago, flags(True, True, True), 600, 15, 1e-6,
                                 lo_hi=((-1, 1), (-2, 2), (-4, 4))))
    ew = flags(False, False, False, is_joint_torques=False, is_external_wrench=True, force_torque=["All"])
    cases.append(structural_case("talos_ew_all", talos, ew, 300, 16, 1e-6, lo_hi=((-1, 1), (-1, 1), (-2, 2))))
    ew2 = dict(ew, force_torque=["Fz", "Mx"], has_friction=True, has_joint_offset=True)
    arrays.update(regressor_case("ur10_on", ur10, on, 8, 21))
    arrays.update(regressor_case("tx40_cpl", tx40, dict(TX40_PARAM), 8, 22, coupling=True))
    arrays.update(regressor_case("tiago_on", tiago, flags(True, True, True), 5, 23))
    arrays.update(regressor_case("talos_ew", talos, ew, 4, 24))
    arrays.update(regressor_case("talos_ew2", talos, ew2, 3, 25))
    with open(os.path.join(HERE, "structural.json"), "w") as f:
        json.dump(cases, f)
    np.savez_compressed(os.path.join(HERE, "regressor_cases.npz"), **arrays)
    tiago_kat()
    tx40_kat()


if __name__ == "__main__":
    main()
