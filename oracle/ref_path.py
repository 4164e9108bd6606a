"""numpy restatement of the reference's Python-level hot path.

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.  /root/reference cannot travel
to the GPU box, so the host-level algorithm of
    /root/reference/src/figaroh/tools/regressor.py        (assembly, elimination)
    /root/reference/src/figaroh/tools/qrdecomposition.py  (QR base parameters, OLS)
    /root/reference/src/figaroh/tools/robot.py:76-155     (parameter naming)
    /root/reference/src/figaroh/identification/identification_tools.py:204-234 (relative_stdev)
is restated here on top of oracle/fg_oracle.c.  In this container
tests/test_oracle_vs_reference.py checks every function against the unmodified
reference files (imported by path through oracle/pinshim.py) on the same inputs.
"""
import numpy as np

from . import kernels

TOL_QR = 1e-8
_INERTIAL = ("Ixx", "Ixy", "Ixz", "Iyy", "Iyz", "Izz", "mx", "my", "mz", "m")
# position of pinocchio's [m, mx, my, mz, Ixx, Ixy, Iyy, Ixz, Iyz, Izz] inside FIGAROH's 10-block
_PIN_TO_FIG = np.array([9, 6, 7, 8, 0, 1, 3, 2, 4, 5])


def standard_parameters(om, param):
    """robot.py:76-155 - ordered name->value dict that defines the column order."""
    out = {}
    dyn = om.dynamic_parameters()
    for i in range(1, om.njoints):
        blk = np.empty(10)
        blk[_PIN_TO_FIG] = dyn[i - 1]
        for nm, val in zip(_INERTIAL, blk):
            out[f"{nm}{i}"] = val

        def pick(flag, key):
            if not param[flag]:
                return 0
            try:
                return param[key][i - 1]
            except Exception:
                return 0
        out[f"Ia{i}"] = pick("has_actuator_inertia", "Ia")
        if param["has_friction"]:
            try:
                fv, fs = param["fv"][i - 1], param["fs"][i - 1]
            except Exception:
                fv, fs = 0, 0
        else:
            fv, fs = 0, 0
        out[f"fv{i}"] = fv
        out[f"fs{i}"] = fs
        out[f"off{i}"] = pick("has_joint_offset", "off")
    return out


_FT_ROW = {"Fx": 0, "Fy": 1, "Fz": 2, "Mx": 3, "My": 4, "Mz": 5}


def build_regressor_basic(om, q, v, a, param, per_sample=True):
    """regressor.py:20-194.  per_sample=True keeps the reference's one-call-per-sample
    loop (used for the timed CPU baseline); False does the whole batch in one C call."""
    q = np.asarray(q, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    a = np.asarray(a, dtype=np.float64)
    N = len(q)
    nb, nv = om.njoints - 1, om.nv
    flags = (1 if param["has_friction"] else 0) | (2 if param["has_actuator_inertia"] else 0) | \
            (4 if param["has_joint_offset"] else 0)
    if param["is_joint_torques"]:
        mode, mask = 0, 0x3F
    elif param["is_external_wrench"]:
        mode, mask = 1, 0
        for ft in param["force_torque"]:
            if ft == "All":
                mask |= 0x3F
            elif ft in _FT_ROW:
                mask |= 1 << _FT_ROW[ft]
            else:
                raise ValueError("Please enter valid parameters")
    else:
        raise UnboundLocalError("cannot access local variable 'W_mod' where it is not associated with a value")
    if not per_sample:
        return kernels.build_regressor_basic_c(om, q, v, a, mode, flags, mask)
    if mode == 0 and nv != nb:
        raise ValueError(f"could not broadcast input array from shape ({10 * nb},) into shape ({10 * nv},)")
    nblk = nv if mode == 0 else 6
    W = np.zeros((nblk * N, 14 * nb))
    W3 = W.reshape(nblk, N, 14 * nb)              # W3[j, i, :] is row j*N+i of the reference layout
    body_cols = (14 * np.arange(nb)[:, None] + _PIN_TO_FIG[None, :]).reshape(-1)
    has_mass = om.mass[1:] != 0.0
    ew_cols = (14 * np.nonzero(has_mass)[0][:, None] + _PIN_TO_FIG[None, :]).reshape(-1)
    ew_src = (10 * np.nonzero(has_mass)[0][:, None] + np.arange(10)[None, :]).reshape(-1)
    sel = np.array([j for j in range(6) if (mask >> j) & 1], dtype=np.int64)
    jj = np.arange(nv if mode == 0 else nb)
    c_fv, c_fs, c_ia, c_off = 14 * jj + 11, 14 * jj + 12, 14 * jj + 10, 14 * jj + 13
    Wt = np.empty((nv, 10 * nb))
    call = kernels.lib().orc_joint_torque_regressor
    mp, wp = om.ptr, Wt.ctypes.data
    qp, vp, ap = q.ctypes.data, v.ctypes.data, a.ctypes.data
    sq, sv = 8 * om.nq, 8 * nv
    for i in range(N):
        # one native call per sample, like pin.computeJointTorqueRegressor at regressor.py:49
        call(mp, qp + i * sq, vp + i * sv, ap + i * sv, wp)
        if mode == 0:
            W3[:, i, body_cols] = Wt
            if flags & 1:
                W3[jj, i, c_fv] = v[i]
                W3[jj, i, c_fs] = np.sign(v[i])
            if flags & 2:
                W3[jj, i, c_ia] = a[i]
            if flags & 4:
                W3[jj, i, c_off] = 1.0
        else:
            if len(sel):
                W3[sel[:, None], i, ew_cols[None, :]] = Wt[np.ix_(sel, ew_src)]
            if flags & 1:
                W3[:, i, c_fv] = v[i, :nb]
                W3[:, i, c_fs] = np.sign(v[i, :nb])
            if flags & 2:
                W3[:, i, c_ia] = a[i, :nb]
            if flags & 4:
                W3[:, i, c_off] = 1.0
    return W


def add_coupling_TX40(W, N, v, a):
    """regressor.py:198-227 - three wrist-coupling columns on the rows of joints 5 and 6."""
    v = np.asarray(v)
    a = np.asarray(a)
    out = np.hstack([W, np.zeros((W.shape[0], 3))])
    sg = np.sign(v[:N, 4] + v[:N, 5])
    out[4 * N:5 * N, -3:] = np.stack([a[:N, 5], v[:N, 5], sg], axis=1)
    out[5 * N:6 * N, -3:] = np.stack([a[:N, 4], v[:N, 4], sg], axis=1)
    return out


def get_index_eliminate(W, params_std, tol_e=1e-6):
    """regressor.py:258-279 - squared column norm below an absolute tolerance."""
    sq = np.diag(np.dot(W.T, W))
    names = list(params_std.keys())
    idx_e = [i for i in range(sq.shape[0]) if sq[i] < tol_e]
    params_r = [names[i] for i in range(sq.shape[0]) if not sq[i] < tol_e]
    return idx_e, params_r


def eliminate_non_dynaffect(W, params_std, tol_e=1e-6):
    idx_e, params_r = get_index_eliminate(W, params_std, tol_e)
    return np.delete(W, tuple(idx_e), 1), params_r


def build_regressor_reduced(W, idx_e):
    return np.delete(W, idx_e, 1)


def _split_by_diag(W_e, params_r, tol_qr):
    R = np.linalg.qr(W_e, mode="r")
    d = np.abs(np.diag(R))
    assert d.shape[0] == len(params_r), "params_r does not have same length with R"
    base = [i for i in range(len(params_r)) if d[i] > tol_qr]
    regroup = [i for i in range(len(params_r)) if not d[i] > tol_qr]
    return base, regroup


def get_baseIndex(W_e, params_r, tol_qr=TOL_QR):
    """qrdecomposition.py:274-296."""
    base, _ = _split_by_diag(W_e, params_r, tol_qr)
    return tuple(base)


def _expressions(names_base, names_regroup, beta):
    """qrdecomposition.py:160-183 - 'name + |b|*other' strings for |b| >= 1e-6."""
    out = list(names_base)
    for i in range(len(out)):
        for j in range(beta.shape[1]):
            b = beta[i, j]
            if abs(b) < 1e-6:
                continue
            out[i] = out[i] + (" - " if b < -1e-6 else " + ") + str(abs(b)) + "*" + str(names_regroup[j])
    return out


def _regrouped_qr(W_e, params_r, tol_qr):
    base, regroup = _split_by_diag(W_e, params_r, tol_qr)
    r = len(base)
    W1 = W_e[:, base]
    Wg = np.concatenate([W1, W_e[:, regroup]], axis=1)
    Q, R = np.linalg.qr(Wg)
    R1, R2, Q1 = R[:r, :r], R[:r, r:], Q[:, :r]
    beta = np.around(np.dot(np.linalg.inv(R1), R2), 6)
    return base, regroup, W1, Q1, R1, beta


def get_baseParams(W_e, params_r, params_std=None, tol_qr=TOL_QR):
    """qrdecomposition.py:190-271."""
    base, regroup, W1, Q1, R1, beta = _regrouped_qr(W_e, params_r, tol_qr)
    names = _expressions([params_r[i] for i in base], [params_r[i] for i in regroup], beta)
    W_b = np.dot(Q1, R1)
    assert np.allclose(W1, W_b), "base regressors is wrongly calculated!"
    return W_b, names, base


def double_QR(tau, W_e, params_r, params_std=None, tol_qr=TOL_QR):
    """qrdecomposition.py:89-187."""
    base, regroup, W1, Q1, R1, beta = _regrouped_qr(W_e, params_r, tol_qr)
    phi_b = np.round(np.dot(np.linalg.inv(R1), np.dot(Q1.T, tau)), 6)
    W_b = np.dot(Q1, R1)
    assert np.allclose(W1, W_b), "base regressors is wrongly calculated!"
    nb_, ng_ = [params_r[i] for i in base], [params_r[i] for i in regroup]
    phi_std = None
    if params_std is not None:
        phi_std = [params_std[x] for x in nb_]
        for i in range(len(base)):
            for j in range(beta.shape[1]):
                phi_std[i] = phi_std[i] + beta[i, j] * params_std[ng_[j]]
        phi_std = np.around(phi_std, 5)
    names = _expressions(nb_, ng_, beta)
    bp = dict(zip(names, phi_b))
    if params_std is not None:
        return W_b, bp, names, phi_b, phi_std
    return W_b, bp, names, phi_b


def QR_pivoting(tau, W_e, params_r, tol_qr=TOL_QR):
    """qrdecomposition.py:24-86 (full Q: only usable for small row counts)."""
    from scipy import linalg
    Q, R, P = linalg.qr(W_e, pivoting=True)
    sorted_names = [params_r[P[i]] for i in range(P.shape[0])]
    d = np.abs(np.diag(R))
    rank = 0
    for i in range(d.shape[0]):
        if not d[i] > tol_qr:
            rank = i
            break
    R1, Q1, R2 = R[:rank, :rank], Q[:, :rank], R[:rank, rank:]
    beta = np.around(np.dot(np.linalg.inv(R1), R2), 6)
    phi_b = np.round(np.dot(np.linalg.inv(R1), np.dot(Q1.T, tau)), 6)
    W_b = np.dot(Q1, R1)
    names = _expressions(sorted_names[:rank], sorted_names[rank:], beta)
    return W_b, dict(zip(names, phi_b))


def build_baseRegressor(W_e, idx_base):
    return np.ascontiguousarray(W_e[:, list(idx_base)])


def cond_num(W_b, norm_type=None):
    if norm_type == "fro":
        return np.linalg.cond(W_b, "fro")
    if norm_type == "max_over_min_sigma":
        return np.linalg.cond(W_b, 2) / np.linalg.cond(W_b, -2)
    return np.linalg.cond(W_b)


def relative_stdev(W_b, phi_b, tau):
    """identification_tools.py:204-234."""
    sig2 = np.linalg.norm(tau - np.dot(W_b, phi_b)) ** 2 / (W_b.shape[0] - phi_b.shape[0])
    C = sig2 * np.linalg.inv(np.dot(W_b.T, W_b))
    return np.round(100 * np.sqrt(np.diag(C)) / np.abs(phi_b), 2)


def _block_diag_columns(x, nblocks):
    """(len(x), nblocks) matrix whose column j holds block j of x (len(x) / nblocks entries) in its own rows and zeros
    elsewhere - what regressor.py:319-345 / :436-470 assemble with nested concatenations."""
    x = np.asarray(x, dtype=float).reshape(-1)
    n = len(x) // nblocks
    out = np.zeros((len(x), nblocks))
    for j in range(nblocks):
        out[j * n:(j + 1) * n, j] = x[j * n:(j + 1) * n]
    return out


def _tls_solution(W_tot, mass_load):
    """regressor.py:405-410 / :495-500: last right singular vector, scaled so that its last entry is the payload mass."""
    _, _, Vh = np.linalg.svd(W_tot, full_matrices=False)
    V = Vh.T.conj()
    V_norm = mass_load * np.divide(V[:, -1], V[-1, -1])
    return V_norm, W_tot @ V_norm


def build_total_regressor_current(W_b_u, W_b_l, W_l, I_u, I_l, param_standard_l, param):
    """regressor.py:296-410 - total-least-squares regressor of a payload identification from joint currents:
    [-W_b (unloaded; loaded) | currents, one column per joint | payload columns of the loaded body | its mass column]."""
    n_samples = param["nb_samples"]
    nb_joints = int(len(I_u) / n_samples)
    top = len(W_b_u)
    cur = np.vstack([_block_diag_columns(np.asarray(I_u)[:nb_joints * n_samples], nb_joints),
                     _block_diag_columns(np.asarray(I_l)[:nb_joints * n_samples], nb_joints)])
    if param["has_friction"]:
        width, ks = 12, [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 11]
    elif param["has_actuator_inertia"]:
        width, ks = 14, [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13]
    else:
        width, ks = 10, list(range(9))
    base = param["which_body_loaded"] * width
    W_l_temp = np.zeros((len(W_l), width if width != 10 else 9))
    for k in ks:
        W_l_temp[:, k] = W_l[:, base + k]
    idx_e, _ = get_index_eliminate(W_l_temp, param_standard_l, 1e-6)
    W_e_l = build_regressor_reduced(W_l_temp, idx_e)
    upayload = np.vstack([np.zeros((len(W_l), W_e_l.shape[1])), -W_e_l])
    kpayload = np.vstack([np.zeros((len(W_l), 1)), -W_l[:, base + 9].reshape(-1, 1)])
    W_tot = np.hstack([np.vstack([-W_b_u, -W_b_l]), cur, upayload, kpayload])
    assert top + len(W_b_l) == len(W_tot)
    V_norm, residue = _tls_solution(W_tot, param["mass_load"])
    return W_tot, V_norm, residue


def build_total_regressor_wrench(W_b_u, W_b_l, W_l, tau_u, tau_l, param_standard_l, param):
    """regressor.py:413-500 - the same with external-wrench measurements: six wrench columns instead of the currents,
    the nine payload columns kept without elimination."""
    wr = np.vstack([_block_diag_columns(tau_u, 6), _block_diag_columns(tau_l, 6)])
    base = param["which_body_loaded"] * 10
    upayload = np.vstack([np.zeros((len(W_l), 9)), -W_l[:, base:base + 9]])
    kpayload = np.vstack([np.zeros((len(W_l), 1)), -W_l[:, base + 9].reshape(-1, 1)])
    W_tot = np.hstack([np.vstack([-W_b_u, -W_b_l]), wr, upayload, kpayload])
    V_norm, residue = _tls_solution(W_tot, param["mass_load"])
    return W_tot, V_norm, residue
