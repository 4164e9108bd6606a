"""figaroh.tools.regressor drop-in (B200).

Same functions, arguments and return values as
/root/reference/src/figaroh/tools/regressor.py; the arithmetic runs in the CUDA library
(kernel (1): batched joint-torque regressor; kernel (2): TSQR state for column norms).

Inputs may be numpy arrays (results come back as fresh float64 C-contiguous numpy
arrays, like the reference) or CUDA tensors / DeviceMatrix handles (results stay in
HBM as DeviceMatrix, so a whole pipeline can run without the tall matrix crossing PCIe).
Set ``figaroh.tools.regressor.RETURN_DEVICE = True`` to get DeviceMatrix results from
numpy inputs as well.
"""
import numpy as np
import torch

from figaroh_b200 import _lib
from figaroh_b200 import engine as _eng
from figaroh_b200._lib import FgError
from figaroh_b200._lib import check as _check

RETURN_DEVICE = False


def _device_model(robot):
    """DeviceModel of robot.model, cached on the model and keyed on (content of the flattened arrays, CUDA device):
    a script that edits jointPlacements / gravity after the first call, or switches GPU, gets a fresh upload."""
    from figaroh_b200.model import flatten_model
    model = robot.model
    fm = flatten_model(model)
    key = (hash(b"".join(np.ascontiguousarray(x).tobytes() for x in (fm.parent, fm.jtype, fm.axis, fm.plc_R, fm.plc_p,
                                                                      fm.idx_q, fm.idx_v, fm.gravity, fm.mass))),
           torch.cuda.current_device() if torch.cuda.is_available() else -1)
    cached = getattr(model, "_fg_device_model", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    dm = _eng.DeviceModel(fm)
    try:
        model._fg_device_model = (key, dm)
    except Exception:  # noqa: BLE001 - e.g. a pinocchio Model without __dict__
        pass
    return dm


def _on_device(x):
    return isinstance(x, _eng.DeviceMatrix) or (isinstance(x, torch.Tensor) and x.is_cuda)


def _wrap(t, like_device):
    if like_device or RETURN_DEVICE:
        return _eng.DeviceMatrix(t)
    return t.cpu().numpy()


def build_regressor_basic(robot, q, v, a, param, tau=None):
    """Tall regressor of the standard parameters (reference: regressor.py:20-194).

    Rows are joint-major (row j*N+i), columns follow get_standard_parameters.  `tau` is
    accepted and ignored, as in the reference."""
    N = len(q)
    # evaluated by the reference too (regressor.py:36-39); keeps duck-typed models honest
    _ = [jj for jj in range(len(robot.model.inertias.tolist())) if robot.model.inertias.tolist()[jj].mass != 0]
    cfg = _eng.RegressorConfig.from_param(param)
    dm = _device_model(robot)
    try:
        W = dm.materialize(cfg, q, v, a)
    except FgError as e:
        if e.code == -2:  # what numpy raises in the reference for free-flyer models in joint-torque mode
            raise ValueError(str(e)) from None
        raise
    assert W.shape[0] % max(N, 1) == 0
    return _wrap(W, _on_device(q))


def add_coupling_TX40(W, model, data, N, nq, nv, njoints, q, v, a):
    """Append the Iam6 / fvm6 / fsm6 wrist-coupling columns (reference: regressor.py:198-227)."""
    dev = _on_device(W)
    Wd = _eng.as_device(W)
    vd, ad = _eng.as_device(v), _eng.as_device(a)
    out = torch.zeros((Wd.shape[0], Wd.shape[1] + 3), dtype=torch.float64, device=Wd.device)
    out[:, :-3] = Wd
    sg = torch.sign(vd[:N, 4] + vd[:N, 5])
    out[4 * N:5 * N, -3], out[4 * N:5 * N, -2], out[4 * N:5 * N, -1] = ad[:N, 5], vd[:N, 5], sg
    out[5 * N:6 * N, -3], out[5 * N:6 * N, -2], out[5 * N:6 * N, -1] = ad[:N, 4], vd[:N, 4], sg
    return _wrap(out, dev)


def _column_sq_norms(W):
    """diag(W^T W) (regressor.py:243,271) by a column reduction on the device: any width, no QR."""
    Wd = _eng.as_device(W)
    out = np.empty(Wd.shape[1])
    _check(_lib.lib().fg_colsq_dense(Wd.data_ptr(), Wd.stride(0), Wd.shape[0], Wd.shape[1], out.ctypes.data,
                                     _eng._stream()))
    return out


def eliminate_non_dynaffect(W, params_std, tol_e=1e-6):
    """Drop columns whose squared norm is below tol_e (reference: regressor.py:230-255)."""
    idx_e, params_r = get_index_eliminate(W, params_std, tol_e)
    return build_regressor_reduced(W, tuple(idx_e)), params_r


def get_index_eliminate(W, params_std, tol_e=1e-6):
    """Indices of columns with squared norm < tol_e and names of the others (regressor.py:258-279)."""
    col_norm = _column_sq_norms(W)
    names = list(params_std.keys())
    idx_e, params_r = [], []
    for i in range(col_norm.shape[0]):
        if col_norm[i] < tol_e:
            idx_e.append(i)
        else:
            params_r.append(names[i])
    return idx_e, params_r


def build_regressor_reduced(W, idx_e):
    """W without the columns idx_e (reference: regressor.py:282-293)."""
    if _on_device(W):
        Wm = W if isinstance(W, _eng.DeviceMatrix) else _eng.DeviceMatrix(W)
        return Wm.drop_columns(list(idx_e))
    return np.delete(W, idx_e, 1)


# ---- total-least-squares payload regressors (reference: regressor.py:296-500) --------------------------------
def _block_columns(x, nblocks):
    """(len(x), nblocks) device matrix: column j carries block j of x in its own rows, zeros elsewhere."""
    xd = _eng.as_device(x).reshape(-1)
    n = xd.shape[0] // nblocks
    out = torch.zeros((xd.shape[0], nblocks), dtype=torch.float64, device=xd.device)
    for j in range(nblocks):
        out[j * n:(j + 1) * n, j] = xd[j * n:(j + 1) * n]
    return out


def _tls_finish(W_tot, mass_load, dev):
    """SVD step of regressor.py:405-410 on the R factor: W_tot = Q R has the right singular vectors of R, so the
    tall matrix is reduced once on the GPU (TSQR) and only the small triangle is decomposed."""
    p = W_tot.shape[1]
    st = _eng.TsqrState(p).accumulate_dense(W_tot, None)
    st.reduce(np.arange(p))
    _, _, R, _ = st.regroup_solve(np.arange(p), np.zeros(0, dtype=np.int32))
    _, _, Vh = np.linalg.svd(R, full_matrices=False)
    V = Vh.T.conj()
    V_norm = mass_load * np.divide(V[:, -1], V[-1, -1])
    residue = W_tot @ torch.from_numpy(V_norm).to(W_tot.device)
    if dev or RETURN_DEVICE:
        return _eng.DeviceMatrix(W_tot), V_norm, residue
    return W_tot.cpu().numpy(), V_norm, residue.cpu().numpy()


def build_total_regressor_current(W_b_u, W_b_l, W_l, I_u, I_l, param_standard_l, param):
    """Total-least-squares regressor of a payload identification from joint currents (regressor.py:296-410).

    Returns (W_tot, V_norm, residue) like the reference."""
    dev = _on_device(W_l)
    Wbu, Wbl, Wl = _eng.as_device(W_b_u), _eng.as_device(W_b_l), _eng.as_device(W_l)
    n_samples = param["nb_samples"]
    nb_joints = int(len(I_u) / n_samples)
    cur = torch.cat([_block_columns(_eng.as_device(I_u).reshape(-1)[:nb_joints * n_samples], nb_joints),
                     _block_columns(_eng.as_device(I_l).reshape(-1)[:nb_joints * n_samples], nb_joints)], dim=0)
    if param["has_friction"]:
        width, ks, wt = 12, [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 11], 12
    elif param["has_actuator_inertia"]:
        width, ks, wt = 14, [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13], 14
    else:
        width, ks, wt = 10, list(range(9)), 9
    base = param["which_body_loaded"] * width
    W_l_temp = torch.zeros((Wl.shape[0], wt), dtype=torch.float64, device=Wl.device)
    for k in ks:
        W_l_temp[:, k] = Wl[:, base + k]
    idx_e, _ = get_index_eliminate(_eng.DeviceMatrix(W_l_temp), param_standard_l, 1e-6)
    keep = [k for k in range(wt) if k not in set(idx_e)]
    W_e_l = W_l_temp[:, keep]
    zeros = torch.zeros((Wl.shape[0], W_e_l.shape[1] + 1), dtype=torch.float64, device=Wl.device)
    payload = torch.cat([zeros, torch.cat([-W_e_l, -Wl[:, base + 9:base + 10]], dim=1)], dim=0)
    W_tot = torch.cat([torch.cat([-Wbu, -Wbl], dim=0), cur, payload], dim=1).contiguous()
    return _tls_finish(W_tot, param["mass_load"], dev)


def build_total_regressor_wrench(W_b_u, W_b_l, W_l, tau_u, tau_l, param_standard_l, param):
    """Total-least-squares regressor with external-wrench measurements (regressor.py:413-500)."""
    dev = _on_device(W_l)
    Wbu, Wbl, Wl = _eng.as_device(W_b_u), _eng.as_device(W_b_l), _eng.as_device(W_l)
    wr = torch.cat([_block_columns(tau_u, 6), _block_columns(tau_l, 6)], dim=0)
    base = param["which_body_loaded"] * 10
    zeros = torch.zeros((Wl.shape[0], 10), dtype=torch.float64, device=Wl.device)
    payload = torch.cat([zeros, -Wl[:, base:base + 10]], dim=0)
    W_tot = torch.cat([torch.cat([-Wbu, -Wbl], dim=0), wr, payload], dim=1).contiguous()
    return _tls_finish(W_tot, param["mass_load"], dev)
