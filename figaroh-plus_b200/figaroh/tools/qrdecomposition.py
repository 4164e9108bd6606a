"""figaroh.tools.qrdecomposition drop-in (B200).

Same functions as /root/reference/src/figaroh/tools/qrdecomposition.py.  The two tall
np.linalg.qr calls of the reference are replaced by ONE Householder TSQR pass over
[W_e | tau] on the GPU (kernel (2)); everything after that (rank decision, regrouping,
beta = R1^-1 R2, phi_b = R1^-1 Q1^T tau) is a function of the small R factor (kernel (3)).
Q is never formed; W_b is returned as the selected columns of W_e, which is what the
reference's Q1 R1 equals (it asserts np.allclose(W1, Q1 R1), qrdecomposition.py:149,269).
"""
import numpy as np
import torch

from figaroh_b200 import engine as _eng

TOL_QR = 1e-8


def _on_device(x):
    return isinstance(x, _eng.DeviceMatrix) or (isinstance(x, torch.Tensor) and x.is_cuda)


def _state_of(W_e, tau=None):
    Wd = _eng.as_device(W_e)
    return Wd, _eng.TsqrState(Wd.shape[1]).accumulate_dense(Wd, tau)


def _columns(W_e, Wd, idx):
    if _on_device(W_e):
        return _eng.DeviceMatrix(Wd).take_columns(idx)
    return np.ascontiguousarray(np.asarray(W_e, dtype=np.float64)[:, list(idx)])


def QR_pivoting(tau, W_e, params_r, tol_qr=TOL_QR):
    """Base parameters from a column-pivoted QR - reference: qrdecomposition.py:24-86.

    The pivoting runs on the GPU on the small R factor of [W_e | tau]; returns (W_b, {expr: phi_b}).
    Quirk kept (qrdecomposition.py:45-51): the rank is the index of the FIRST diagonal entry with
    |R_ii| <= tol_qr and stays 0 when there is none."""
    Wd, st = _state_of(W_e, np.asarray(tau, dtype=np.float64) if not _on_device(tau) else tau)
    p = Wd.shape[1]
    perm, R = st.pivot_qr(np.arange(p))
    params_rsorted = [params_r[int(perm[i])] for i in range(p)]
    numrank_W = 0
    for i in range(p):
        if abs(R[i, i]) > tol_qr:
            continue
        numrank_W = i
        break
    r = numrank_W
    R1inv = np.linalg.inv(R[:r, :r])
    beta = np.around(R1inv @ R[:r, r:p], 6)
    phi_b = np.round(R1inv @ R[:r, p], 6)
    W_b = _columns(W_e, Wd, [int(perm[i]) for i in range(r)])
    params_base = _eng.expressions(params_rsorted[:r], params_rsorted[r:], beta)
    return W_b, dict(zip(params_base, phi_b))


def get_baseIndex(W_e, params_r, tol_qr=TOL_QR):
    """Indices i with |R_ii| > tol_qr in the unpivoted QR of W_e (qrdecomposition.py:274-296)."""
    Wd, st = _state_of(W_e)
    assert Wd.shape[1] == len(params_r), "params_r does not have same length with R"
    diag = st.reduce(np.arange(Wd.shape[1]))
    return tuple(i for i in range(len(params_r)) if diag[i] > tol_qr)


def get_baseParams(W_e, params_r, params_std=None, tol_qr=TOL_QR):
    """(W_b, base-parameter expressions, idx_base) - reference: qrdecomposition.py:190-271."""
    Wd, st = _state_of(W_e)
    assert Wd.shape[1] == len(params_r), "params_r does not have same length with R"
    sol = _eng.solve_state(st, list(params_r), idx_e=[], tol_qr=tol_qr)
    return _columns(W_e, Wd, sol.idx_base), sol.params_base, list(sol.idx_base)


def double_QR(tau, W_e, params_r, params_std=None, tol_qr=TOL_QR):
    """Base parameters and their OLS values - reference: qrdecomposition.py:89-187.

    Returns (W_b, {expr: phi_b}, exprs, phi_b[, phi_std])."""
    Wd, st = _state_of(W_e, np.asarray(tau, dtype=np.float64) if not _on_device(tau) else tau)
    assert Wd.shape[1] == len(params_r), "params_r does not have same length with R"
    sol = _eng.solve_state(st, list(params_r), idx_e=[], tol_qr=tol_qr, params_std=params_std)
    W_b = _columns(W_e, Wd, sol.idx_base)
    base_parameters = dict(zip(sol.params_base, sol.phi_b))
    if params_std is not None:
        return W_b, base_parameters, sol.params_base, sol.phi_b, sol.phi_std
    return W_b, base_parameters, sol.params_base, sol.phi_b


def build_baseRegressor(W_e, idx_base):
    """Columns idx_base of W_e (reference: qrdecomposition.py:299-313)."""
    return _columns(W_e, _eng.as_device(W_e) if _on_device(W_e) else None, list(idx_base))


def cond_num(W_b, norm_type=None):
    """Condition number of W_b (reference: qrdecomposition.py:316-332).

    cond_2(W_b) equals cond_2 of its R factor, so the tall matrix is reduced on the GPU
    and only the small triangular factor is decomposed."""
    if norm_type == "fro" and W_b.shape[0] != W_b.shape[1]:
        # np.linalg.cond(W, "fro") inverts W: the reference raises for a tall matrix
        raise np.linalg.LinAlgError("Last 2 dimensions of the array must be square")
    Wd, st = _state_of(W_b)
    p = Wd.shape[1]
    st.reduce(np.arange(p))
    _, _, R1, _ = st.regroup_solve(np.arange(p), np.zeros(0, dtype=np.int32))
    if norm_type == "fro":
        # ||W||_F = ||R||_F ; ||pinv(W)||_F = ||R^-1||_F
        return float(np.linalg.norm(R1, "fro") * np.linalg.norm(np.linalg.inv(R1), "fro"))
    sv = np.linalg.svd(R1, compute_uv=False)
    if norm_type == "max_over_min_sigma":
        # the reference evaluates cond(W,2) / cond(W,-2) = (smax/smin)^2
        return float((sv.max() / sv.min()) ** 2)
    return float(sv.max() / sv.min())
