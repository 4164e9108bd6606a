"""figaroh.identification.identification_tools drop-in (B200): the preprocessing step immediately ahead of
the regressor - the two functions of /root/reference/src/figaroh/identification/identification_tools.py
that turn a logged position trajectory into the (q, dq, ddq) the regressor consumes.

Same names, arguments and return values as the reference; numpy in -> numpy out, CUDA tensors in -> CUDA
tensors out (so a raw log can go to base parameters without leaving the GPU).  The Butterworth DESIGN
(scipy.signal.butter / lfilter_zi: a handful of coefficients) stays on the host; the filtering and the
differentiation of the N-sample signals run on the GPU (fg_filtfilt / fg_differentiate).
"""
import ctypes as C

import numpy as np
import torch
from scipy import signal

from figaroh_b200 import _lib
from figaroh_b200 import engine as _eng


def _dev(x):
    on_dev = isinstance(x, torch.Tensor) and x.is_cuda
    return _eng.as_device(x), on_dev


def low_pass_filter_data(data, param, nbutter=5):
    """Zero-phase Butterworth low-pass (reference: identification_tools.py:390-424).

    b, a = butter(nbutter, ts * fc / 2); filtfilt(padtype="odd", padlen=3 * (max(len(a), len(b)) - 1));
    5 * nbutter samples removed at both ends."""
    cutoff = param["ts"] * param["cut_off_frequency_butterworth"] / 2
    b, a = signal.butter(nbutter, cutoff, "low")
    padlen = 3 * (max(len(b), len(a)) - 1)
    zi = np.ascontiguousarray(signal.lfilter_zi(b, a), dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    one_d = np.ndim(data) == 1
    x, on_dev = _dev(np.asarray(data, dtype=np.float64).reshape(len(data), -1) if not isinstance(data, torch.Tensor)
                     else data.reshape(data.shape[0], -1))
    N, ncols = x.shape
    if N <= padlen:
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % padlen)
    nbord = 5 * nbutter
    rows = N - 2 * nbord
    if rows <= 0:  # the reference deletes everything (and wraps around for very short signals)
        out = torch.empty((0, ncols), dtype=torch.float64, device=x.device)
    else:
        out = torch.empty((rows, ncols), dtype=torch.float64, device=x.device)
        _lib.check(_lib.lib().fg_filtfilt(x.data_ptr(), N, ncols, x.stride(0), b.ctypes.data, len(b), a.ctypes.data,
                                          len(a), zi.ctypes.data, padlen, nbord, out.data_ptr(), out.stride(0),
                                          _eng._stream()))
    if one_d:
        out = out[:, 0]
    return out if on_dev else out.cpu().numpy()


def calculate_first_second_order_differentiation(model, q, param, dt=None):
    """(q, dq, ddq) from positions (reference: identification_tools.py:334-387).

    dq[i] = pin.difference(q[i], q[i+1]) / dt, ddq[:, j] = np.gradient(dq[:, j]) / dt for j < nq - 1 (the last
    joint's acceleration stays 0, as in the reference), two trailing samples dropped.  Joint-torque mode on
    models whose configuration space is a vector space (nq == nv: pin.difference is a subtraction)."""
    if not param["is_joint_torques"]:
        # external-wrench mode needs pin.difference on the free-flyer (SE3 log): not on the GPU path
        raise NotImplementedError("calculate_first_second_order_differentiation: joint-torque mode only")
    nq = getattr(model, "nq", None)
    qd, on_dev = _dev(q)
    N, ncol = qd.shape
    if nq is not None and (nq != ncol or getattr(model, "nv", nq) != nq):
        raise NotImplementedError("configuration space is not a vector space (nq != nv)")
    dq = torch.empty((N - 2, ncol), dtype=torch.float64, device=qd.device)
    ddq = torch.empty((N - 2, ncol), dtype=torch.float64, device=qd.device)
    if dt is None:
        dts, dtv = float(param["ts"]), None
    else:
        dtv = _eng.as_device(np.asarray(dt, dtype=np.float64).reshape(-1) if not isinstance(dt, torch.Tensor) else dt.reshape(-1))
        if dtv.numel() != N - 1:
            raise ValueError("dt must hold one step per pair of consecutive samples (N - 1 entries)")
        dts = 0.0
    _lib.check(_lib.lib().fg_differentiate(qd.data_ptr(), N, ncol, qd.stride(0), C.c_double(dts),
                                           None if dtv is None else dtv.data_ptr(), dq.data_ptr(), ddq.data_ptr(),
                                           dq.stride(0), _eng._stream()))
    q_out = qd[:N - 2]
    if on_dev:
        return q_out, dq, ddq
    return q_out.cpu().numpy(), dq.cpu().numpy(), ddq.cpu().numpy()


def _factor_of(W_b, tau):
    """(R1, z = Q1^T tau, |residual|^2, rows) of the least-squares problem [W_b | tau], from ONE TSQR pass."""
    Wd = _eng.as_device(W_b)
    r = Wd.shape[1]
    st = _eng.TsqrState(r).accumulate_dense(Wd, _eng.as_device(np.asarray(tau, dtype=np.float64).reshape(-1))
                                            if not isinstance(tau, torch.Tensor) else tau.reshape(-1))
    fused = st.solve(None, (), 0.0)   # no elimination; tol_qr = 0: every column with a non-zero pivot is a base column
    if fused is None:
        diag = st.reduce(np.arange(r))
        base = [i for i in range(r) if diag[i] > 0.0]
        if len(base) != r:
            raise np.linalg.LinAlgError("Singular matrix")
        _, phi, R1, r2 = st.regroup_solve(base, [])
    else:
        if len(fused["base"]) != r:
            raise np.linalg.LinAlgError("Singular matrix")   # what np.linalg.inv(W_b^T W_b) raises in the reference
        phi, R1, r2 = fused["phi"], fused["R1"], fused["resid2"]
    return R1, R1 @ phi, r2, Wd.shape[0]


def relative_stdev(W_b, phi_b, tau):
    """Relative standard deviation (%) of the base parameters (reference: identification_tools.py:204-234).

    |tau - W_b phi_b|^2 = |R1 phi_b - Q1^T tau|^2 + |residual|^2 and (W_b^T W_b)^-1 = R1^-1 R1^-T: one TSQR pass
    over [W_b | tau] on the GPU replaces the tall products of the reference."""
    phi_b = np.asarray(phi_b, dtype=np.float64)
    R1, z, r2, rows = _factor_of(W_b, tau)
    sig_ro_sqr = (float(np.sum((R1 @ phi_b - z) ** 2)) + r2) / (rows - phi_b.shape[0])
    Rinv = np.linalg.inv(R1)
    std_x_sqr = sig_ro_sqr * np.sum(Rinv * Rinv, axis=1)   # diag(R1^-1 R1^-T)
    std_xr = np.zeros(std_x_sqr.shape[0])
    for i in range(std_x_sqr.shape[0]):
        std_xr[i] = np.round(100 * np.sqrt(std_x_sqr[i]) / np.abs(phi_b[i]), 2)
    return std_xr


def weigthed_least_squares(robot, phi_b, W_b, tau_meas, tau_est, param):
    """Weighted least squares (reference: identification_tools.py:291-331; the name keeps the reference's
    spelling).  sigma_j = |tau_meas_j - tau_est_j| / (n_j - len(phi_b)) per joint block, weights 1 / sigma_j on
    the nb_samples = idx_tau_stop[0] rows of every block; the solution pinv(P W_b) (P tau_meas) of the LAST pass
    of the reference's loop is the least-squares solution of the row-scaled system: rows are scaled on the
    device and folded by one TSQR pass."""
    tau_meas = np.asarray(tau_meas, dtype=np.float64).reshape(-1)
    tau_est = np.asarray(tau_est, dtype=np.float64).reshape(-1)
    nq = robot.model.nq
    nb_samples = int(param["idx_tau_stop"][0])
    w = np.zeros(len(tau_meas))
    start_idx = 0
    for ii in range(nq):
        stop = int(param["idx_tau_stop"][ii])
        diff = tau_meas[start_idx:stop] - tau_est[start_idx:stop]
        sigma = np.linalg.norm(diff) / (len(diff) - len(phi_b))
        start_idx = stop
        if sigma == 0.0:  # the reference divides by it: infinite weights, np.linalg.pinv then fails in the SVD
            raise np.linalg.LinAlgError("SVD did not converge")
        w[ii * nb_samples:(ii + 1) * nb_samples] = 1 / sigma
    Wd = _eng.as_device(W_b)
    wd = torch.from_numpy(w).to(Wd.device)
    Ws, ts = Wd * wd[:, None], torch.from_numpy(tau_meas).to(Wd.device) * wd
    # The reference's pinv(P W_b) (P tau) - the minimum-norm solution when the weighted system is rank-deficient -
    # on the small factor: with P W_b = Q R (Q with orthonormal columns) it is pinv(R) Q^T (P tau)
    r = Ws.shape[1]
    st = _eng.TsqrState(r).accumulate_dense(Ws, ts)
    R = st.export().cpu().numpy()[:(r + 1) * (r + 1)].reshape(r + 1, r + 1)
    return np.around(np.linalg.pinv(R[:r, :r]) @ R[:r, r], 6)
