"""CPU restatement of the reference's preprocessing step (test infrastructure only; see oracle/__init__.py).

Follows /root/reference/src/figaroh/identification/identification_tools.py:334-424.  scipy.signal is the
reference's own dependency for the filter and is used as is; pin.difference is restated for vector-space
joints (nq == nv), where it is a subtraction."""
import numpy as np
from scipy import signal


def calculate_first_second_order_differentiation(nq, q, param, dt=None):
    """identification_tools.py:334-387 (joint-torque branch)."""
    q = np.asarray(q, dtype=float)
    dq = np.zeros([q.shape[0] - 1, q.shape[1]])
    ddq = np.zeros([q.shape[0] - 1, q.shape[1]])
    if dt is None:
        dt = param["ts"]
        for ii in range(q.shape[0] - 1):
            dq[ii, :] = (q[ii + 1, :] - q[ii, :]) / dt
    else:
        for ii in range(q.shape[0] - 1):
            dq[ii, :] = (q[ii + 1, :] - q[ii, :]) / dt[ii]
    for jj in range(nq - 1):  # the reference's range(model.nq - 1): the last joint keeps ddq = 0
        ddq[:, jj] = np.gradient(dq[:, jj], edge_order=1) / dt
    q = np.delete(q, len(q) - 1, 0)
    q = np.delete(q, len(q) - 1, 0)
    dq = np.delete(dq, len(dq) - 1, 0)
    ddq = np.delete(ddq, len(ddq) - 1, 0)
    return q, dq, ddq


def low_pass_filter_data(data, param, nbutter=5):
    """identification_tools.py:390-424."""
    cutoff = param["ts"] * param["cut_off_frequency_butterworth"] / 2
    b, a = signal.butter(nbutter, cutoff, "low")
    padlen = 3 * (max(len(b), len(a)) - 1)
    data = signal.filtfilt(b, a, data, axis=0, padtype="odd", padlen=padlen)
    nbord = 5 * nbutter
    data = np.delete(data, np.s_[0:nbord], axis=0)
    end_slice = slice(data.shape[0] - nbord, data.shape[0])
    data = np.delete(data, end_slice, axis=0)
    return data


def relative_stdev(W_b, phi_b, tau):
    """identification_tools.py:204-234."""
    sig_ro_sqr = np.linalg.norm((tau - np.dot(W_b, phi_b))) ** 2 / (W_b.shape[0] - phi_b.shape[0])
    C_x = sig_ro_sqr * np.linalg.inv(np.dot(W_b.T, W_b))
    std_x_sqr = np.diag(C_x)
    std_xr = np.zeros(std_x_sqr.shape[0])
    for i in range(std_x_sqr.shape[0]):
        std_xr[i] = np.round(100 * np.sqrt(std_x_sqr[i]) / np.abs(phi_b[i]), 2)
    return std_xr


def weigthed_least_squares(nq, phi_b, W_b, tau_meas, tau_est, param):
    """identification_tools.py:291-331 (dense P as in the reference: small problems only)."""
    sigma = np.zeros(nq)
    P = np.zeros((len(tau_meas), len(tau_meas)))
    nb_samples = int(param["idx_tau_stop"][0])
    start_idx = int(0)
    for ii in range(nq):
        tau_slice = slice(int(start_idx), int(param["idx_tau_stop"][ii]))
        diff = (tau_meas[tau_slice] - tau_est[tau_slice])
        denom = len(tau_meas[tau_slice]) - len(phi_b)
        sigma[ii] = np.linalg.norm(diff) / denom
        start_idx = param["idx_tau_stop"][ii]
        for jj in range(nb_samples):
            idx = jj + ii * nb_samples
            P[idx, idx] = 1 / sigma[ii]
        phi_b = np.matmul(np.linalg.pinv(np.matmul(P, W_b)), np.matmul(P, tau_meas))
    return np.around(phi_b, 6)
