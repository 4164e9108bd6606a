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
