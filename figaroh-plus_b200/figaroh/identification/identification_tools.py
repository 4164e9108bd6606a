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
