"""Column-wise signal operations of the identification scripts on the GPU:

    medfilt   scipy.signal.medfilt(x[:, j], k)                      examples/tiago/identification.py:70,80
    filtfilt  scipy.signal.filtfilt(b, a, x, padtype="odd", padlen) examples/tiago/identification.py:71-87
    gradient  np.gradient(x[:, j]) [/ np.gradient(t)]               examples/tiago/identification.py:91-98
    decimate  scipy.signal.decimate(x[:, j], q, zero_phase=True)    examples/tiago/identification.py:142-187

numpy in -> numpy out, CUDA tensors / DeviceMatrix in -> CUDA tensors out.  Filter DESIGN (a handful of
coefficients: scipy.signal.cheby1 / lfilter_zi) stays on the host; the N-sample work runs in
libfigaroh_b200.so (fg_medfilt, fg_filtfilt, fg_gradient, fg_filtfilt_decimate).
"""
import numpy as np
import torch
from scipy import signal as _sp

from . import _lib
from . import engine as _eng


def _prep(x):
    if isinstance(x, _eng.DeviceMatrix):
        x = x.t
    on_dev = isinstance(x, torch.Tensor) and x.is_cuda
    one_d = (x.dim() if isinstance(x, torch.Tensor) else np.ndim(x)) == 1
    t = _eng.as_device(x if isinstance(x, torch.Tensor) else np.asarray(x, dtype=np.float64))
    t = t.reshape(t.shape[0], -1)
    return t, on_dev, one_d


def _finish(y, on_dev, one_d):
    if one_d:
        y = y[:, 0]
    return y if on_dev else y.cpu().numpy()


def medfilt(x, kernel_size=3):
    t, on_dev, one_d = _prep(x)
    y = torch.empty_like(t)
    _lib.check(_lib.lib().fg_medfilt(t.data_ptr(), t.shape[0], t.shape[1], t.stride(0), int(kernel_size), y.data_ptr(),
                                     y.stride(0), _eng._stream()))
    return _finish(y, on_dev, one_d)


def _design(b, a):
    b = np.ascontiguousarray(b, dtype=np.float64)
    a = np.ascontiguousarray(a, dtype=np.float64)
    zi = np.ascontiguousarray(_sp.lfilter_zi(b, a), dtype=np.float64)
    return b, a, zi


def filtfilt(b, a, x, padlen=None):
    """scipy.signal.filtfilt(b, a, x, axis=0, padtype="odd", padlen=padlen)."""
    b, a, zi = _design(b, a)
    if padlen is None:
        padlen = 3 * max(len(a), len(b))
    t, on_dev, one_d = _prep(x)
    if t.shape[0] <= padlen:
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % padlen)
    y = torch.empty_like(t)
    _lib.check(_lib.lib().fg_filtfilt(t.data_ptr(), t.shape[0], t.shape[1], t.stride(0), b.ctypes.data, len(b),
                                      a.ctypes.data, len(a), zi.ctypes.data, int(padlen), 0, y.data_ptr(), y.stride(0),
                                      _eng._stream()))
    return _finish(y, on_dev, one_d)


def gradient(x, t=None):
    """np.gradient along axis 0 for every column; divided by np.gradient(t) when t (N,) is given."""
    xt, on_dev, one_d = _prep(x)
    td = None if t is None else _eng.as_device(t if isinstance(t, torch.Tensor) else np.asarray(t, dtype=np.float64)).reshape(-1)
    y = torch.empty_like(xt)
    _lib.check(_lib.lib().fg_gradient(xt.data_ptr(), xt.shape[0], xt.shape[1], xt.stride(0),
                                      None if td is None else td.data_ptr(), y.data_ptr(), y.stride(0), _eng._stream()))
    return _finish(y, on_dev, one_d)


def sosfiltfilt(sos, x, padlen=None, step=1):
    """scipy.signal.sosfiltfilt(sos, x, axis=0, padtype="odd", padlen)[::step] for every column."""
    sos = np.ascontiguousarray(sos, dtype=np.float64).reshape(-1, 6)
    nsec = sos.shape[0]
    zi = np.ascontiguousarray(_sp.sosfilt_zi(sos), dtype=np.float64)
    if padlen is None:  # scipy/signal/_signaltools.py sosfiltfilt()
        ntaps = 2 * nsec + 1 - min(int((sos[:, 2] == 0).sum()), int((sos[:, 5] == 0).sum()))
        padlen = 3 * ntaps
    t, on_dev, one_d = _prep(x)
    if t.shape[0] <= padlen:
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % padlen)
    step = int(step)
    rows = (t.shape[0] + step - 1) // step
    y = torch.empty((rows, t.shape[1]), dtype=torch.float64, device=t.device)
    _lib.check(_lib.lib().fg_sosfiltfilt(t.data_ptr(), t.shape[0], t.shape[1], t.stride(0), sos.ctypes.data, nsec,
                                         zi.ctypes.data, int(padlen), 0, step, y.data_ptr(), y.stride(0), _eng._stream()))
    return _finish(y, on_dev, one_d)


def decimate(x, q, n=None, zero_phase=True):
    """scipy.signal.decimate(x, q, n, ftype="iir", axis=0, zero_phase=True) for every column: Chebyshev type I
    low-pass of order n (default 8) as second-order sections, forwards and backwards, then every q-th sample."""
    if not zero_phase:
        raise NotImplementedError("decimate: only the zero-phase (filtfilt) variant the scripts use")
    q = int(q)
    n = 8 if n is None else int(n)
    sos = _sp.cheby1(n, 0.05, 0.8 / q, output="sos")      # scipy/signal/_signaltools.py decimate(), ftype="iir"
    return sosfiltfilt(sos, x, step=q)


def decimate_blocks(W, nblocks, q=10):
    """Decimate every column of each of the `nblocks` equal row blocks of W separately and stack the results
    (examples/tiago/identification.py:157-187: one decimation per active joint's row block)."""
    if isinstance(W, _eng.DeviceMatrix):
        W = W.t
    on_dev = isinstance(W, torch.Tensor) and W.is_cuda
    t = _eng.as_device(W if isinstance(W, torch.Tensor) else np.asarray(W, dtype=np.float64))
    t = t.reshape(t.shape[0], -1)
    n = t.shape[0] // nblocks
    out = torch.cat([decimate(t[i * n:(i + 1) * n], q) for i in range(nblocks)], dim=0)
    return out if on_dev else out.cpu().numpy()
