"""ctypes front end of oracle/fg_oracle.c (the restated pinocchio arithmetic).

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.
"""
import ctypes as C
import json

import numpy as np

from . import build as _build

_I32 = ("parent", "jtype", "idx_q", "idx_v")
_F64 = ("axis", "plc_R", "plc_p", "gravity", "mass", "lever", "inertia")


class _CModel(C.Structure):
    _fields_ = [("njoints", C.c_int), ("nq", C.c_int), ("nv", C.c_int),
                ("parent", C.c_void_p), ("jtype", C.c_void_p), ("axis", C.c_void_p),
                ("plc_R", C.c_void_p), ("plc_p", C.c_void_p), ("idx_q", C.c_void_p), ("idx_v", C.c_void_p),
                ("gravity", C.c_void_p), ("mass", C.c_void_p), ("lever", C.c_void_p), ("inertia", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.orc_joint_torque_regressor.argtypes = [C.c_void_p] * 5
        _lib.orc_rnea.argtypes = [C.c_void_p] * 5
        _lib.orc_build_regressor_basic.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long,
                                                   C.c_int, C.c_int, C.c_int, C.c_void_p]
    return _lib


class OracleModel:
    """Flattened kinematic tree held as numpy arrays + the C struct pointing at them."""

    def __init__(self, src):
        if isinstance(src, str):
            with open(src) as f:
                src = json.load(f)
        get = (lambda k: src[k]) if isinstance(src, dict) else (lambda k: getattr(src, k))
        self.name = get("name")
        self.names = list(get("names"))
        self.njoints, self.nq, self.nv = int(get("njoints")), int(get("nq")), int(get("nv"))
        for k in _I32:
            setattr(self, k, np.ascontiguousarray(get(k), dtype=np.int32))
        for k in _F64:
            setattr(self, k, np.ascontiguousarray(get(k), dtype=np.float64))
        self._c = _CModel(self.njoints, self.nq, self.nv,
                          *[getattr(self, k).ctypes.data for k in
                            ("parent", "jtype", "axis", "plc_R", "plc_p", "idx_q", "idx_v",
                             "gravity", "mass", "lever", "inertia")])

    @property
    def ptr(self):
        return C.addressof(self._c)

    def dynamic_parameters(self):
        """(n,10) pinocchio-ordered [m, mc, Ixx,Ixy,Iyy,Ixz,Iyz,Izz] with inertia at the joint origin."""
        out = np.zeros((self.njoints - 1, 10))
        for i in range(1, self.njoints):
            m, c, I = self.mass[i], self.lever[i], self.inertia[i]
            Ic = np.array([[I[0], I[1], I[3]], [I[1], I[2], I[4]], [I[3], I[4], I[5]]])
            I0 = Ic - m * (np.outer(c, c) - c.dot(c) * np.eye(3))
            out[i - 1] = [m, m * c[0], m * c[1], m * c[2], I0[0, 0], I0[0, 1], I0[1, 1], I0[0, 2], I0[1, 2], I0[2, 2]]
        return out


def _vec(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def joint_torque_regressor(model: OracleModel, q, v, a):
    """pin.computeJointTorqueRegressor(model, data, q, v, a) -> (nv, 10(njoints-1))."""
    q, v, a = _vec(q), _vec(v), _vec(a)
    assert q.shape == (model.nq,) and v.shape == (model.nv,) and a.shape == (model.nv,)
    W = np.empty((model.nv, 10 * (model.njoints - 1)))
    rc = lib().orc_joint_torque_regressor(model.ptr, q.ctypes.data, v.ctypes.data, a.ctypes.data, W.ctypes.data)
    assert rc == 0
    return W


def rnea(model: OracleModel, q, v, a):
    q, v, a = _vec(q), _vec(v), _vec(a)
    tau = np.zeros(model.nv)
    rc = lib().orc_rnea(model.ptr, q.ctypes.data, v.ctypes.data, a.ctypes.data, tau.ctypes.data)
    assert rc == 0
    return tau


def build_regressor_basic_c(model: OracleModel, q, v, a, mode=0, flags=0, ft_mask=0x3F):
    """Whole tall regressor in one C call (same result as ref_path.build_regressor_basic)."""
    q, v, a = _vec(q), _vec(v), _vec(a)
    N = q.shape[0]
    nb = model.njoints - 1
    rows = (model.nv if mode == 0 else 6) * N
    W = np.empty((rows, 14 * nb))
    rc = lib().orc_build_regressor_basic(model.ptr, q.ctypes.data, v.ctypes.data, a.ctypes.data, N,
                                         mode, flags, ft_mask, W.ctypes.data)
    if rc == -2:
        raise ValueError("joint-torque mode needs nv == njoints-1 (regressor.py:53)")
    assert rc == 0
    return W
