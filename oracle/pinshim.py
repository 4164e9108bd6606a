"""A stand-in ``pinocchio`` module so the reference's UNMODIFIED Python files can run here.

TEST INFRASTRUCTURE ONLY.  Used in this container (where /root/reference exists) to
(1) pin the restated C arithmetic against the reference's golden CSVs and
(2) generate the fixtures under tests/golden/ (tests/golden/make_golden.py).

``install()`` registers fake ``pinocchio``, ``pinocchio.robot_wrapper``,
``pinocchio.visualize`` and ``quadprog`` modules; ``load_reference(name)`` imports
one reference source file by path (the ``figaroh`` package itself cannot be
imported: its ``tools/__init__`` pulls in meshcat/hppfcl).
The model objects come from the product's URDF loader (the loader is part of
what the TIAGo golden pins); the arithmetic comes from oracle/fg_oracle.c.
"""
import importlib.util
import os
import sys
import types

import numpy as np

from . import kernels

REF_ROOT = "/root/reference"
_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "figaroh-plus_b200")


def _product():
    if _PKG not in sys.path:
        sys.path.insert(0, _PKG)
    import figaroh_b200.model as pm
    import figaroh_b200.urdf as pu
    return pm, pu


def _oracle_model(model):
    om = getattr(model, "_oracle_model", None)
    if om is None or om.njoints != model.njoints:
        pm, _ = _product()
        om = kernels.OracleModel(pm.flatten_model(model))
        model._oracle_model = om
    return om


def refresh(model):
    """Call after mutating a model (e.g. jointPlacements) so the C view is rebuilt."""
    model._oracle_model = None


def computeJointTorqueRegressor(model, data, q, v, a):
    return kernels.joint_torque_regressor(_oracle_model(model), q, v, a)


def rnea(model, data, q, v, a):
    return kernels.rnea(_oracle_model(model), q, v, a)


def difference(model, q0, q1):
    om = _oracle_model(model)
    out = np.zeros(model.nv)
    for i in range(1, om.njoints):
        jt, iq, iv = om.jtype[i], om.idx_q[i], om.idx_v[i]
        if jt in (1, 3):
            out[iv] = q1[iq] - q0[iq]
        elif jt == 2:
            c0, s0, c1, s1 = q0[iq], q0[iq + 1], q1[iq], q1[iq + 1]
            out[iv] = np.arctan2(s1 * c0 - c1 * s0, c1 * c0 + s1 * s0)
        else:
            raise NotImplementedError("difference() for free-flyer joints")
    return out


def neutral(model):
    om = _oracle_model(model)
    q = np.zeros(model.nq)
    for i in range(1, om.njoints):
        if om.jtype[i] == 2:
            q[om.idx_q[i]] = 1.0
        elif om.jtype[i] == 4:
            q[om.idx_q[i] + 6] = 1.0
    return q


def computeSubtreeMasses(model, data):
    om = _oracle_model(model)
    mass = om.mass.copy()
    for i in range(om.njoints - 1, 0, -1):
        mass[om.parent[i]] += mass[i]
    data.mass = mass
    return mass


class RobotWrapper:
    def initFromURDF(self, filename, package_dirs=None, root_joint=None, verbose=False, meshLoader=None):
        pm, pu = _product()
        self.model = pu.build_model_from_urdf(filename, root_joint=root_joint)
        self.data = self.model.createData()
        self.collision_model = None
        self.visual_model = None
        self.q0 = neutral(self.model)
        self.v0 = np.zeros(self.model.nv)


def install():
    pm, _ = _product()
    pin = types.ModuleType("pinocchio")
    pin.computeJointTorqueRegressor = computeJointTorqueRegressor
    pin.rnea = rnea
    pin.difference = difference
    pin.neutral = neutral
    pin.computeSubtreeMasses = computeSubtreeMasses
    pin.JointModelFreeFlyer = pm.JointModelFreeFlyer
    pin.SE3 = pm.SE3
    rw = types.ModuleType("pinocchio.robot_wrapper")
    rw.RobotWrapper = RobotWrapper
    vis = types.ModuleType("pinocchio.visualize")
    vis.GepettoVisualizer = type("GepettoVisualizer", (), {})
    vis.MeshcatVisualizer = type("MeshcatVisualizer", (), {})
    pin.robot_wrapper = rw
    pin.visualize = vis
    pin.RobotWrapper = RobotWrapper
    sys.modules["pinocchio"] = pin
    sys.modules["pinocchio.robot_wrapper"] = rw
    sys.modules["pinocchio.visualize"] = vis
    sys.modules.setdefault("quadprog", types.ModuleType("quadprog"))
    return pin


_REF_FILES = {
    "regressor": "src/figaroh/tools/regressor.py",
    "qrdecomposition": "src/figaroh/tools/qrdecomposition.py",
    "robot": "src/figaroh/tools/robot.py",
    "randomdata": "src/figaroh/tools/randomdata.py",
    "identification_tools": "src/figaroh/identification/identification_tools.py",
}


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "src", "figaroh"))


def load_reference(name):
    """Import one unmodified reference source file by path (container only)."""
    install()
    key = "_figaroh_ref_" + name
    if key in sys.modules:
        return sys.modules[key]
    path = os.path.join(REF_ROOT, _REF_FILES[name])
    spec = importlib.util.spec_from_file_location(key, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod
