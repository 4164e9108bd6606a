"""Kinematic-tree model objects and the flattening loader.

Two jobs:

1. ``Model`` and friends: a minimal, duck-typed stand-in for the attributes of
   ``pinocchio.Model`` that the FIGAROH hot path reads (``nq, nv, njoints,
   names, parents, jointPlacements, joints, inertias, gravity``; reference call
   sites: /root/reference/src/figaroh/tools/regressor.py:36-49,
   /root/reference/src/figaroh/tools/robot.py:105-121).  It exists because
   pinocchio is neither in this image nor on the GPU box; when pinocchio *is*
   installed its own ``Model`` is accepted by :func:`flatten_model` unchanged.

2. ``flatten_model``: turns either kind of model into ``FlatModel`` - plain
   contiguous numpy arrays (parents, joint types, axes, placements, idx_q/idx_v,
   inertias, gravity) that the C-ABI ``fg_model_create`` copies to the device.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field

import numpy as np

# joint type codes (FG_JT_* in include/figaroh_b200.h)
JT_NONE = 0          # universe
JT_REVOLUTE = 1      # nq=1, nv=1, rotation about `axis` (any unit axis)
JT_REVOLUTE_UNB = 2  # nq=2 (cos,sin), nv=1  - URDF "continuous"
JT_PRISMATIC = 3     # nq=1, nv=1, translation along `axis`
JT_FREEFLYER = 4     # nq=7 (xyz, quat xyzw), nv=6 (local linear, angular)

_JT_NQ = {JT_NONE: 0, JT_REVOLUTE: 1, JT_REVOLUTE_UNB: 2, JT_PRISMATIC: 1, JT_FREEFLYER: 7}
_JT_NV = {JT_NONE: 0, JT_REVOLUTE: 1, JT_REVOLUTE_UNB: 1, JT_PRISMATIC: 1, JT_FREEFLYER: 6}

_SHORTNAMES = {
    JT_REVOLUTE: "JointModelRevoluteUnaligned",
    JT_REVOLUTE_UNB: "JointModelRevoluteUnboundedUnaligned",
    JT_PRISMATIC: "JointModelPrismaticUnaligned",
    JT_FREEFLYER: "JointModelFreeFlyer",
}


def _aligned_suffix(axis):
    for k, s in enumerate("XYZ"):
        e = np.zeros(3)
        e[k] = 1.0
        if np.allclose(axis, e, rtol=1e-12, atol=0.0):
            return s
    return None


class SE3:
    """Rigid placement (rotation, translation); child frame expressed in parent."""

    __slots__ = ("rotation", "translation")

    def __init__(self, rotation=None, translation=None):
        self.rotation = np.eye(3) if rotation is None else np.array(rotation, dtype=np.float64).reshape(3, 3)
        self.translation = np.zeros(3) if translation is None else np.array(translation, dtype=np.float64).reshape(3)

    @staticmethod
    def Identity():
        return SE3()

    def __mul__(self, other):
        if isinstance(other, SE3):
            return SE3(self.rotation @ other.rotation, self.rotation @ other.translation + self.translation)
        return NotImplemented

    def act_inertia(self, Y: "Inertia") -> "Inertia":
        R = self.rotation
        return Inertia(Y.mass, R @ Y.lever + self.translation, R @ Y.inertia @ R.T)

    def copy(self):
        return SE3(self.rotation.copy(), self.translation.copy())


def _skew_square(v):
    """[v]x^2 = v v^T - |v|^2 I."""
    return np.outer(v, v) - np.dot(v, v) * np.eye(3)


class Inertia:
    """Spatial inertia (mass, centre of mass `lever`, rotational inertia about the CoM)."""

    __slots__ = ("mass", "lever", "inertia")

    def __init__(self, mass=0.0, lever=None, inertia=None):
        self.mass = float(mass)
        self.lever = np.zeros(3) if lever is None else np.array(lever, dtype=np.float64).reshape(3)
        self.inertia = np.zeros((3, 3)) if inertia is None else np.array(inertia, dtype=np.float64).reshape(3, 3)

    @staticmethod
    def Zero():
        return Inertia()

    def __add__(self, other: "Inertia") -> "Inertia":
        eps = np.finfo(np.float64).eps
        mab = self.mass + other.mass
        mab_inv = 1.0 / max(mab, eps)
        ab = self.lever - other.lever
        lever = (self.mass * self.lever + other.mass * other.lever) * mab_inv
        inertia = self.inertia + other.inertia - (self.mass * other.mass * mab_inv) * _skew_square(ab)
        return Inertia(mab, lever, inertia)

    def toDynamicParameters(self):
        """[m, mc_x, mc_y, mc_z, Ixx, Ixy, Iyy, Ixz, Iyz, Izz], inertia taken at the joint origin."""
        I0 = self.inertia - self.mass * _skew_square(self.lever)
        out = np.empty(10)
        out[0] = self.mass
        out[1:4] = self.mass * self.lever
        out[4], out[5], out[6] = I0[0, 0], I0[0, 1], I0[1, 1]
        out[7], out[8], out[9] = I0[0, 2], I0[1, 2], I0[2, 2]
        return out


class JointModel:
    """One joint of the tree; mirrors the pinocchio JointModel attributes the path reads."""

    def __init__(self, jtype, axis=None):
        self.jtype = int(jtype)
        self.axis = np.zeros(3) if axis is None else np.array(axis, dtype=np.float64).reshape(3)
        self.id = -1
        self.idx_q = -1
        self.idx_v = -1

    @property
    def nq(self):
        return _JT_NQ[self.jtype]

    @property
    def nv(self):
        return _JT_NV[self.jtype]

    def shortname(self):
        if self.jtype == JT_FREEFLYER:
            return "JointModelFreeFlyer"
        suf = _aligned_suffix(self.axis)
        if suf is None:
            return _SHORTNAMES[self.jtype]
        return {JT_REVOLUTE: "JointModelR", JT_REVOLUTE_UNB: "JointModelRUB", JT_PRISMATIC: "JointModelP"}[self.jtype] + suf


def JointModelFreeFlyer():
    return JointModel(JT_FREEFLYER)


class _InertiaList(list):
    def tolist(self):  # regressor.py:36-39 calls robot.model.inertias.tolist()
        return list(self)


class Model:
    """Kinematic tree with pinocchio's conventions (joint 0 = universe)."""

    def __init__(self, name="robot"):
        self.name = name
        self.names = ["universe"]
        self.parents = [0]
        self.jointPlacements = [SE3()]
        self.joints = [JointModel(JT_NONE)]
        self.joints[0].id = 0
        self.inertias = _InertiaList([Inertia()])
        self.gravity_linear = np.array([0.0, 0.0, -9.81])
        self.nq = 0
        self.nv = 0
        self.lowerPositionLimit = np.zeros(0)
        self.upperPositionLimit = np.zeros(0)

    @property
    def njoints(self):
        return len(self.joints)

    @property
    def nbodies(self):
        return len(self.joints)

    def addJoint(self, parent, jmodel: JointModel, placement: SE3, name, lower=None, upper=None):
        jmodel.id = len(self.joints)
        jmodel.idx_q = self.nq
        jmodel.idx_v = self.nv
        self.nq += jmodel.nq
        self.nv += jmodel.nv
        self.joints.append(jmodel)
        self.parents.append(int(parent))
        self.jointPlacements.append(placement.copy())
        self.names.append(name)
        self.inertias.append(Inertia())
        lo = np.full(jmodel.nq, -np.inf) if lower is None else np.asarray(lower, dtype=np.float64)
        up = np.full(jmodel.nq, np.inf) if upper is None else np.asarray(upper, dtype=np.float64)
        self.lowerPositionLimit = np.concatenate([self.lowerPositionLimit, lo])
        self.upperPositionLimit = np.concatenate([self.upperPositionLimit, up])
        return jmodel.id

    def appendBodyToJoint(self, joint_id, Y: Inertia, placement: SE3):
        self.inertias[joint_id] = self.inertias[joint_id] + placement.act_inertia(Y)

    def getJointId(self, name):
        return self.names.index(name) if name in self.names else len(self.names)

    def existJointName(self, name):
        return name in self.names

    def createData(self):
        return Data(self)


class Data:
    """Placeholder for pinocchio.Data; the B200 path keeps its scratch on the device."""

    def __init__(self, model):
        self.model = model


# --------------------------------------------------------------------------- flatten


@dataclass
class FlatModel:
    """Contiguous arrays consumed by ``fg_model_create`` (include/figaroh_b200.h)."""

    name: str
    njoints: int
    nq: int
    nv: int
    parent: np.ndarray      # int32 [njoints]
    jtype: np.ndarray       # int32 [njoints]
    axis: np.ndarray        # f64 [njoints,3]
    plc_R: np.ndarray       # f64 [njoints,9] row-major rotation of jointPlacement
    plc_p: np.ndarray       # f64 [njoints,3]
    idx_q: np.ndarray       # int32 [njoints]
    idx_v: np.ndarray       # int32 [njoints]
    gravity: np.ndarray     # f64 [3]  (linear part of model.gravity)
    mass: np.ndarray        # f64 [njoints]
    lever: np.ndarray       # f64 [njoints,3]
    inertia: np.ndarray     # f64 [njoints,6]  I_C as xx,xy,yy,xz,yz,zz
    names: list = field(default_factory=list)

    def to_json(self) -> str:
        d = {}
        for k, v in self.__dict__.items():
            d[k] = v.tolist() if isinstance(v, np.ndarray) else v
        # repr() round-trips doubles exactly in Python
        return json.dumps(d)

    @staticmethod
    def from_json(s: str) -> "FlatModel":
        d = json.loads(s)
        i32 = ("parent", "jtype", "idx_q", "idx_v")
        f64 = ("axis", "plc_R", "plc_p", "gravity", "mass", "lever", "inertia")
        for k in i32:
            d[k] = np.ascontiguousarray(d[k], dtype=np.int32)
        for k in f64:
            d[k] = np.ascontiguousarray(d[k], dtype=np.float64)
        return FlatModel(**d)

    def dynamic_parameters(self) -> np.ndarray:
        """(njoints-1, 10) pinocchio-ordered inertial parameters of bodies 1..n."""
        out = np.zeros((self.njoints - 1, 10))
        for i in range(1, self.njoints):
            I = self.inertia[i]
            Ic = np.array([[I[0], I[1], I[3]], [I[1], I[2], I[4]], [I[3], I[4], I[5]]])
            out[i - 1] = Inertia(self.mass[i], self.lever[i], Ic).toDynamicParameters()
        return out


_PIN_SHORTNAME_TO_TYPE = {
    "JointModelRX": (JT_REVOLUTE, (1, 0, 0)), "JointModelRY": (JT_REVOLUTE, (0, 1, 0)),
    "JointModelRZ": (JT_REVOLUTE, (0, 0, 1)), "JointModelRevoluteUnaligned": (JT_REVOLUTE, None),
    "JointModelRUBX": (JT_REVOLUTE_UNB, (1, 0, 0)), "JointModelRUBY": (JT_REVOLUTE_UNB, (0, 1, 0)),
    "JointModelRUBZ": (JT_REVOLUTE_UNB, (0, 0, 1)),
    "JointModelRevoluteUnboundedUnaligned": (JT_REVOLUTE_UNB, None),
    "JointModelPX": (JT_PRISMATIC, (1, 0, 0)), "JointModelPY": (JT_PRISMATIC, (0, 1, 0)),
    "JointModelPZ": (JT_PRISMATIC, (0, 0, 1)), "JointModelPrismaticUnaligned": (JT_PRISMATIC, None),
    "JointModelFreeFlyer": (JT_FREEFLYER, (0, 0, 0)),
}


def _joint_type_axis(j):
    """(type code, axis) from either our JointModel or a pinocchio JointModel."""
    if isinstance(j, JointModel):
        return j.jtype, j.axis
    sn = j.shortname()
    if sn not in _PIN_SHORTNAME_TO_TYPE:
        raise NotImplementedError(f"joint model {sn} is not supported by the B200 regressor path")
    jt, ax = _PIN_SHORTNAME_TO_TYPE[sn]
    if ax is None:  # unaligned variants expose .axis (pinocchio >= 2.6 via extract())
        jj = j.extract() if hasattr(j, "extract") else j
        ax = np.asarray(jj.axis, dtype=np.float64).reshape(3)
    return jt, np.asarray(ax, dtype=np.float64)


def flatten_model(model) -> FlatModel:
    """Flatten a pinocchio-like ``model`` into device-ready arrays.

    Reads only public attributes that pinocchio.Model also offers, so a real
    pinocchio model is a valid argument (reference construction site:
    /root/reference/src/figaroh/tools/robot.py:52-58).
    """
    if isinstance(model, FlatModel):
        return model
    n = int(model.njoints)
    parent = np.zeros(n, np.int32)
    jtype = np.zeros(n, np.int32)
    axis = np.zeros((n, 3))
    plc_R = np.zeros((n, 9))
    plc_p = np.zeros((n, 3))
    idx_q = np.zeros(n, np.int32)
    idx_v = np.zeros(n, np.int32)
    mass = np.zeros(n)
    lever = np.zeros((n, 3))
    inertia = np.zeros((n, 6))
    plc_R[0] = np.eye(3).reshape(9)
    for i in range(n):
        Y = model.inertias[i]
        mass[i] = float(Y.mass)
        lever[i] = np.asarray(Y.lever, dtype=np.float64).reshape(3)
        Ic = np.asarray(Y.inertia, dtype=np.float64).reshape(3, 3)
        inertia[i] = [Ic[0, 0], Ic[0, 1], Ic[1, 1], Ic[0, 2], Ic[1, 2], Ic[2, 2]]
        if i == 0:
            continue
        parent[i] = int(model.parents[i])
        jt, ax = _joint_type_axis(model.joints[i])
        jtype[i] = jt
        axis[i] = ax
        M = model.jointPlacements[i]
        plc_R[i] = np.asarray(M.rotation, dtype=np.float64).reshape(9)
        plc_p[i] = np.asarray(M.translation, dtype=np.float64).reshape(3)
        idx_q[i] = int(model.joints[i].idx_q)
        idx_v[i] = int(model.joints[i].idx_v)
        if parent[i] >= i:
            raise ValueError("model joints must be ordered parents-first")
    if hasattr(model, "gravity_linear"):
        g = np.asarray(model.gravity_linear, dtype=np.float64)
    else:  # pinocchio: Motion with .linear
        g = np.asarray(model.gravity.linear, dtype=np.float64)
    return FlatModel(
        name=str(getattr(model, "name", "robot")), njoints=n, nq=int(model.nq), nv=int(model.nv),
        parent=parent, jtype=jtype, axis=np.ascontiguousarray(axis), plc_R=np.ascontiguousarray(plc_R),
        plc_p=np.ascontiguousarray(plc_p), idx_q=idx_q, idx_v=idx_v, gravity=g.reshape(3).copy(),
        mass=mass, lever=np.ascontiguousarray(lever), inertia=np.ascontiguousarray(inertia),
        names=[str(x) for x in model.names],
    )


def model_from_flat(fm: FlatModel) -> Model:
    """Rebuild a ``Model`` (e.g. from a committed JSON fixture) without any URDF."""
    m = Model(fm.name)
    for i in range(1, fm.njoints):
        jm = JointModel(int(fm.jtype[i]), fm.axis[i])
        m.addJoint(int(fm.parent[i]), jm, SE3(fm.plc_R[i].reshape(3, 3), fm.plc_p[i]), fm.names[i])
        assert jm.idx_q == fm.idx_q[i] and jm.idx_v == fm.idx_v[i]
    for i in range(fm.njoints):
        I = fm.inertia[i]
        Ic = np.array([[I[0], I[1], I[3]], [I[1], I[2], I[4]], [I[3], I[4], I[5]]])
        m.inertias[i] = Inertia(fm.mass[i], fm.lever[i], Ic)
    m.gravity_linear = fm.gravity.copy()
    return m
