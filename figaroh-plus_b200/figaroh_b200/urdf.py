"""URDF -> Model, reproducing pinocchio's ``urdf::buildModel`` semantics.

The reference builds its model with ``RobotWrapper.initFromURDF``
(/root/reference/src/figaroh/tools/robot.py:52-58).  Pinocchio is absent from
this image, so this module restates the parts of that loader which determine
the regressor's column order and values ("Loader
semantics"):

* joints are numbered by depth-first traversal, children visited in
  alphabetical order of *joint* name (urdfdom keeps joints in a ``std::map``);
* ``fixed`` joints create no joint: the child link's inertia is appended to
  the supporting joint through the accumulated fixed transform, and that
  transform is folded into the placement of downstream joints;
* ``rpy`` goes through urdfdom's quaternion (``Rotation::setFromRPY``) and
  back to a matrix, as pinocchio's ``convertFromUrdf`` does;
* the root link's inertia goes to the universe (fixed base) or to a
  free-flyer ``root_joint``.

Only ``<robot>/<link>/<inertial>`` and ``<robot>/<joint>`` elements matter;
meshes, transmissions, gazebo tags and mimic tags are ignored like pinocchio
ignores them for dynamics.
"""
from __future__ import annotations

import math
import xml.etree.ElementTree as ET

import numpy as np

from .model import (JT_FREEFLYER, JT_PRISMATIC, JT_REVOLUTE, JT_REVOLUTE_UNB, Inertia, JointModel, Model, SE3)

_HALF_PI_TOL = 1e-6


def _floats(s, n, default):
    if s is None:
        return list(default)
    vals = [float(x) for x in s.split()]
    if len(vals) != n:
        raise ValueError(f"expected {n} numbers, got {s!r}")
    return vals


def _snap_angle(a):
    """Snap an angle that is a decimal rounding of k*pi/2 to the exact double (opt-in)."""
    k = round(a / (math.pi / 2))
    if k != 0 and abs(a - k * (math.pi / 2)) < _HALF_PI_TOL:
        return k * (math.pi / 2)
    return a


def rpy_to_matrix(r, p, y):
    """urdfdom setFromRPY -> quaternion (normalised) -> Eigen Quaternion::matrix()."""
    phi, the, psi = r / 2.0, p / 2.0, y / 2.0
    qx = math.sin(phi) * math.cos(the) * math.cos(psi) - math.cos(phi) * math.sin(the) * math.sin(psi)
    qy = math.cos(phi) * math.sin(the) * math.cos(psi) + math.sin(phi) * math.cos(the) * math.sin(psi)
    qz = math.cos(phi) * math.cos(the) * math.sin(psi) - math.sin(phi) * math.sin(the) * math.cos(psi)
    qw = math.cos(phi) * math.cos(the) * math.cos(psi) + math.sin(phi) * math.sin(the) * math.sin(psi)
    s = math.sqrt(qx * qx + qy * qy + qz * qz + qw * qw)
    if s == 0.0:
        qx = qy = qz = 0.0
        qw = 1.0
    else:
        qx, qy, qz, qw = qx / s, qy / s, qz / s, qw / s
    return quat_to_matrix(qx, qy, qz, qw)


def quat_to_matrix(x, y, z, w):
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([
        [1.0 - (tyy + tzz), txy - twz, txz + twy],
        [txy + twz, 1.0 - (txx + tzz), tyz - twx],
        [txz - twy, tyz + twx, 1.0 - (txx + tyy)],
    ])


def _origin(elem, snap):
    if elem is None:
        return SE3()
    xyz = _floats(elem.get("xyz"), 3, (0, 0, 0))
    rpy = _floats(elem.get("rpy"), 3, (0, 0, 0))
    if snap:
        rpy = [_snap_angle(a) for a in rpy]
    return SE3(rpy_to_matrix(*rpy), xyz)


def _link_inertia(link_elem, snap):
    ine = link_elem.find("inertial")
    if ine is None:
        return Inertia()
    M = _origin(ine.find("origin"), snap)
    mass_e = ine.find("mass")
    mass = float(mass_e.get("value")) if mass_e is not None else 0.0
    I = np.zeros((3, 3))
    ie = ine.find("inertia")
    if ie is not None:
        ixx, ixy, ixz = float(ie.get("ixx", 0)), float(ie.get("ixy", 0)), float(ie.get("ixz", 0))
        iyy, iyz, izz = float(ie.get("iyy", 0)), float(ie.get("iyz", 0)), float(ie.get("izz", 0))
        I = np.array([[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]])
    R = M.rotation
    return Inertia(mass, M.translation, R @ I @ R.T)


def build_model_from_urdf(path, root_joint=None, snap_right_angles=False) -> Model:
    """Parse ``path`` into a :class:`Model`.

    root_joint: ``None`` for a fixed base, or a free-flyer ``JointModel``
      (``JointModelFreeFlyer()``), mirroring ``initFromURDF(..., root_joint=...)``.
    snap_right_angles: opt-in clean-up that replaces ``rpy`` entries within 1e-6
      of a multiple of pi/2 by the exact double.  Pinocchio does *not* do this;
      it exists for the conditioning study in DESIGN.md.
    """
    tree = ET.parse(path)
    robot = tree.getroot()
    if robot.tag != "robot":
        raise ValueError("not a URDF: root element is <%s>" % robot.tag)
    links = {l.get("name"): l for l in robot.findall("link")}
    joints = {}
    for j in robot.findall("joint"):
        joints[j.get("name")] = j
    children = {name: [] for name in links}
    has_parent = set()
    for jname in sorted(joints):  # std::map order = alphabetical joint names
        j = joints[jname]
        pl = j.find("parent").get("link")
        cl = j.find("child").get("link")
        children[pl].append((jname, cl))
        has_parent.add(cl)
    roots = [n for n in links if n not in has_parent]
    if len(roots) != 1:
        raise ValueError(f"URDF must have exactly one root link, found {roots}")
    root = roots[0]

    model = Model(robot.get("name", "robot"))
    Y_root = _link_inertia(links[root], snap_right_angles)
    if root_joint is None:
        # addFixedJointAndBody(0, Identity, "root_joint", Y, root_link)
        model.appendBodyToJoint(0, Y_root, SE3())
        root_frame = (0, SE3())
    else:
        jid = model.addJoint(0, root_joint, SE3(), "root_joint",
                             lower=np.full(root_joint.nq, -np.inf), upper=np.full(root_joint.nq, np.inf))
        model.appendBodyToJoint(jid, Y_root, SE3())
        root_frame = (jid, SE3())

    # Depth-first traversal, children in alphabetical joint-name order.  A stack item is a
    # link still to be attached: (joint name, child link, supporting joint id, placement of
    # the parent link's frame in that joint).
    def push_children(stack, link, sup, frame_plc):
        for jname, child in reversed(children[link]):
            stack.append((jname, child, sup, frame_plc))

    stack = []
    push_children(stack, root, *root_frame)
    while stack:
        jname, child, sup, frame_plc = stack.pop()
        j = joints[jname]
        jt = j.get("type")
        plc = _origin(j.find("origin"), snap_right_angles)
        Y = _link_inertia(links[child], snap_right_angles)
        ax_e = j.find("axis")
        axis = np.array(_floats(ax_e.get("xyz") if ax_e is not None else None, 3, (1, 0, 0)))
        lim = j.find("limit")
        lo = float(lim.get("lower", 0.0)) if lim is not None else 0.0
        up = float(lim.get("upper", 0.0)) if lim is not None else 0.0
        if jt == "fixed":
            child_plc = frame_plc * plc
            model.appendBodyToJoint(sup, Y, child_plc)
            push_children(stack, child, sup, child_plc)
            continue
        if jt in ("revolute", "continuous", "prismatic"):
            aligned = any(np.allclose(axis, e, rtol=1e-12, atol=0.0) for e in np.eye(3))
            ax = axis if aligned else axis / np.linalg.norm(axis)
            code = {"revolute": JT_REVOLUTE, "continuous": JT_REVOLUTE_UNB, "prismatic": JT_PRISMATIC}[jt]
            jm = JointModel(code, ax)
            if jt == "continuous":
                lower, upper = [-1.01, -1.01], [1.01, 1.01]
            else:
                lower, upper = [lo], [up]
        elif jt == "floating":
            jm = JointModel(JT_FREEFLYER)
            lower, upper = np.full(7, -np.inf), np.full(7, np.inf)
        else:
            raise NotImplementedError(f"URDF joint type {jt!r} ({jname}) is not supported")
        jid = model.addJoint(sup, jm, frame_plc * plc, jname, lower=lower, upper=upper)
        model.appendBodyToJoint(jid, Y, SE3())
        push_children(stack, child, jid, SE3())
    return model
