"""figaroh.tools.robot drop-in: Robot wrapper + standard-parameter naming.

Mirrors /root/reference/src/figaroh/tools/robot.py (Robot.__init__ :26-74,
get_standard_parameters :76-155).  When pinocchio is installed its RobotWrapper is used
unchanged; otherwise the URDF is read by figaroh_b200.urdf, which reproduces pinocchio's
joint ordering and inertia merging.  Either way `robot.model` is accepted by
figaroh.tools.regressor.
"""
import numpy as np

try:  # pragma: no cover - pinocchio is not part of this image
    import pinocchio as _pin
    from pinocchio.robot_wrapper import RobotWrapper as _Base
    _HAVE_PIN = True
except Exception:  # noqa: BLE001
    _pin = None
    _HAVE_PIN = False

    class _Base:  # minimal RobotWrapper stand-in
        def initFromURDF(self, filename, package_dirs=None, root_joint=None, verbose=False, meshLoader=None):
            from figaroh_b200.urdf import build_model_from_urdf
            self.model = build_model_from_urdf(filename, root_joint=root_joint)
            self.data = self.model.createData()
            self.collision_model = None
            self.visual_model = None
            self.q0 = _neutral(self.model)
            self.v0 = np.zeros(self.model.nv)

from figaroh_b200.model import JointModelFreeFlyer, flatten_model

_INERTIAL_NAMES = ("Ixx", "Ixy", "Ixz", "Iyy", "Iyz", "Izz", "mx", "my", "mz", "m")
# slot of pinocchio's [m, mx, my, mz, Ixx, Ixy, Iyy, Ixz, Iyz, Izz] in the FIGAROH order above
_PIN_TO_FIG = (9, 6, 7, 8, 0, 1, 3, 2, 4, 5)


def _neutral(model):
    fm = flatten_model(model)
    q = np.zeros(fm.nq)
    for i in range(1, fm.njoints):
        if fm.jtype[i] == 2:
            q[fm.idx_q[i]] = 1.0
        elif fm.jtype[i] == 4:
            q[fm.idx_q[i] + 6] = 1.0
    return q


class Robot(_Base):
    """Robot model loaded from a URDF, optionally with a free-flyer root joint."""

    def __init__(self, robot_urdf, package_dirs, isFext=False, freeflyer_ori=None):
        self.params_name = _INERTIAL_NAMES
        self.isFext = isFext
        self.robot_urdf = robot_urdf
        if not isFext:
            self.initFromURDF(robot_urdf, package_dirs=package_dirs)
        else:
            ff = _pin.JointModelFreeFlyer() if _HAVE_PIN else JointModelFreeFlyer()
            self.initFromURDF(robot_urdf, package_dirs=package_dirs, root_joint=ff)
        if freeflyer_ori is not None and isFext:
            jid = self.model.getJointId("root_joint")
            self.model.jointPlacements[jid].rotation = freeflyer_ori
            ub = np.array(self.model.upperPositionLimit, dtype=float)
            lb = np.array(self.model.lowerPositionLimit, dtype=float)
            ub[:7] = 1
            lb[:7] = -1
            self.model.upperPositionLimit = ub
            self.model.lowerPositionLimit = lb
            self.data = self.model.createData()
            if hasattr(self.model, "_fg_device_model"):
                del self.model._fg_device_model
        self.geom_model = self.collision_model

    def get_standard_parameters(self, param):
        """Ordered dict name -> value; its key order IS the regressor's column order."""
        model = self.model
        out = {}
        for i in range(1, len(model.inertias)):
            dyn = np.asarray(model.inertias[i].toDynamicParameters(), dtype=float)
            block = np.empty(10)
            for k_pin, k_fig in enumerate(_PIN_TO_FIG):
                block[k_fig] = dyn[k_pin]
            for nm, val in zip(_INERTIAL_NAMES, block):
                out[nm + str(i)] = val
            extras = (("Ia", "has_actuator_inertia", ("Ia",)), ("fv", "has_friction", ("fv", "fs")),
                      ("off", "has_joint_offset", ("off",)))
            vals = {}
            for label, flag, keys in extras:
                got = [0] * len(keys)
                if param[flag]:
                    try:
                        got = [param[k][i - 1] for k in keys]
                    except Exception as e:  # noqa: BLE001 - same tolerance as the reference
                        print("Warning: ", "%s_%d" % (flag, i), e)
                        got = [0] * len(keys)
                vals[label] = got
            out["Ia" + str(i)] = vals["Ia"][0]
            out["fv" + str(i)] = vals["fv"][0]
            out["fs" + str(i)] = vals["fv"][1]
            out["off" + str(i)] = vals["off"][0]
        return out
