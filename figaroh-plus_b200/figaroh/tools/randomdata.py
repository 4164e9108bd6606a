"""figaroh.tools.randomdata drop-in (B200): the torque generator of
/root/reference/src/figaroh/tools/randomdata.py:93-147.

get_torque_rand evaluates pin.rnea sample by sample in the reference.  Here the inverse dynamics come from
the regressor identity  rnea(q, v, a) = W_inertial(q, v, a) . phi_inertial  (the identity the whole
identification rests on), evaluated on the GPU in chunks (kernel (1) + one GEMV per chunk), and the
friction / actuator-inertia / offset / coupled-wrist terms are added element-wise on the device exactly as
the reference writes them (including its use of the VELOCITY with Iam6).  The random waypoint generators of
the same file are host-side numpy.random calls and are not re-implemented."""
import numpy as np
import torch

from figaroh_b200 import engine as _eng

from .regressor import _device_model

_PIN2FIG = [9, 6, 7, 8, 0, 1, 3, 2, 4, 5]  # pinocchio [m, mc, Ixx Ixy Iyy Ixz Iyz Izz] -> figaroh slot (robot.py:108-121)


def get_torque_rand(N, robot, q, v, a, param):
    """Joint torques (nv * N,), joint-major (tau[j * N + i]) - reference: randomdata.py:93-147."""
    dm = _device_model(robot)
    fm = dm.flat
    nv = fm.nv
    cfg = _eng.RegressorConfig(_eng._lib.FG_MODE_JOINT_TORQUES, 0)   # inertial columns only
    rps, cols = dm.shape(cfg)
    phi = np.zeros(cols)
    dyn = fm.dynamic_parameters()
    for i in range(fm.njoints - 1):
        blk = np.empty(10)
        blk[_PIN2FIG] = dyn[i]
        phi[14 * i:14 * i + 10] = blk
    qd, vd, ad = _eng.as_device(q), _eng.as_device(v), _eng.as_device(a)
    N = int(N)
    phi_d = torch.from_numpy(phi).to(qd.device)
    tau = torch.empty((nv, N), dtype=torch.float64, device=qd.device)
    chunk = max(1, min(N, (256 << 20) // (8 * rps * cols)))          # <= 256 MB of regressor at a time
    for i0 in range(0, N, chunk):
        n = min(chunk, N - i0)
        W = dm.materialize(cfg, qd[i0:i0 + n], vd[i0:i0 + n], ad[i0:i0 + n])
        tau[:, i0:i0 + n] = (W @ phi_d).view(nv, n)
        del W
    vt, at = vd[:N, :nv].t(), ad[:N, :nv].t()                        # (nv, N) views
    col = lambda key: torch.as_tensor(np.asarray(param[key], dtype=np.float64)[:nv], device=qd.device)[:, None]  # noqa: E731
    if param["has_friction"]:
        tau += vt * col("fv") + torch.sign(vt) * col("fs")
    if param["has_actuator_inertia"]:
        tau += col("Ia") * at
    if param["has_joint_offset"]:
        tau += col("off")
    if param.get("has_coupled_wrist", False):
        s = torch.sign(vt[nv - 2] + vt[nv - 1])
        tau[nv - 2] += param["Iam6"] * vt[nv - 1] + param["fvm6"] * vt[nv - 1] + param["fsm6"] * s
        tau[nv - 1] += param["Iam6"] * vt[nv - 2] + param["fvm6"] * vt[nv - 2] + param["fsm6"] * s
    return tau.reshape(-1).cpu().numpy()
