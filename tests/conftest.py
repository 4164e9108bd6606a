import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "figaroh-plus_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_path(*parts):
    return os.path.join(GOLDEN, *parts)


def load_flat(name):
    from figaroh_b200.model import FlatModel
    with open(golden_path("models", name + ".json")) as f:
        return FlatModel.from_json(f.read())


def load_oracle_model(name):
    from oracle import kernels
    return kernels.OracleModel(golden_path("models", name + ".json"))


def rand_inputs(fm, N, seed, lo_hi=((-6, 6), (-6, 6), (-6, 6))):
    """Same generator as tests/golden/make_golden.py::rand_inputs (must stay in sync)."""
    rng = np.random.default_rng(seed)
    q = rng.uniform(*lo_hi[0], size=(N, fm.nq))
    v = rng.uniform(*lo_hi[1], size=(N, fm.nv))
    a = rng.uniform(*lo_hi[2], size=(N, fm.nv))
    for i in range(1, fm.njoints):
        iq = fm.idx_q[i]
        if fm.jtype[i] == 2:
            th = rng.uniform(-np.pi, np.pi, N)
            q[:, iq], q[:, iq + 1] = np.cos(th), np.sin(th)
        elif fm.jtype[i] == 4:
            quat = rng.normal(size=(N, 4))
            q[:, iq + 3:iq + 7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    return q, v, a


TX40_PARAM = dict(
    is_joint_torques=True, is_external_wrench=False, force_torque=None,
    has_friction=True, has_actuator_inertia=True, has_joint_offset=True, has_coupled_wrist=True,
    fv=[8.05e0, 5.53e0, 1.97e0, 1.11e0, 1.86e0, 6.5e-1], fs=[7.14e0, 8.26e0, 6.34e0, 2.48e0, 3.03e0, 2.82e-1],
    Ia=[3.62e-1, 3.62e-1, 9.88e-2, 3.13e-2, 4.68e-2, 1.05e-2], off=[3.92e-1, 1.37, 3.26e-1, -1.02e-1, -2.88e-2, 1.27e-1],
    Iam6=9.64e-3, fvm6=6.16e-1, fsm6=1.95e0)


def make_param(fr, ia, jo, **kw):
    d = dict(is_joint_torques=True, is_external_wrench=False, force_torque=None,
             has_friction=fr, has_actuator_inertia=ia, has_joint_offset=jo, fv=None, fs=None, Ia=None, off=None)
    d.update(kw)
    return d


@pytest.fixture(scope="session")
def structural_cases():
    with open(golden_path("structural.json")) as f:
        return {c["name"]: c for c in json.load(f)}
