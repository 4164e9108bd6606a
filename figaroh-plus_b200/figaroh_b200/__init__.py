"""figaroh_b200 - B200-native engine behind the figaroh.tools regressor / QR API.

Layout:
  model.py   pinocchio-like model objects and the flattening loader
  urdf.py    URDF front end with pinocchio's joint ordering
  _lib.py    ctypes binding of the C ABI (include/figaroh_b200.h)
  engine.py  DeviceModel / TsqrState wrappers and the fused identification pipeline
  dist.py    multi-GPU sharding: all-gather of the per-GPU R factors over NCCL
"""
from .model import (FlatModel, Inertia, JointModel, JointModelFreeFlyer, Model, SE3, flatten_model,  # noqa: F401
                    model_from_flat)

__version__ = "0.1.0"
