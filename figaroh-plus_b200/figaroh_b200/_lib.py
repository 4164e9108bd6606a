"""ctypes binding of lib/libfigaroh_b200.so (C ABI: include/figaroh_b200.h).

There is deliberately no fallback: if the CUDA library is missing or a call fails,
an exception is raised - the product path never routes through CPU code.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FG_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libfigaroh_b200.so")

FG_MODE_JOINT_TORQUES = 0
FG_MODE_EXTERNAL_WRENCH = 1
FG_FLAG_FRICTION = 1
FG_FLAG_ACTUATOR_INERTIA = 2
FG_FLAG_JOINT_OFFSET = 4
FG_FLAG_TX40_COUPLING = 8

SYMBOLS = (
    "fg_last_error", "fg_abi_version", "fg_model_create", "fg_model_destroy", "fg_regressor_shape",
    "fg_regressor_materialize", "fg_state_create", "fg_state_destroy", "fg_state_reset", "fg_state_p",
    "fg_state_rows", "fg_state_accumulate_dense", "fg_state_accumulate_samples", "fg_state_packed_doubles",
    "fg_state_export", "fg_state_merge_packed", "fg_state_colsq", "fg_state_reduce", "fg_state_regroup_solve",
    "fg_state_solve", "fg_state_pivot_qr", "fg_filtfilt", "fg_differentiate", "fg_state_accumulate_samples_host", "fg_state_get_perm", "fg_state_set_perm", "fg_launch_count", "fg_state_set_timing", "fg_state_get_timing", "fg_plan_describe", "fg_colsq_dense", "fg_sosfiltfilt", "fg_filtfilt_decimate", "fg_medfilt", "fg_gradient",
)


class FgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"figaroh_b200 error {code}: {msg}")
        self.code = code


class ModelDesc(C.Structure):
    _fields_ = [("njoints", C.c_int32), ("nq", C.c_int32), ("nv", C.c_int32),
                ("parent", C.c_void_p), ("jtype", C.c_void_p), ("axis", C.c_void_p), ("plc_R", C.c_void_p),
                ("plc_p", C.c_void_p), ("idx_q", C.c_void_p), ("idx_v", C.c_void_p), ("gravity", C.c_void_p),
                ("mass", C.c_void_p)]


class RegressorCfg(C.Structure):
    _fields_ = [("mode", C.c_int32), ("flags", C.c_int32), ("ft_mask", C.c_int32), ("n_row_sel", C.c_int32),
                ("row_sel", C.c_void_p), ("n_col_sel", C.c_int32), ("col_sel", C.c_void_p)]


_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' at the repo root). "
            "figaroh_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dp = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
    L.fg_last_error.restype = C.c_char_p
    L.fg_abi_version.restype = C.c_int
    L.fg_model_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(vp)]
    L.fg_model_destroy.argtypes = [vp]
    L.fg_model_destroy.restype = None
    L.fg_regressor_shape.argtypes = [vp, C.POINTER(RegressorCfg), C.POINTER(i32), C.POINTER(i32)]
    L.fg_regressor_materialize.argtypes = [vp, C.POINTER(RegressorCfg), dp, dp, dp, i64, dp, i64, vp]
    L.fg_state_create.argtypes = [i32, C.POINTER(vp)]
    L.fg_state_destroy.argtypes = [vp]
    L.fg_state_destroy.restype = None
    L.fg_state_reset.argtypes = [vp, vp]
    L.fg_state_p.argtypes = [vp]
    L.fg_state_rows.argtypes = [vp]
    L.fg_state_rows.restype = i64
    L.fg_state_accumulate_dense.argtypes = [vp, dp, i64, dp, i64, vp]
    L.fg_state_accumulate_samples.argtypes = [vp, vp, C.POINTER(RegressorCfg), dp, dp, dp, dp, i64, i64, vp]
    L.fg_state_accumulate_samples_host.argtypes = [vp, vp, C.POINTER(RegressorCfg), dp, dp, dp, dp, i64, i64, vp]
    L.fg_state_packed_doubles.argtypes = [i32]
    L.fg_state_packed_doubles.restype = i64
    L.fg_state_get_perm.argtypes = [vp, vp]
    L.fg_state_set_perm.argtypes = [vp, vp]
    L.fg_state_export.argtypes = [vp, dp, vp]
    L.fg_state_merge_packed.argtypes = [vp, dp, i32, vp]
    L.fg_state_colsq.argtypes = [vp, dp, vp]
    L.fg_state_reduce.argtypes = [vp, vp, i32, dp, vp]
    L.fg_state_regroup_solve.argtypes = [vp, vp, i32, vp, i32, dp, dp, dp, dp, vp]
    L.fg_filtfilt.argtypes = [dp, i64, i32, i64, vp, i32, vp, i32, vp, i32, i64, dp, i64, vp]
    L.fg_differentiate.argtypes = [dp, i64, i32, i64, C.c_double, dp, dp, dp, i64, vp]
    L.fg_state_pivot_qr.argtypes = [vp, vp, i32, vp, dp, vp]
    L.fg_state_solve.argtypes = [vp, C.c_double, vp, i32, C.c_double, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.fg_launch_count.restype = C.c_longlong
    L.fg_state_set_timing.argtypes = [vp, C.c_int]
    L.fg_state_get_timing.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    L.fg_filtfilt_decimate.argtypes = [dp, i64, i32, i64, vp, i32, vp, i32, vp, i32, i32, dp, i64, vp]
    L.fg_sosfiltfilt.argtypes = [dp, i64, i32, i64, vp, i32, vp, i32, i64, i32, dp, i64, vp]
    L.fg_medfilt.argtypes = [dp, i64, i32, i64, i32, dp, i64, vp]
    L.fg_gradient.argtypes = [dp, i64, i32, i64, dp, dp, i64, vp]
    L.fg_colsq_dense.argtypes = [dp, i64, i64, i32, vp, vp]
    L.fg_plan_describe.argtypes = [C.POINTER(ModelDesc), C.POINTER(RegressorCfg), i32, i32, C.c_char_p, i32]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise FgError(rc, lib().fg_last_error().decode())
