"""Python face of the C ABI: device model, TSQR state and the fused identification pipeline.

PyTorch is used only to own device / pinned memory and to hand pointers and the
current CUDA stream across the ABI; all arithmetic happens in lib/libfigaroh_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import FgError, ModelDesc, RegressorCfg, check
from .model import FlatModel, flatten_model

TOL_QR = 1e-8
_FT_ROW = {"Fx": 0, "Fy": 1, "Fz": 2, "Mx": 3, "My": 4, "Mz": 5}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("figaroh_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def as_device(x, cols=None) -> torch.Tensor:
    """float64, C-contiguous CUDA tensor from a numpy array / tensor (copied only if needed)."""
    if isinstance(x, DeviceMatrix):
        x = x.t
    if isinstance(x, torch.Tensor):
        t = x
        if t.dtype != torch.float64:
            t = t.to(torch.float64)
        if not t.is_cuda:
            t = t.cuda(non_blocking=True)
        return t.contiguous()
    arr = np.ascontiguousarray(x, dtype=np.float64)
    return torch.from_numpy(arr).cuda(non_blocking=True)


class DeviceMatrix:
    """A regressor that lives in HBM.  Accepted wherever the figaroh.tools functions take W."""

    def __init__(self, t: torch.Tensor):
        assert t.is_cuda and t.dtype == torch.float64 and t.dim() == 2
        self.t = t.contiguous()

    @property
    def shape(self):
        return tuple(self.t.shape)

    def numpy(self):
        return self.t.cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def take_columns(self, idx):
        idx_t = torch.as_tensor(np.asarray(idx, dtype=np.int64), device=self.t.device)
        return DeviceMatrix(self.t.index_select(1, idx_t))

    def drop_columns(self, idx):
        keep = np.setdiff1d(np.arange(self.t.shape[1]), np.asarray(idx, dtype=np.int64))
        return self.take_columns(keep)


class RegressorConfig:
    """Rows/columns of the regressor; built from the reference's `param` dict (regressor.py:45-140)."""

    def __init__(self, mode=_lib.FG_MODE_JOINT_TORQUES, flags=0, ft_mask=0x3F, row_sel=None, col_sel=None):
        self.mode, self.flags, self.ft_mask = int(mode), int(flags), int(ft_mask)
        self.row_sel = None if row_sel is None else np.ascontiguousarray(row_sel, dtype=np.int32)
        self.col_sel = None if col_sel is None else np.ascontiguousarray(col_sel, dtype=np.int32)

    @staticmethod
    def from_param(param, coupling=False, row_sel=None, col_sel=None) -> "RegressorConfig":
        flags = 0
        if param["has_friction"]:
            flags |= _lib.FG_FLAG_FRICTION
        if param["has_actuator_inertia"]:
            flags |= _lib.FG_FLAG_ACTUATOR_INERTIA
        if param["has_joint_offset"]:
            flags |= _lib.FG_FLAG_JOINT_OFFSET
        if coupling:
            flags |= _lib.FG_FLAG_TX40_COUPLING
        if param["is_joint_torques"]:
            return RegressorConfig(_lib.FG_MODE_JOINT_TORQUES, flags, 0x3F, row_sel, col_sel)
        if param["is_external_wrench"]:
            mask = 0
            for ft in param["force_torque"]:
                if ft == "All":
                    mask |= 0x3F
                elif ft in _FT_ROW:
                    mask |= 1 << _FT_ROW[ft]
                else:
                    raise ValueError("Please enter valid parameters")  # regressor.py:140
            return RegressorConfig(_lib.FG_MODE_EXTERNAL_WRENCH, flags, mask, row_sel, col_sel)
        # regressor.py:194 returns an unbound W_mod when neither branch is taken
        raise UnboundLocalError("cannot access local variable 'W_mod' where it is not associated with a value")

    def with_columns(self, col_sel):
        return RegressorConfig(self.mode, self.flags, self.ft_mask, self.row_sel, col_sel)

    def cstruct(self) -> RegressorCfg:
        c = RegressorCfg()
        c.mode, c.flags, c.ft_mask = self.mode, self.flags, self.ft_mask
        c.n_row_sel = 0 if self.row_sel is None else len(self.row_sel)
        c.row_sel = None if self.row_sel is None else self.row_sel.ctypes.data
        c.n_col_sel = 0 if self.col_sel is None else len(self.col_sel)
        c.col_sel = None if self.col_sel is None else self.col_sel.ctypes.data
        return c


def _model_desc(fm: FlatModel):
    """fg_model_desc over the arrays of a FlatModel (returned with the arrays that must stay alive)."""
    d = ModelDesc()
    d.njoints, d.nq, d.nv = fm.njoints, fm.nq, fm.nv
    keep = [np.ascontiguousarray(fm.parent, np.int32), np.ascontiguousarray(fm.jtype, np.int32),
            np.ascontiguousarray(fm.axis), np.ascontiguousarray(fm.plc_R), np.ascontiguousarray(fm.plc_p),
            np.ascontiguousarray(fm.idx_q, np.int32), np.ascontiguousarray(fm.idx_v, np.int32),
            np.ascontiguousarray(fm.gravity), np.ascontiguousarray(fm.mass)]
    (d.parent, d.jtype, d.axis, d.plc_R, d.plc_p, d.idx_q, d.idx_v, d.gravity, d.mass) = [a.ctypes.data for a in keep]
    return d, keep


def describe_plan(model, cfg: "RegressorConfig", sm_count=148, max_smem=227 * 1024) -> dict:
    """Host-only: the fused path's symbolic pass for (model, cfg) - walk program, row-group problems, warp split."""
    import json
    d, keep = _model_desc(flatten_model(model))
    cs = cfg.cstruct()
    buf = C.create_string_buffer(1 << 16)
    check(_lib.lib().fg_plan_describe(C.byref(d), C.byref(cs), sm_count, max_smem, buf, len(buf)))
    return json.loads(buf.value.decode())


class DeviceModel:
    """A flattened kinematic tree resident on the current CUDA device (fg_model)."""

    def __init__(self, model):
        _require_cuda()
        self.flat: FlatModel = flatten_model(model)
        d, self._keep = _model_desc(self.flat)
        self._h = C.c_void_p()
        check(_lib.lib().fg_model_create(C.byref(d), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and _lib is not None and _lib._lib is not None:
            _lib._lib.fg_model_destroy(h)
            self._h = None

    @property
    def nq(self):
        return self.flat.nq

    @property
    def nv(self):
        return self.flat.nv

    def shape(self, cfg: RegressorConfig):
        rps, cols = C.c_int32(), C.c_int32()
        cs = cfg.cstruct()
        check(_lib.lib().fg_regressor_shape(self._h, C.byref(cs), C.byref(rps), C.byref(cols)))
        return rps.value, cols.value

    def materialize(self, cfg: RegressorConfig, q, v, a, out: torch.Tensor | None = None) -> torch.Tensor:
        """Kernel (1): tall regressor in the reference layout, returned as a CUDA tensor."""
        qd, vd, ad = as_device(q), as_device(v), as_device(a)
        N = qd.shape[0]
        if qd.dim() != 2 or qd.shape[1] != self.nq or tuple(vd.shape) != (N, self.nv) or tuple(ad.shape) != (N, self.nv):
            raise ValueError(f"expected q (N,{self.nq}), v (N,{self.nv}), a (N,{self.nv})")
        full = RegressorConfig(cfg.mode, cfg.flags, cfg.ft_mask)
        rps, cols = self.shape(full)
        if out is None:
            out = torch.empty((rps * N, cols), dtype=torch.float64, device=qd.device)
        cs = full.cstruct()
        check(_lib.lib().fg_regressor_materialize(self._h, C.byref(cs), qd.data_ptr(), vd.data_ptr(), ad.data_ptr(),
                                                  N, out.data_ptr(), out.stride(0), _stream()))
        return out


class TsqrState:
    """(p+1)x(p+1) Householder R factor of [W | tau] + exact column norms (fg_state)."""

    def __init__(self, p: int):
        _require_cuda()
        self._h = C.c_void_p()
        check(_lib.lib().fg_state_create(int(p), C.byref(self._h)))
        self.p = int(p)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value and _lib is not None and _lib._lib is not None:
            _lib._lib.fg_state_destroy(h)
            self._h = None

    @property
    def rows(self):
        return int(_lib.lib().fg_state_rows(self._h))

    def reset(self):
        check(_lib.lib().fg_state_reset(self._h, _stream()))

    def accumulate_dense(self, W, tau=None):
        Wd = as_device(W)
        if Wd.dim() != 2 or Wd.shape[1] != self.p:
            raise ValueError(f"W must have {self.p} columns")
        td = None
        if tau is not None:
            td = as_device(tau).reshape(-1)
            if td.shape[0] != Wd.shape[0]:
                raise ValueError("tau must have one entry per row of W")
        check(_lib.lib().fg_state_accumulate_dense(self._h, Wd.data_ptr(), Wd.stride(0),
                                                   None if td is None else td.data_ptr(), Wd.shape[0], _stream()))
        return self

    def accumulate_samples(self, dm: DeviceModel, cfg: RegressorConfig, q, v, a, tau=None):
        """Fused kernel (1)+(2) on device-resident inputs.  tau: (rows_per_sample, N) joint-major or flat."""
        qd, vd, ad = as_device(q), as_device(v), as_device(a)
        N = qd.shape[0]
        td, ld = None, 0
        if tau is not None:
            td = as_device(tau)
            rps, _ = dm.shape(cfg)
            if td.numel() != rps * N:
                raise ValueError(f"tau must hold {rps}*{N} entries (joint-major)")
            ld = N
        cs = cfg.cstruct()
        check(_lib.lib().fg_state_accumulate_samples(self._h, dm._h, C.byref(cs), qd.data_ptr(), vd.data_ptr(),
                                                     ad.data_ptr(), None if td is None else td.data_ptr(), ld, N,
                                                     _stream()))
        return self

    def accumulate_samples_host(self, dm: DeviceModel, cfg: RegressorConfig, q, v, a, tau=None):
        """Same, from HOST arrays (numpy or CPU tensors); H2D staging happens inside the call."""
        def host(x):
            if isinstance(x, torch.Tensor):
                return x if (x.dtype == torch.float64 and x.is_contiguous() and not x.is_cuda) else x.double().cpu().contiguous()
            return np.ascontiguousarray(x, dtype=np.float64)

        def ptr(x):
            return x.data_ptr() if isinstance(x, torch.Tensor) else x.ctypes.data
        qh, vh, ah = host(q), host(v), host(a)
        N = qh.shape[0]
        th = None if tau is None else host(tau)
        cs = cfg.cstruct()
        check(_lib.lib().fg_state_accumulate_samples_host(self._h, dm._h, C.byref(cs), ptr(qh), ptr(vh), ptr(ah),
                                                          None if th is None else ptr(th), N, N, _stream()))
        self._host_keep = (qh, vh, ah, th)  # keep alive until the stream is synchronised
        return self

    # ---- diagnostics
    def set_timing(self, on=True):
        check(_lib.lib().fg_state_set_timing(self._h, 1 if on else 0))

    def get_timing(self):
        """(total ms, launches) of the fold kernel since set_timing(True)."""
        ms, n = C.c_double(), C.c_longlong()
        check(_lib.lib().fg_state_get_timing(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # ---- multi-GPU exchange
    def packed_doubles(self):
        return int(_lib.lib().fg_state_packed_doubles(self.p))

    def get_perm(self) -> np.ndarray:
        """State column i holds caller column perm[i] (the fused kernel picks the order)."""
        perm = np.empty(self.p, dtype=np.int32)
        check(_lib.lib().fg_state_get_perm(self._h, perm.ctypes.data))
        return perm

    def set_perm(self, perm):
        perm = np.ascontiguousarray(perm, dtype=np.int32)
        assert perm.shape == (self.p,)
        check(_lib.lib().fg_state_set_perm(self._h, perm.ctypes.data))
        return self

    def export(self, out: torch.Tensor | None = None) -> torch.Tensor:
        if out is None:
            out = torch.empty(self.packed_doubles(), dtype=torch.float64, device="cuda")
        check(_lib.lib().fg_state_export(self._h, out.data_ptr(), _stream()))
        return out

    def merge_packed(self, packed: torch.Tensor, n: int):
        assert packed.is_cuda and packed.dtype == torch.float64 and packed.is_contiguous()
        assert packed.numel() == n * self.packed_doubles()
        check(_lib.lib().fg_state_merge_packed(self._h, packed.data_ptr(), int(n), _stream()))
        return self

    # ---- small stage
    def colsq(self) -> np.ndarray:
        out = np.empty(self.p + 1)
        check(_lib.lib().fg_state_colsq(self._h, out.ctypes.data, _stream()))
        return out

    def reduce(self, keep) -> np.ndarray:
        keep = np.ascontiguousarray(keep, dtype=np.int32)
        diag = np.empty(len(keep) + 1)
        check(_lib.lib().fg_state_reduce(self._h, keep.ctypes.data, len(keep), diag.ctypes.data, _stream()))
        return diag

    def solve(self, tol_e, drop, tol_qr):
        """The whole N-independent stage in one launch (fg_state_solve).  Returns None when the library
        cannot take this state (permuted factor / too large): the caller uses the three-call sequence."""
        p = self.p
        drop = np.ascontiguousarray(sorted(drop), dtype=np.int32)
        keep = np.empty(p, dtype=np.int32)
        order = np.empty(p, dtype=np.int32)
        colsq = np.empty(p + 1)
        diag = np.empty(p + 1)
        beta = np.empty(p * p)
        phi = np.empty(p)
        R1 = np.empty(p * p)
        nk, nb, r2 = C.c_int32(), C.c_int32(), C.c_double()
        rc = _lib.lib().fg_state_solve(self._h, -1.0 if tol_e is None else float(tol_e),
                                       drop.ctypes.data if len(drop) else None, len(drop), float(tol_qr),
                                       C.addressof(nk), keep.ctypes.data, colsq.ctypes.data, diag.ctypes.data,
                                       C.addressof(nb), order.ctypes.data, beta.ctypes.data, phi.ctypes.data,
                                       R1.ctypes.data, C.addressof(r2), _stream())
        if rc == -2:  # FG_ERR_UNSUPPORTED
            return None
        check(rc)
        nk, nb = nk.value, nb.value
        ng = nk - nb
        return dict(keep=[int(x) for x in keep[:nk]], colsq=colsq, diag=diag[:nk + 1].copy(),
                    base=[int(x) for x in order[:nb]], regroup=[int(x) for x in order[nb:nk]],
                    beta=beta[:nb * ng].reshape(nb, ng).copy(), phi=phi[:nb].copy(),
                    R1=R1[:nb * nb].reshape(nb, nb).copy(), resid2=r2.value)

    def pivot_qr(self, keep):
        """Column-pivoted QR of the kept columns, rhs carried along: (perm, R) with R (nk+1) x (nk+1)."""
        keep = np.ascontiguousarray(keep, dtype=np.int32)
        nk = len(keep)
        perm = np.empty(nk, dtype=np.int32)
        R = np.empty((nk + 1, nk + 1))
        check(_lib.lib().fg_state_pivot_qr(self._h, keep.ctypes.data, nk, perm.ctypes.data, R.ctypes.data, _stream()))
        return perm, R

    def regroup_solve(self, base, regroup):
        base = np.ascontiguousarray(base, dtype=np.int32)
        regroup = np.ascontiguousarray(regroup, dtype=np.int32)
        nb, ng = len(base), len(regroup)
        beta = np.empty((nb, ng))
        phi = np.empty(nb)
        R1 = np.empty((nb, nb))
        r2 = C.c_double()
        check(_lib.lib().fg_state_regroup_solve(self._h, base.ctypes.data, nb, regroup.ctypes.data if ng else None, ng,
                                                beta.ctypes.data if ng else None, phi.ctypes.data, R1.ctypes.data,
                                                C.byref(r2), _stream()))
        return beta, phi, R1, r2.value


# --------------------------------------------------------------------------- reference semantics on a state


def launch_count() -> int:
    return int(_lib.lib().fg_launch_count())


def expressions(names_base, names_regroup, beta):
    """'name + |b|*other' strings for |b| >= 1e-6, formatted like qrdecomposition.py:160-183."""
    out = list(names_base)
    for i in range(len(out)):
        for j in range(beta.shape[1]):
            b = beta[i, j]
            if abs(b) < 1e-6:
                continue
            out[i] = out[i] + (" - " if b < -1e-6 else " + ") + str(abs(b)) + "*" + str(names_regroup[j])
    return out


class BaseSolution:
    """Everything get_index_eliminate + double_QR return, computed from a TsqrState."""

    def __init__(self):
        self.idx_e = []          # eliminated columns (indices into the state's columns)
        self.keep = []           # kept columns
        self.params_r = []       # names of kept columns
        self.diag = None         # |diag R| of the reduced factor (last entry: residual norm vs all kept columns)
        self.idx_base = []       # base columns, indices into the reduced (kept) columns
        self.idx_regroup = []
        self.beta_raw = None     # unrounded R1^-1 R2
        self.beta = None         # np.around(., 6)
        self.phi_b_raw = None
        self.phi_b = None        # np.round(., 6)
        self.R1 = None
        self.resid2 = None       # |tau - W_b phi_b|^2
        self._params_base = None  # expression strings, built on first access (string work stays off the hot loop)
        self.phi_std = None
        self._rows = None        # rows of the regressor (fetched lazily: after a multi-GPU merge it lives on the device)
        self._state = None

    @property
    def params_base(self):
        if self._params_base is None:
            nb_ = [self.params_r[i] for i in self.idx_base]
            ng_ = [self.params_r[i] for i in self.idx_regroup]
            self._params_base = expressions(nb_, ng_, self.beta)
        return self._params_base

    @params_base.setter
    def params_base(self, value):
        self._params_base = value

    @property
    def rows(self):
        if self._rows is None and self._state is not None:
            self._rows = self._state.rows
        return self._rows or 0

    @rows.setter
    def rows(self, value):
        self._rows = value


def structural_columns(dm: DeviceModel, cfg: RegressorConfig, nprobe=64, seed=20241017) -> np.ndarray:
    """Columns of the regressor that are not identically zero for this model / configuration.

    The N-independent "symbolic" pass: W is evaluated on `nprobe` random generic samples and the
    columns whose every entry is exactly 0.0 are dropped before the QR.  Their squared norm is 0 for
    any data, so get_index_eliminate (regressor.py:271-274) eliminates them for every tol_e > 0;
    skipping them changes neither idx_e nor anything downstream.  Cached on the DeviceModel.
    """
    key = (cfg.mode, cfg.flags, cfg.ft_mask, None if cfg.row_sel is None else tuple(int(i) for i in cfg.row_sel))
    cache = dm.__dict__.setdefault("_structural", {})
    if key in cache:
        return cache[key]
    fm = dm.flat
    rng = np.random.default_rng(seed)
    q = rng.uniform(-np.pi, np.pi, size=(nprobe, fm.nq))
    v = rng.uniform(-2.0, 2.0, size=(nprobe, fm.nv))
    a = rng.uniform(-3.0, 3.0, size=(nprobe, fm.nv))
    for i in range(1, fm.njoints):
        iq = fm.idx_q[i]
        if fm.jtype[i] == 2:
            th = rng.uniform(-np.pi, np.pi, nprobe)
            q[:, iq], q[:, iq + 1] = np.cos(th), np.sin(th)
        elif fm.jtype[i] == 4:
            quat = rng.normal(size=(nprobe, 4))
            q[:, iq + 3:iq + 7] = quat / np.linalg.norm(quat, axis=1, keepdims=True)
    W = dm.materialize(cfg, q, v, a)
    rps = W.shape[0] // nprobe
    amax = W.abs().view(rps, nprobe, -1).amax(dim=1)          # (row block, column)
    if cfg.mode == _lib.FG_MODE_JOINT_TORQUES and cfg.row_sel is not None:
        amax = amax[torch.as_tensor(np.asarray(cfg.row_sel, dtype=np.int64), device=amax.device)]
    nz = (amax.amax(dim=0) != 0).cpu().numpy()
    sel = np.flatnonzero(nz).astype(np.int32)
    cache[key] = sel
    return sel


def solve_state(state: TsqrState, names, tol_e=None, idx_e=None, tol_qr=TOL_QR, params_std=None,
                col_map=None, p_full=None) -> BaseSolution:
    """Reference post-processing (regressor.py:258-293 + qrdecomposition.py:89-187) on the R factor.

    names: names of the regressor's columns (list(params_std.keys())[:p] in the reference).
    Either tol_e (squared-norm threshold, regressor.py:271-274) or an explicit idx_e.
    col_map / p_full: when the state holds only a subset of the regressor's p_full columns
    (structural_columns), col_map[i] is the regressor column of state column i; idx_e / keep /
    names are then expressed in regressor columns, the dropped ones counting as eliminated.
    """
    sol = BaseSolution()
    p = state.p
    sol._state = state
    if col_map is None:
        col_map = np.arange(p)
        p_full = p
    col_map = np.asarray(col_map, dtype=np.int64)
    if idx_e is None:
        fused = state.solve(tol_e, (), tol_qr)
    else:
        inv = {int(c): i for i, c in enumerate(col_map)}
        fused = state.solve(None, set(inv[int(c)] for c in idx_e if int(c) in inv), tol_qr)
    if fused is not None:  # one launch, one device-to-host copy
        keep_state = fused["keep"]
        sol.keep = [int(col_map[i]) for i in keep_state]
        kept = set(sol.keep)
        sol.idx_e = [i for i in range(int(p_full)) if i not in kept]
        sol.params_r = [names[i] for i in sol.keep]
        sol.diag = fused["diag"]
        sol.idx_base, sol.idx_regroup = fused["base"], fused["regroup"]
        if not sol.idx_base:
            raise FgError(-1, "no column of the regressor exceeds tol_qr")
        beta, phi, R1, r2 = fused["beta"], fused["phi"], fused["R1"], fused["resid2"]
    else:
        if idx_e is None:
            sq = state.colsq()[:p]
            drop = set(i for i in range(p) if sq[i] < tol_e) if tol_e is not None else set()
        else:
            drop = set(inv[int(c)] for c in idx_e if int(c) in inv)
        keep_state = [i for i in range(p) if i not in drop]
        sol.keep = [int(col_map[i]) for i in keep_state]
        kept = set(sol.keep)
        sol.idx_e = [i for i in range(int(p_full)) if i not in kept]
        sol.params_r = [names[i] for i in sol.keep]
        diag = state.reduce(keep_state)
        sol.diag = diag
        nk = len(sol.keep)
        sol.idx_base = [i for i in range(nk) if diag[i] > tol_qr]
        sol.idx_regroup = [i for i in range(nk) if not diag[i] > tol_qr]
        if not sol.idx_base:
            raise FgError(-1, "no column of the regressor exceeds tol_qr")
        beta, phi, R1, r2 = state.regroup_solve(sol.idx_base, sol.idx_regroup)
    sol.beta_raw, sol.phi_b_raw, sol.R1, sol.resid2 = beta, phi, R1, r2
    sol.beta = np.around(beta, 6)
    sol.phi_b = np.round(phi, 6)
    nb_ = [sol.params_r[i] for i in sol.idx_base]
    ng_ = [sol.params_r[i] for i in sol.idx_regroup]
    if params_std is not None:
        ps = [params_std[x] for x in nb_]
        for i in range(len(nb_)):
            for j in range(sol.beta.shape[1]):
                ps[i] = ps[i] + sol.beta[i, j] * params_std[ng_[j]]
        sol.phi_std = np.around(ps, 5)
    return sol


def relative_stdev_from_solution(sol: BaseSolution) -> np.ndarray:
    """identification_tools.py:204-234 evaluated on the factor: (W_b^T W_b)^-1 = R1^-1 R1^-T.

    Uses the rounded phi_b like the reference (its residual is taken with the rounded values)."""
    r = len(sol.idx_base)
    d = sol.phi_b - sol.phi_b_raw
    resid2 = sol.resid2 + float(np.sum((sol.R1 @ d) ** 2))
    sig2 = resid2 / (sol.rows - r)
    Rinv = np.linalg.inv(sol.R1)
    C_x = sig2 * (Rinv @ Rinv.T)
    return np.round(100 * np.sqrt(np.diag(C_x)) / np.abs(sol.phi_b), 2)


def identify(dm: DeviceModel, cfg: RegressorConfig, q, v, a, tau, names, tol_e=1e-6, tol_qr=TOL_QR,
             params_std=None, host_inputs=False, state: TsqrState | None = None, prune=True) -> BaseSolution:
    """Samples -> base parameters in one fused pass: the whole hot path of the reference
    (build_regressor_basic -> get_index_eliminate -> build_regressor_reduced -> double_QR)
    without ever forming W.  Columns that are identically zero for the model (structural_columns)
    are not generated at all unless prune=False."""
    _, cols = dm.shape(cfg)
    col_map = None
    # identically-zero columns are eliminated by get_index_eliminate only for tol_e > 0 (regressor.py:271-274);
    # with tol_e None / <= 0 the reference keeps them (they end up regrouped), so they are generated here too
    prune = prune and tol_e is not None and tol_e > 0
    if cfg.col_sel is None and prune:
        col_map = structural_columns(dm, cfg)
        cfg = cfg.with_columns(col_map)
    elif cfg.col_sel is not None:
        col_map = np.asarray(cfg.col_sel)
    p = cols if col_map is None else len(col_map)
    st = state if state is not None else TsqrState(p)
    if host_inputs:
        st.accumulate_samples_host(dm, cfg, q, v, a, tau)
    else:
        st.accumulate_samples(dm, cfg, q, v, a, tau)
    return solve_state(st, names, tol_e=tol_e, tol_qr=tol_qr, params_std=params_std, col_map=col_map, p_full=cols)
