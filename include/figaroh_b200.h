/*
 * figaroh_b200.h - C ABI of the B200-native FIGAROH dynamic-identification path.
 *
 * The reference (thanhndv212/figaroh-plus) has no FFI of its own: its boundary is the
 * Python module API of figaroh.tools.regressor / figaroh.tools.qrdecomposition, which
 * reaches native code only through pinocchio and LAPACK.  Each entry point below names
 * the reference interface it replaces (paths relative to /root/reference).  The Python
 * drop-in (figaroh-plus_b200/figaroh/tools/*.py) binds these symbols with ctypes;
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *  - every function returns 0 on success or a negative fg_status; fg_last_error()
 *    returns a thread-local message for the last failure; no C++ exceptions cross the ABI;
 *  - "dev" pointers are CUDA device pointers on the current device, "host" pointers are
 *    ordinary host memory; all arrays are IEEE fp64 / int32 / int64, row-major;
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream); device work is
 *    enqueued on it and, unless stated otherwise, not synchronised;
 *  - objects are not thread-safe; use one fg_state per thread.  No hidden global state.
 */
#ifndef FIGAROH_B200_H
#define FIGAROH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FG_ABI_VERSION 1

typedef enum {
  FG_OK = 0,
  FG_ERR_INVALID = -1,     /* bad argument */
  FG_ERR_UNSUPPORTED = -2, /* model / mode outside what the kernels implement */
  FG_ERR_CUDA = -3,        /* CUDA runtime error (message has the cudaError string) */
  FG_ERR_NOMEM = -4
} fg_status;

/* joint type codes (figaroh_b200/model.py) */
enum { FG_JT_NONE = 0, FG_JT_REVOLUTE = 1, FG_JT_REVOLUTE_UNB = 2, FG_JT_PRISMATIC = 3, FG_JT_FREEFLYER = 4 };

/* regressor row modes: regressor.py:45 (is_joint_torques) / :89 (is_external_wrench) */
enum { FG_MODE_JOINT_TORQUES = 0, FG_MODE_EXTERNAL_WRENCH = 1 };

/* flag bits: param["has_friction"|"has_actuator_inertia"|"has_joint_offset"] (regressor.py:55,62,67)
 * and the three TX40 wrist-coupling columns of add_coupling_TX40 (regressor.py:198-227). */
enum { FG_FLAG_FRICTION = 1, FG_FLAG_ACTUATOR_INERTIA = 2, FG_FLAG_JOINT_OFFSET = 4, FG_FLAG_TX40_COUPLING = 8 };

/* Flattened pinocchio::Model - host arrays, copied to the device by fg_model_create.
 * Replaces the `robot.model` argument of build_regressor_basic (regressor.py:20,49). */
typedef struct {
  int32_t njoints, nq, nv;      /* njoints includes the universe (joint 0) */
  const int32_t* parent;        /* [njoints] */
  const int32_t* jtype;         /* [njoints] FG_JT_* */
  const double* axis;           /* [njoints][3] unit axis (revolute / prismatic) */
  const double* plc_R;          /* [njoints][9] jointPlacements[i].rotation, row-major */
  const double* plc_p;          /* [njoints][3] jointPlacements[i].translation */
  const int32_t* idx_q;         /* [njoints] */
  const int32_t* idx_v;         /* [njoints] */
  const double* gravity;        /* [3] linear part of model.gravity */
  const double* mass;           /* [njoints] body masses (external-wrench mode copies only mass != 0 bodies) */
} fg_model_desc;

typedef struct fg_model fg_model;
typedef struct fg_state fg_state;

/* Which rows / columns the regressor has.  Mirrors the `param` dict keys read at
 * regressor.py:45-140 plus the TIAGo script's active-joint row selection
 * (examples/tiago/identification.py:157-166). */
typedef struct {
  int32_t mode;        /* FG_MODE_* */
  int32_t flags;       /* FG_FLAG_* */
  int32_t ft_mask;     /* external wrench: bit j enables row block j of [Fx,Fy,Fz,Mx,My,Mz] (regressor.py:96-140) */
  int32_t n_row_sel;   /* 0 = all row blocks; else number of entries in row_sel (fused path only) */
  const int32_t* row_sel; /* host [n_row_sel]: velocity indices (JT) whose row blocks are used, ascending */
  int32_t n_col_sel;   /* 0 = all columns; else number of entries in col_sel (fused path only) */
  const int32_t* col_sel; /* host [n_col_sel]: regressor columns that are generated and folded, ascending;
                             the state's p is then n_col_sel (columns known to be structurally zero are skipped) */
} fg_regressor_cfg;

const char* fg_last_error(void);
int fg_abi_version(void);

/* ---- model ------------------------------------------------------------------------ */
int fg_model_create(const fg_model_desc* desc, fg_model** out);
void fg_model_destroy(fg_model* m);

/* rows per sample (nv in JT mode, 6 in EW mode) and total column count 14*(njoints-1)
 * (+3 with FG_FLAG_TX40_COUPLING) of the regressor described by cfg. */
int fg_regressor_shape(const fg_model* m, const fg_regressor_cfg* cfg, int32_t* rows_per_sample, int32_t* cols);

/* ---- kernel (1): batched joint-torque regressor, W materialised --------------------
 * Replaces build_regressor_basic (+ add_coupling_TX40 when FG_FLAG_TX40_COUPLING):
 * regressor.py:20-194, :198-227.  q (N,nq), v (N,nv), a (N,nv) dev, row-major.
 * W dev, (rows_per_sample*N) x cols with leading dimension ldw >= cols, reference
 * layout: row j*N+i holds row block j of sample i; FIGAROH per-body column order
 * [Ixx Ixy Ixz Iyy Iyz Izz mx my mz m Ia fv fs off].  Every entry of W is written. */
int fg_regressor_materialize(const fg_model* m, const fg_regressor_cfg* cfg, const double* q_dev,
                             const double* v_dev, const double* a_dev, int64_t N, double* W_dev, int64_t ldw,
                             void* stream);

/* ---- kernel (2): TSQR state ---------------------------------------------------------
 * A state summarises a tall matrix [W | tau] with p regressor columns by the
 * (p+1)x(p+1) upper-triangular Householder R factor of [W | tau], the exact squared
 * column norms and the row count.  It replaces the tall np.linalg.qr calls of
 * qrdecomposition.py:107,140,205,238,285 and diag(W^T W) of regressor.py:243,271:
 * R^T R = [W|tau]^T [W|tau], so every later step of the reference is a function of it. */
int fg_state_create(int32_t p, fg_state** out);
void fg_state_destroy(fg_state* s);
int fg_state_reset(fg_state* s, void* stream);
int fg_state_p(const fg_state* s);
int64_t fg_state_rows(const fg_state* s);

/* Fold a dense device matrix W (rows x p, leading dimension ldw) and optional rhs
 * tau (rows, may be NULL = zeros) into the state. */
int fg_state_accumulate_dense(fg_state* s, const double* W_dev, int64_t ldw, const double* tau_dev, int64_t rows,
                              void* stream);

/* Fused path: regressor rows are generated in registers from (q, v, a) and folded
 * straight into the state; W never exists in HBM.  tau_dev is joint-major like the
 * reference's 1-D tau: entry (row block j, sample i) at tau_dev[j*ld_tau + i]
 * (ld_tau >= N; with row_sel, j counts selected row blocks).  The state's p must be
 * the regressor column count (fg_regressor_shape). */
int fg_state_accumulate_samples(fg_state* s, const fg_model* m, const fg_regressor_cfg* cfg, const double* q_dev,
                                const double* v_dev, const double* a_dev, const double* tau_dev, int64_t ld_tau,
                                int64_t N, void* stream);

/* Column order of the factor.  The fused kernel is free to factor [W P | tau] for a column
 * permutation P of its choosing (it eliminates the Ia/fv/fs/off columns, which are non-zero in one
 * row block only, before the inertial columns): state column i holds caller column perm[i].
 * fg_state_colsq / fg_state_reduce / fg_state_regroup_solve always speak caller columns; only the
 * packed exchange format below is in state order (and carries perm with it); identity after
 * create / reset. */
int fg_state_get_perm(const fg_state* s, int32_t* perm_out /* host [p] */);
int fg_state_set_perm(fg_state* s, const int32_t* perm /* host [p] */);

/* Packed form for the multi-GPU exchange: (p+1)^2 doubles of R (row-major, zeros below the
 * diagonal, columns in state order), (p+1) squared column norms, the row count as a double, then the
 * p entries of perm as doubles: (p+1)^2 + 2(p+1) doubles.  fg_state_merge_packed adopts the order of
 * the packed factors when the state is empty and refuses factors in a different order otherwise. */
int64_t fg_state_packed_doubles(int32_t p);
int fg_state_export(const fg_state* s, double* packed_dev, void* stream);
/* Merge `n` packed states (contiguous, n*fg_state_packed_doubles(p) doubles) into s,
 * in index order (deterministic). */
int fg_state_merge_packed(fg_state* s, const double* packed_dev, int32_t n, void* stream);

/* ---- kernel (3): small-matrix stage on the R factor ----------------------------------
 * All of get_index_eliminate / build_regressor_reduced / get_baseIndex / get_baseParams /
 * double_QR (regressor.py:258-293, qrdecomposition.py:89-296) expressed on the state.
 * These calls synchronise the stream (their outputs are host arrays). */

/* squared column norms, host [p+1] (last = |tau|^2): np.diag(W.T @ W) of regressor.py:271 */
int fg_state_colsq(fg_state* s, double* colsq_host, void* stream);

/* Keep the listed columns (ascending indices into the state's p columns; the rhs column
 * is always kept, last), re-triangularise, and return |diag| of the reduced R factor:
 * np.diag(np.linalg.qr(W_e)[1]) of qrdecomposition.py:107-121.  diag_host has n_keep+1
 * entries; the last one is the residual norm of tau against the kept columns. */
int fg_state_reduce(fg_state* s, const int32_t* keep_host, int32_t n_keep, double* diag_host, void* stream);

/* After fg_state_reduce: regroup the reduced columns as [base | regroup | rhs]
 * (indices into the n_keep reduced columns), re-triangularise, and solve
 *   beta  = R1^-1 R2      (n_base x n_regroup, row-major)    qrdecomposition.py:146
 *   phi_b = R1^-1 Q1^T tau (n_base)                          qrdecomposition.py:147
 * Also returns R1 (n_base x n_base, row-major, for cond / covariance) when R1_host != NULL
 * and resid2 = |tau - W_b phi_b|^2.  Unrounded values; the reference's np.round(.,6) is
 * applied by the Python layer. */
int fg_state_regroup_solve(fg_state* s, const int32_t* base_host, int32_t n_base, const int32_t* regroup_host,
                           int32_t n_regroup, double* beta_host, double* phi_host, double* R1_host,
                           double* resid2_host, void* stream);

/* Column-pivoted QR of the kept columns with the rhs carried along: scipy.linalg.qr(W_e, pivoting=True) of
 * QR_pivoting (qrdecomposition.py:24-86) evaluated on the factor (R^T R = W_e^T W_e, so the partial column
 * norms the pivoting looks at are the same).  perm_host[n_keep]: pivoted column j is kept column perm[j];
 * R_host: (n_keep+1) x (n_keep+1) row-major, [R P | Q^T tau]. */
int fg_state_pivot_qr(fg_state* s, const int32_t* keep_host, int32_t n_keep, int32_t* perm_host, double* R_host,
                      void* stream);

/* The whole N-independent stage in one launch and one device-to-host copy:
 *   keep  = columns with squared norm >= tol_e (get_index_eliminate, regressor.py:258-279; tol_e < 0: none
 *           dropped by norm) that are not listed in drop_host,
 *   base  = kept columns with |R_ii| > tol_qr in the re-triangularised factor (get_baseIndex,
 *           qrdecomposition.py:274-296), regroup = the others,
 *   beta, phi_b, R1, resid2 as fg_state_regroup_solve (double_QR, qrdecomposition.py:89-187).
 * keep_host[p], colsq_host[p+1], diag_host[p+1] (n_keep+1 valid), order_host[p]: indices into keep_host, the
 * n_base base columns first (ascending), then the regroup columns (ascending).  beta_host (n_base x
 * n_regroup), phi_host (n_base), R1_host (n_base x n_base) must have room for p*p / p doubles.
 * Returns FG_ERR_UNSUPPORTED for a permuted factor or one that does not fit shared memory: use the
 * three-call sequence then.  Leaves the state as after fg_state_reduce(keep). */
int fg_state_solve(fg_state* s, double tol_e, const int32_t* drop_host, int32_t n_drop, double tol_qr,
                   int32_t* n_keep, int32_t* keep_host, double* colsq_host, double* diag_host, int32_t* n_base,
                   int32_t* order_host, double* beta_host, double* phi_host, double* R1_host, double* resid2_host,
                   void* stream);

/* ---- preprocessing ahead of the regressor (device matrices in, device matrices out) -----------------
 * fg_filtfilt: scipy.signal.filtfilt(b, a, x, axis=0, padtype="odd", padlen=padlen) and removal of `trim`
 * rows at both ends = low_pass_filter_data (src/figaroh/identification/identification_tools.py:390-424;
 * b, a = butter(nbutter, ts*fc/2), padlen = 3*(max(len(a),len(b))-1), trim = 5*nbutter).  b_host, a_host,
 * zi_host (= scipy.signal.lfilter_zi(b, a)) are host arrays; order <= 8.  y: (N - 2*trim) x ncols.
 * fg_differentiate: calculate_first_second_order_differentiation (identification_tools.py:334-387) for
 * models whose configuration is a vector space (nq == nv, pin.difference = subtraction): dq = forward
 * difference / dt, ddq[:, j] = np.gradient(dq[:, j]) / dt for j < nq-1 and 0 for the last column (the
 * reference's range(model.nq - 1)), N - 2 rows out; dt_dev (N-1 entries) overrides the constant dt. */
int fg_filtfilt(const double* x_dev, int64_t N, int32_t ncols, int64_t ldx, const double* b_host, int32_t nb,
                const double* a_host, int32_t na, const double* zi_host, int32_t padlen, int64_t trim,
                double* y_dev, int64_t ldy, void* stream);
/* diag(W^T W) of a dense device matrix (rows x p, ld = ldw) - the column norms get_index_eliminate and
 * eliminate_non_dynaffect threshold (regressor.py:243,271: np.diag(np.dot(W.T, W))).  Any width. */
int fg_colsq_dense(const double* W_dev, int64_t ldw, int64_t rows, int32_t p, double* colsq_host, void* stream);
/* filtfilt(b, a, padtype="odd", padlen) per column followed by y[::q] - the transfer-function form of a zero-phase
 * decimation (scipy < 1.5 and complex designs; current scipy decimates through fg_sosfiltfilt's second-order
 * sections).  Rows 0, q, 2q, ... are written to y (ceil(N / q) rows). */
int fg_filtfilt_decimate(const double* x_dev, int64_t N, int32_t ncols, int64_t ldx, const double* b_host, int32_t nb,
                         const double* a_host, int32_t na, const double* zi_host, int32_t padlen, int32_t q,
                         double* y_dev, int64_t ldy, void* stream);
/* scipy.signal.sosfiltfilt(sos, x, axis=0, padtype="odd", padlen)[trim : N - trim : step] per column: what
 * scipy.signal.decimate(x, q, zero_phase=True) runs on its cheby1(8, 0.05, 0.8 / q, output="sos") design
 * (examples/tiago/identification.py:142-187).  sos: nsec x 6 host array (a0 == 1), zi = sosfilt_zi(sos) (nsec x 2). */
int fg_sosfiltfilt(const double* x_dev, int64_t N, int32_t ncols, int64_t ldx, const double* sos_host, int32_t nsec,
                   const double* zi_host, int32_t padlen, int64_t trim, int32_t step, double* y_dev, int64_t ldy,
                   void* stream);
/* scipy.signal.medfilt(x[:, c], k) per column, zero padded (examples/tiago/identification.py:70,80). */
int fg_medfilt(const double* x_dev, int64_t N, int32_t ncols, int64_t ldx, int32_t k, double* y_dev, int64_t ldy,
               void* stream);
/* np.gradient(x[:, c]) per column, divided by np.gradient(t) when t_dev != NULL (examples/tiago/identification.py:91-98). */
int fg_gradient(const double* x_dev, int64_t N, int32_t ncols, int64_t ldx, const double* t_dev, double* y_dev,
                int64_t ldy, void* stream);
int fg_differentiate(const double* q_dev, int64_t N, int32_t nq, int64_t ldq, double dt, const double* dt_dev,
                     double* dq_dev, double* ddq_dev, int64_t ldo, void* stream);

/* ---- host-buffer conveniences (end-to-end path: H2D copies inside the call) ----------
 * q,v,a,tau are HOST arrays (pinned or pageable); the call stages them through device
 * buffers in chunks, overlapping copies with the fused kernel, and folds them into s. */
int fg_state_accumulate_samples_host(fg_state* s, const fg_model* m, const fg_regressor_cfg* cfg,
                                     const double* q_host, const double* v_host, const double* a_host,
                                     const double* tau_host, int64_t ld_tau, int64_t N, void* stream);

/* ---- diagnostics (used by bench.py; not part of the reference-facing surface) ---------
 * fg_launch_count: kernels launched by this library since load.
 * fg_state_set_timing(on): bracket every launch of the dominant fold kernel
 * (accumulate_samples / accumulate_dense) with CUDA events on the caller's stream;
 * fg_state_get_timing returns the summed duration in ms and the number of launches. */
long long fg_launch_count(void);
/* Host-only (no device needed): JSON description of the fused path's symbolic pass for (model, cfg) - the
 * producers' depth-first walk program over the kinematic tree, the local column problem of every row group
 * (the non-zero columns of the rows of regressor.py:49-70 for that joint) and the warp split.  There is no
 * reference counterpart; it exposes the plan for tests and documentation. */
int fg_plan_describe(const fg_model_desc* desc, const fg_regressor_cfg* cfg, int32_t sm_count, int32_t max_smem,
                     char* buf, int32_t buflen);
int fg_state_set_timing(fg_state* s, int on);
int fg_state_get_timing(fg_state* s, double* total_ms, long long* launches);

#ifdef __cplusplus
}
#endif
#endif /* FIGAROH_B200_H */
