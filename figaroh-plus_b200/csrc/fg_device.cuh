// fg_device.cuh - device-side building blocks of the B200 identification path.
//
//  * spatial algebra + per-sample forward kinematics (what pinocchio's
//    JointTorqueRegressorForwardStep does; call site regressor.py:49 of the reference),
//  * per-column regressor generation in the "column-owner" layout: lane l of a warp owns
//    columns l, l+32, l+64, ... of the regressor and computes exactly those entries,
//  * the Householder fold of a register-resident row tile into a triangular R factor.
//
// Everything is fp64, no fast-math, sm_100a only.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fg {

constexpr int kWarp = 32;
constexpr int kMaxJoints = 64;     // incl. universe
constexpr int kMaxRowGroups = 48;  // row blocks per sample

enum ColKind : int { COL_ZERO = 0, COL_INERTIAL = 1, COL_EXTRA = 2, COL_COUPLING = 3, COL_RHS = 4 };
enum ExtraSrc : int { SRC_QDD = 0, SRC_QD = 1, SRC_SIGN_QD = 2, SRC_ONE = 3 };

struct DevJoint {
  double plcR[9];
  double plcp[3];
  double axis[3];
  int parent, jtype, idx_q, idx_v;
};

// one regressor column (32 bytes)
struct ColDesc {
  int kind;   // ColKind
  int body;   // INERTIAL: joint index of the body (1-based)
  int k;      // INERTIAL: pinocchio parameter 0..9; EXTRA: ExtraSrc; COUPLING: 0 Iam6, 1 fvm6, 2 fsm6
  int vidx;   // EXTRA: velocity index of the source
  int grp;    // EXTRA: target row group, or -1 = all row groups, -2 = none
  unsigned anc_lo, anc_hi;  // INERTIAL: bit g set iff row group g's joint is the body's joint or an ancestor
  int pad;
};
__host__ __device__ inline uint64_t col_ancmask(const ColDesc& c) { return ((uint64_t)c.anc_hi << 32) | c.anc_lo; }

// one row group of a sample (64 bytes): the rows S^T B of one joint DoF
struct RowDesc {
  double S[6];   // selection vector applied to (f, n): [0;axis] revolute, [axis;0] prismatic, e_k free-flyer
  int joint;     // joint whose frame the wrench is expressed in
  int rowblock;  // block index j of the reference's row j*N+i
  int start_col; // first column in which this row group can be non-zero
  int pad;
};

struct Plan {
  int nj;        // njoints incl. universe
  int nq, nv;
  int ms;        // row groups per sample
  int ncols;     // regressor columns (without rhs)
  int p1;        // ncols + 1
  int mode, flags;
  double gravity[3];
};

__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }

struct DevPlanPtrs {
  const Plan* plan;
  const DevJoint* joints;
  const RowDesc* rows;
  const ColDesc* cols;  // [p1] (last = rhs)
  const int* nact;      // [ms + 1] gstart: first column at which row group g is active; [ms] = p1
  const int* fullcol;   // [p1] full (reference) column index of each compact column, -1 for rhs
};

// A "part" is one packed upper-triangular factor (p1 columns) plus p1 squared column norms.
__host__ __device__ inline int64_t part_doubles(int p1) { return (int64_t)p1 * (p1 + 1) / 2 + p1; }

// per-joint forward state of one sample, kept in shared memory
constexpr int FW_R = 0, FW_P = 9, FW_V = 12, FW_A = 18, FW_AP = 24, FW_STRIDE = 27;

__device__ __forceinline__ void cross3(const double* a, const double* b, double* o) {
  double x = a[1] * b[2] - a[2] * b[1];
  double y = a[2] * b[0] - a[0] * b[2];
  double z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}

// R = (1-c) a a^T + c I + s [a]x, with exact 0/1 entries for axis-aligned joints
__device__ __forceinline__ void axis_rot(const double* a, double c, double s, double* R) {
  if (a[0] == 1.0 && a[1] == 0.0 && a[2] == 0.0) {
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = c; R[5] = -s; R[6] = 0; R[7] = s; R[8] = c;
  } else if (a[0] == 0.0 && a[1] == 1.0 && a[2] == 0.0) {
    R[0] = c; R[1] = 0; R[2] = s; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = -s; R[7] = 0; R[8] = c;
  } else if (a[0] == 0.0 && a[1] == 0.0 && a[2] == 1.0) {
    R[0] = c; R[1] = -s; R[2] = 0; R[3] = s; R[4] = c; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
  } else {
    double c1x = (1 - c) * a[0], c1y = (1 - c) * a[1], c1z = (1 - c) * a[2];
    double sx = s * a[0], sy = s * a[1], sz = s * a[2];
    R[0] = c1x * a[0] + c;  R[1] = c1x * a[1] - sz; R[2] = c1x * a[2] + sy;
    R[3] = c1y * a[0] + sz; R[4] = c1y * a[1] + c;  R[5] = c1y * a[2] - sx;
    R[6] = c1z * a[0] - sy; R[7] = c1z * a[1] + sx; R[8] = c1z * a[2] + c;
  }
}

__device__ __forceinline__ void quat_rot(double x, double y, double z, double w, double* R) {
  double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
  double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

// Forward pass for ONE sample, executed by one thread.  fw points at this sample's
// [nj][FW_STRIDE] block in shared memory; q/qd/qdd at this sample's rows in global memory.
// Also copies qd/qdd into shared memory (vs, as) for the friction / inertia / offset columns.
__device__ inline void forward_sample(const Plan& P, const DevJoint* J, const double* __restrict__ q,
                                      const double* __restrict__ qd, const double* __restrict__ qdd,
                                      double* fw, double* vs, double* as, bool valid) {
  // universe
  for (int k = 0; k < FW_STRIDE; ++k) fw[k] = 0.0;
  fw[FW_R + 0] = fw[FW_R + 4] = fw[FW_R + 8] = 1.0;
  fw[FW_A + 0] = -P.gravity[0]; fw[FW_A + 1] = -P.gravity[1]; fw[FW_A + 2] = -P.gravity[2];
  if (vs != nullptr) {
    for (int k = 0; k < P.nv; ++k) { vs[k] = valid ? qd[k] : 0.0; as[k] = valid ? qdd[k] : 0.0; }
  } else {  // no shared-memory copy of (dq, ddq): read them where they are
    vs = const_cast<double*>(qd);
    as = const_cast<double*>(qdd);
  }
  for (int i = 1; i < P.nj; ++i) {
    const DevJoint& jt = J[i];
    double Rj[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pj[3] = {0, 0, 0};
    double vj[6] = {0, 0, 0, 0, 0, 0}, aj[6] = {0, 0, 0, 0, 0, 0};
    const int iq = jt.idx_q, iv = jt.idx_v;
    if (jt.jtype == 1) {
      double s, c;
      sincos(valid ? q[iq] : 0.0, &s, &c);
      axis_rot(jt.axis, c, s, Rj);
      for (int k = 0; k < 3; ++k) { vj[3 + k] = jt.axis[k] * vs[iv]; aj[3 + k] = jt.axis[k] * as[iv]; }
    } else if (jt.jtype == 2) {
      axis_rot(jt.axis, valid ? q[iq] : 1.0, valid ? q[iq + 1] : 0.0, Rj);
      for (int k = 0; k < 3; ++k) { vj[3 + k] = jt.axis[k] * vs[iv]; aj[3 + k] = jt.axis[k] * as[iv]; }
    } else if (jt.jtype == 3) {
      double qq = valid ? q[iq] : 0.0;
      for (int k = 0; k < 3; ++k) { pj[k] = jt.axis[k] * qq; vj[k] = jt.axis[k] * vs[iv]; aj[k] = jt.axis[k] * as[iv]; }
    } else if (jt.jtype == 4) {
      if (valid) quat_rot(q[iq + 3], q[iq + 4], q[iq + 5], q[iq + 6], Rj);
      for (int k = 0; k < 3; ++k) pj[k] = valid ? q[iq + k] : 0.0;
      for (int k = 0; k < 6; ++k) { vj[k] = vs[iv + k]; aj[k] = as[iv + k]; }
    }
    double* o = fw + i * FW_STRIDE;
    const double* Pp = fw + jt.parent * FW_STRIDE;
    // liMi = placement * Mj
    double R[9], p[3];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c)
        R[3 * r + c] = jt.plcR[3 * r] * Rj[c] + jt.plcR[3 * r + 1] * Rj[3 + c] + jt.plcR[3 * r + 2] * Rj[6 + c];
      p[r] = jt.plcR[3 * r] * pj[0] + jt.plcR[3 * r + 1] * pj[1] + jt.plcR[3 * r + 2] * pj[2] + jt.plcp[r];
    }
    // v = vJ + liMi.actInv(v_parent)
    double v[6], a[6], t[3], u[3];
    for (int k = 0; k < 6; ++k) v[k] = vj[k];
    if (jt.parent > 0) {
      cross3(p, Pp + FW_V + 3, t);
      for (int k = 0; k < 3; ++k) t[k] = Pp[FW_V + k] - t[k];
      for (int k = 0; k < 3; ++k) v[k] += R[k] * t[0] + R[3 + k] * t[1] + R[6 + k] * t[2];
      for (int k = 0; k < 3; ++k)
        v[3 + k] += R[k] * Pp[FW_V + 3] + R[3 + k] * Pp[FW_V + 4] + R[6 + k] * Pp[FW_V + 5];
    }
    // a = v x vJ + S qdd + liMi.actInv(a_parent)
    double l1[3], l2[3], an[3];
    cross3(v + 3, vj, l1);
    cross3(v, vj + 3, l2);
    cross3(v + 3, vj + 3, an);
    for (int k = 0; k < 3; ++k) { a[k] = l1[k] + l2[k] + aj[k]; a[3 + k] = an[k] + aj[3 + k]; }
    cross3(p, Pp + FW_A + 3, t);
    for (int k = 0; k < 3; ++k) t[k] = Pp[FW_A + k] - t[k];
    for (int k = 0; k < 3; ++k) a[k] += R[k] * t[0] + R[3 + k] * t[1] + R[6 + k] * t[2];
    for (int k = 0; k < 3; ++k)
      a[3 + k] += R[k] * Pp[FW_A + 3] + R[3 + k] * Pp[FW_A + 4] + R[6 + k] * Pp[FW_A + 5];
    // classical acceleration a' = a_lin + w x v_lin
    cross3(v + 3, v, u);
    for (int k = 0; k < 9; ++k) o[FW_R + k] = R[k];
    for (int k = 0; k < 3; ++k) o[FW_P + k] = p[k];
    for (int k = 0; k < 6; ++k) { o[FW_V + k] = v[k]; o[FW_A + k] = a[k]; }
    for (int k = 0; k < 3; ++k) o[FW_AP + k] = a[k] + u[k];
  }
}

__device__ __forceinline__ double sel3(const double* x, int j) { return j == 0 ? x[0] : (j == 1 ? x[1] : x[2]); }

// column k (pinocchio order m, mc_xyz, Ixx Ixy Iyy Ixz Iyz Izz) of the 6x10 body regressor
__device__ __forceinline__ void body_column(int k, const double* fwj, double* f, double* n) {
  const double* ap = fwj + FW_AP;
  const double* al = fwj + FW_A + 3;
  const double* om = fwj + FW_V + 3;
  if (k == 0) {
    f[0] = ap[0]; f[1] = ap[1]; f[2] = ap[2];
    n[0] = n[1] = n[2] = 0.0;
  } else if (k < 4) {
    const int j = k - 1;
    const double e[3] = {j == 0 ? 1.0 : 0.0, j == 1 ? 1.0 : 0.0, j == 2 ? 1.0 : 0.0};
    const double ww = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
    const double oj = sel3(om, j);
    double x[3];
    cross3(al, e, x);                       // column j of skew(alpha)
    for (int i = 0; i < 3; ++i) f[i] = x[i] + om[i] * oj - ww * e[i];
    cross3(e, ap, n);                       // column j of -skew(a')
  } else {
    // symmetric basis matrix E for Ixx,Ixy,Iyy,Ixz,Iyz,Izz : (iu,iw) = (0,0),(0,1),(1,1),(0,2),(1,2),(2,2)
    const int c = k - 4;
    const int iu = (c == 2 || c == 4) ? 1 : (c == 5 ? 2 : 0);
    const int iw = (c == 0) ? 0 : ((c == 1 || c == 2) ? 1 : 2);
    double Ea[3], Eo[3];
    const double a_w = sel3(al, iw), a_u = sel3(al, iu), o_w = sel3(om, iw), o_u = sel3(om, iu);
    for (int i = 0; i < 3; ++i) {
      const double du = (i == iu) ? 1.0 : 0.0, dw = (i == iw && iw != iu) ? 1.0 : 0.0;
      Ea[i] = du * a_w + dw * a_u;
      Eo[i] = du * o_w + dw * o_u;
    }
    double x[3];
    cross3(om, Eo, x);
    f[0] = f[1] = f[2] = 0.0;
    for (int i = 0; i < 3; ++i) n[i] = Ea[i] + x[i];
  }
}

// (f, n) <- liMi o (f, n): f' = R f, n' = R n + p x f'
__device__ __forceinline__ void force_hop(const double* fwj, double* f, double* n) {
  const double* R = fwj + FW_R;
  const double* p = fwj + FW_P;
  double f2[3], n2[3], x[3];
  for (int i = 0; i < 3; ++i) {
    f2[i] = R[3 * i] * f[0] + R[3 * i + 1] * f[1] + R[3 * i + 2] * f[2];
    n2[i] = R[3 * i] * n[0] + R[3 * i + 1] * n[1] + R[3 * i + 2] * n[2];
  }
  cross3(p, f2, x);
  for (int i = 0; i < 3; ++i) { f[i] = f2[i]; n[i] = n2[i] + x[i]; }
}

__device__ __forceinline__ double sgn(double x) { return (double)((x > 0.0) - (x < 0.0)); }

// Entries of ONE regressor column for S samples and all row groups:
//   out[g*S + s] = W[row group g of sample s, this column].
// fw: [S][nj][FW_STRIDE] forward states, vs/as: [S][nv] velocities / accelerations (shared memory).
// tau: joint-major global rhs (entry (rowblock, sample) at tau[rowblock*ld_tau + sample]) or null.
template <int MSMAX, int S>
__device__ __forceinline__ void gen_column(const Plan& P, const DevJoint* J, const RowDesc* rows, const ColDesc cd,
                                           uint64_t ancmask, const double* fw, const double* vs, const double* as,
                                           const double* __restrict__ tau, int64_t ld_tau, int64_t sample0, int64_t N,
                                           double (&out)[MSMAX * S]) {
#pragma unroll
  for (int i = 0; i < MSMAX * S; ++i) out[i] = 0.0;
  if (cd.kind == COL_INERTIAL) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      if (sample0 + s >= N) continue;  // padding sample of the last tile: all-zero rows
      const double* fws = fw + (size_t)s * P.nj * FW_STRIDE;
      double f[3], n[3];
      body_column(cd.k, fws + cd.body * FW_STRIDE, f, n);
      int cur = cd.body;
#pragma unroll
      for (int g = MSMAX - 1; g >= 0; --g) {
        if (g < P.ms && ((ancmask >> g) & 1ull)) {
          const int jg = rows[g].joint;
          while (cur != jg) {
            force_hop(fws + cur * FW_STRIDE, f, n);
            cur = J[cur].parent;
          }
          const double* Sg = rows[g].S;
          out[g * S + s] = Sg[0] * f[0] + Sg[1] * f[1] + Sg[2] * f[2] + Sg[3] * n[0] + Sg[4] * n[1] + Sg[5] * n[2];
        }
      }
    }
  } else if (cd.kind == COL_EXTRA) {
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const bool valid = sample0 + s < N;
      const double vv = valid ? vs[s * P.nv + cd.vidx] : 0.0, aa = valid ? as[s * P.nv + cd.vidx] : 0.0;
      double val = cd.k == SRC_QDD ? aa : (cd.k == SRC_QD ? vv : (cd.k == SRC_SIGN_QD ? sgn(vv) : 1.0));
      if (!valid) val = 0.0;
#pragma unroll
      for (int g = 0; g < MSMAX; ++g)
        if (g < P.ms && (cd.grp == -1 || cd.grp == g)) out[g * S + s] = val;
    }
  } else if (cd.kind == COL_COUPLING) {
    // add_coupling_TX40 (regressor.py:216-225): row block 4 <- [a5, v5, sign(v4+v5)], row block 5 <- [a4, v4, sign(v4+v5)]
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const bool valid = sample0 + s < N;
      const int so = valid ? s * P.nv : 0;
      const double v4 = vs[so + 4], v5 = vs[so + 5], a4 = as[so + 4], a5 = as[so + 5];
      const double sg = sgn(v4 + v5);
#pragma unroll
      for (int g = 0; g < MSMAX; ++g) {
        if (g < P.ms && valid) {
          const int rb = rows[g].rowblock;
          if (rb == 4) out[g * S + s] = cd.k == 0 ? a5 : (cd.k == 1 ? v5 : sg);
          if (rb == 5) out[g * S + s] = cd.k == 0 ? a4 : (cd.k == 1 ? v4 : sg);
        }
      }
    }
  } else if (cd.kind == COL_RHS) {
    if (tau != nullptr) {
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const bool valid = sample0 + s < N;
#pragma unroll
        for (int g = 0; g < MSMAX; ++g)
          if (g < P.ms && valid) out[g * S + s] = tau[(int64_t)g * ld_tau + sample0 + s];
      }
    }
  }
}

// offset of row c in a row-major packed upper-triangular matrix with p1 columns
__device__ __forceinline__ int tri_off(int c, int p1) { return c * p1 - (c * (c - 1)) / 2; }

// ------------------------------------------------------------------------------------
// Householder fold of a register tile into R.
//   B[r][t]   : tile entry (row r, column lane + 32 t); rows come in NG = TR/GS groups of GS
//               rows whose activation column is non-decreasing with the group index,
//   Rt        : packed upper-triangular (p1 columns) factor, shared or global memory,
//   xs        : per-warp scratch, 2 * (TR + 2) doubles (double-buffered reflector broadcast),
//   gstart    : gstart[g] = first column at which row group g can be non-zero (g < NG) and
//               gstart[NG] = p1; nullptr = every group active from column 0,
//   c_begin   : first column that can be non-zero in this tile.
// Reflector c annihilates column c of the tile against R[c][c] (LAPACK dlarfg convention:
// beta = -sign(alpha) * norm, v = x / (alpha - beta), tau = (beta - alpha) / beta).
// All p1 columns are eliminated, the rhs column last, so R[p1-1][p1-1] carries the residual.
//
// The column loop is split by (slot of the pivot column, number of active row groups) so that
// every inner loop has compile-time bounds: rows that are still structurally zero cost nothing
// (the block-triangular shape of a joint-torque regressor makes that ~60% of the dense work).

// one reflector; NA = number of active rows (a multiple of GS), TO = slot of the pivot column
template <int NSLOT, int TR, int NA, int TO>
__device__ __forceinline__ void fold_column(double (&B)[TR][NSLOT], double* Rt, int p1, double* xs, int c, int lane) {
  constexpr int XS = TR + 2;
  const int o = c - 32 * TO;
  double* xb = xs + (c & 1) * XS;
  const int roff = tri_off(c, p1);
  if (lane == o) {
#pragma unroll
    for (int r = 0; r < NA; ++r) xb[r] = B[r][TO];
    xb[TR] = Rt[roff];  // alpha
  }
  __syncwarp();
  double v[NA];
  double sg[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int r = 0; r < NA; ++r) { v[r] = xb[r]; sg[r & 3] = fma(v[r], v[r], sg[r & 3]); }
  const double sigma = (sg[0] + sg[1]) + (sg[2] + sg[3]);
  if (sigma == 0.0) return;  // nothing to annihilate: H = I (dlarfg with xnorm == 0)
  const double alpha = xb[TR];
  const double mu = fma(alpha, alpha, sigma);
  const double rs = rsqrt(mu);
  const double nrm = mu * rs;
  const double beta = alpha >= 0.0 ? -nrm : nrm;
  const double den = alpha - beta;
  const double inv = __drcp_rn(den);
  const double tauh = fabs(den) * rs;  // (beta - alpha) / beta
#pragma unroll
  for (int t = TO; t < NSLOT; ++t) {
    const int col = lane + 32 * t;
    const bool act = (t > TO || lane > o) && col < p1;
    double d[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int r = 0; r < NA; ++r) d[r & 3] = fma(v[r], B[r][t], d[r & 3]);
    const double dd = (d[0] + d[1]) + (d[2] + d[3]);
    double gco = 0.0;
    if (act) {
      const double rho = Rt[roff + col - c];
      const double sdot = fma(dd, inv, rho);  // w^T [rho; b],  w = [1; x * inv]
      Rt[roff + col - c] = fma(-tauh, sdot, rho);
      gco = tauh * sdot * inv;
    }
#pragma unroll
    for (int r = 0; r < NA; ++r) B[r][t] = fma(-gco, v[r], B[r][t]);
  }
  if (lane == o) Rt[roff] = beta;
}

template <int NSLOT, int TR, int GS, int TO, int NGA>
__device__ __forceinline__ void fold_ranges(double (&B)[TR][NSLOT], double* Rt, int p1, double* xs,
                                            const int* gstart, int c_begin, int lane) {
  constexpr int NG = TR / GS;
  if constexpr (NGA <= NG) {
    int lo = gstart ? gstart[NGA - 1] : 0, hi = gstart ? gstart[NGA] : p1;
    if (gstart == nullptr && NGA < NG) hi = 0;  // dense tiles: only the all-active range exists
    if (lo < 32 * TO) lo = 32 * TO;
    if (lo < c_begin) lo = c_begin;
    if (hi > 32 * TO + 32) hi = 32 * TO + 32;
    if (hi > p1) hi = p1;
    for (int c = lo; c < hi; ++c) fold_column<NSLOT, TR, NGA * GS, TO>(B, Rt, p1, xs, c, lane);
    fold_ranges<NSLOT, TR, GS, TO, NGA + 1>(B, Rt, p1, xs, gstart, c_begin, lane);
  }
}

template <int NSLOT, int TR, int GS, int TO>
__device__ __forceinline__ void fold_slots(double (&B)[TR][NSLOT], double* Rt, int p1, double* xs,
                                           const int* gstart, int c_begin, int lane) {
  if constexpr (TO < NSLOT) {
    if (32 * TO < p1 && 32 * TO + 32 > c_begin) fold_ranges<NSLOT, TR, GS, TO, 1>(B, Rt, p1, xs, gstart, c_begin, lane);
    fold_slots<NSLOT, TR, GS, TO + 1>(B, Rt, p1, xs, gstart, c_begin, lane);
  }
}

template <int NSLOT, int TR, int GS>
__device__ __forceinline__ void fold_tile(double (&B)[TR][NSLOT], double* Rt, int p1, double* xs,
                                          const int* gstart, int c_begin, int lane) {
  fold_slots<NSLOT, TR, GS, 0>(B, Rt, p1, xs, gstart, c_begin, lane);
}

}  // namespace fg
