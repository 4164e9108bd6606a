/*
 * fg_oracle.c - CPU restatement (plain C, fp64) of the third-party arithmetic on
 * FIGAROH's dynamic-identification path.  TEST INFRASTRUCTURE ONLY: nothing under
 * figaroh-plus_b200/ may import, link or call this file (see oracle/README.md).
 *
 * What it restates
 * ----------------
 * pinocchio::computeJointTorqueRegressor and pinocchio::rnea, which the reference
 * calls at /root/reference/src/figaroh/tools/regressor.py:49,93 and
 * /root/reference/src/figaroh/tools/randomdata.py:110.  Pinocchio is a conda
 * dependency (unpinned: /root/reference/environment.yml:7) whose sources are not
 * under /root/reference and which is not installable here, so the published
 * algorithm is restated (SURVEY.md section 8a "Algorithm box"):
 *
 *   forward, i = 1..n (parents first):
 *     liMi = jointPlacement_i * jointTransform_i(q)
 *     v_i  = S_i qd_i + liMi^-1 o v_parent
 *     a_i  = S_i qdd_i + v_i x (S_i qd_i) + liMi^-1 o a_parent,  a_0 = -gravity
 *   body regressor B_i (6x10) from (v_i, a_i), columns [m | mc | Ixx Ixy Iyy Ixz Iyz Izz]
 *   backward, for each body i and each ancestor-or-self j:
 *     W[idx_v(j) .. +nv_j, 10(i-1) .. 10i] = S_j^T B ;  B <- liMi_j o B
 *
 * Spatial vectors are [linear; angular].  Parity of this restatement is pinned by
 * the reference's own golden CSVs (tests/test_oracle_golden.py): the TIAGo base
 * parameter file is reproduced exactly, the TX40 one to all expression strings.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { JT_NONE = 0, JT_REVOLUTE = 1, JT_REVOLUTE_UNB = 2, JT_PRISMATIC = 3, JT_FREEFLYER = 4 };

typedef struct {
  int njoints, nq, nv;
  const int *parent, *jtype;
  const double *axis;   /* [njoints][3] */
  const double *plc_R;  /* [njoints][9] row-major */
  const double *plc_p;  /* [njoints][3] */
  const int *idx_q, *idx_v;
  const double *gravity; /* [3] */
  const double *mass;    /* [njoints] */
  const double *lever;   /* [njoints][3] */
  const double *inertia; /* [njoints][6] xx,xy,yy,xz,yz,zz about the CoM */
} orc_model;

static void cross3(const double *a, const double *b, double *o) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
static void matvec3(const double *R, const double *x, double *o) { /* o = R x */
  double a = R[0] * x[0] + R[1] * x[1] + R[2] * x[2];
  double b = R[3] * x[0] + R[4] * x[1] + R[5] * x[2];
  double c = R[6] * x[0] + R[7] * x[1] + R[8] * x[2];
  o[0] = a; o[1] = b; o[2] = c;
}
static void matTvec3(const double *R, const double *x, double *o) { /* o = R^T x */
  double a = R[0] * x[0] + R[3] * x[1] + R[6] * x[2];
  double b = R[1] * x[0] + R[4] * x[1] + R[7] * x[2];
  double c = R[2] * x[0] + R[5] * x[1] + R[8] * x[2];
  o[0] = a; o[1] = b; o[2] = c;
}
static void matmul3(const double *A, const double *B, double *C) {
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(C, t, sizeof t);
}

/* rotation about a unit axis given cos/sin: (1-c) a a^T + c I + s [a]x (Rodrigues) */
static void axis_rot(const double *a, double c, double s, double *R) {
  /* axis-aligned joints (pinocchio JointModelRX/RY/RZ) produce exact 0/1 entries */
  if (a[0] == 1.0 && a[1] == 0.0 && a[2] == 0.0) {
    double t[9] = {1, 0, 0, 0, c, -s, 0, s, c}; memcpy(R, t, sizeof t); return;
  }
  if (a[0] == 0.0 && a[1] == 1.0 && a[2] == 0.0) {
    double t[9] = {c, 0, s, 0, 1, 0, -s, 0, c}; memcpy(R, t, sizeof t); return;
  }
  if (a[0] == 0.0 && a[1] == 0.0 && a[2] == 1.0) {
    double t[9] = {c, -s, 0, s, c, 0, 0, 0, 1}; memcpy(R, t, sizeof t); return;
  }
  double c1[3] = {(1 - c) * a[0], (1 - c) * a[1], (1 - c) * a[2]};
  double sa[3] = {s * a[0], s * a[1], s * a[2]};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) R[3 * i + j] = c1[i] * a[j];
  R[1] -= sa[2]; R[2] += sa[1]; R[3] += sa[2]; R[5] -= sa[0]; R[6] -= sa[1]; R[7] += sa[0];
  R[0] += c; R[4] += c; R[8] += c;
}

static void quat_rot(const double *q /* x y z w */, double *R) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x;
  double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}

typedef struct {
  double R[9], p[3];   /* liMi */
  double v[6], a[6];   /* spatial velocity / acceleration (with gravity field), local frame */
} orc_jstate;

/* forward kinematics pass shared by the regressor and RNEA */
static void forward_pass(const orc_model *m, const double *q, const double *qd, const double *qdd, orc_jstate *st) {
  memset(&st[0], 0, sizeof(orc_jstate));
  st[0].R[0] = st[0].R[4] = st[0].R[8] = 1.0;
  st[0].a[0] = -m->gravity[0]; st[0].a[1] = -m->gravity[1]; st[0].a[2] = -m->gravity[2];
  for (int i = 1; i < m->njoints; ++i) {
    const double *ax = m->axis + 3 * i;
    const int iq = m->idx_q[i], iv = m->idx_v[i], par = m->parent[i];
    double Rj[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pj[3] = {0, 0, 0};
    double vj[6] = {0, 0, 0, 0, 0, 0}, aj[6] = {0, 0, 0, 0, 0, 0};
    switch (m->jtype[i]) {
      case JT_REVOLUTE: axis_rot(ax, cos(q[iq]), sin(q[iq]), Rj);
        for (int k = 0; k < 3; ++k) { vj[3 + k] = ax[k] * qd[iv]; aj[3 + k] = ax[k] * qdd[iv]; }
        break;
      case JT_REVOLUTE_UNB: axis_rot(ax, q[iq], q[iq + 1], Rj);
        for (int k = 0; k < 3; ++k) { vj[3 + k] = ax[k] * qd[iv]; aj[3 + k] = ax[k] * qdd[iv]; }
        break;
      case JT_PRISMATIC:
        for (int k = 0; k < 3; ++k) { pj[k] = ax[k] * q[iq]; vj[k] = ax[k] * qd[iv]; aj[k] = ax[k] * qdd[iv]; }
        break;
      case JT_FREEFLYER: quat_rot(q + iq + 3, Rj);
        for (int k = 0; k < 3; ++k) pj[k] = q[iq + k];
        for (int k = 0; k < 6; ++k) { vj[k] = qd[iv + k]; aj[k] = qdd[iv + k]; }
        break;
      default: break;
    }
    orc_jstate *s = &st[i];
    const double *Rp = m->plc_R + 9 * i, *pp = m->plc_p + 3 * i;
    matmul3(Rp, Rj, s->R);
    matvec3(Rp, pj, s->p);
    for (int k = 0; k < 3; ++k) s->p[k] += pp[k];
    /* v_i = vJ + liMi.actInv(v_parent) */
    memcpy(s->v, vj, sizeof vj);
    if (par > 0) {
      const orc_jstate *P = &st[par];
      double t[3], u[3];
      cross3(s->p, P->v + 3, t);
      for (int k = 0; k < 3; ++k) t[k] = P->v[k] - t[k];
      matTvec3(s->R, t, u);
      for (int k = 0; k < 3; ++k) s->v[k] += u[k];
      matTvec3(s->R, P->v + 3, u);
      for (int k = 0; k < 3; ++k) s->v[3 + k] += u[k];
    }
    /* a_i = (v_i x vJ) + S qdd + liMi.actInv(a_parent)   [joint bias c_J = 0 for these joints] */
    {
      double l1[3], l2[3], an[3];
      cross3(s->v + 3, vj, l1);      /* w1 x v2 */
      cross3(s->v, vj + 3, l2);      /* v1 x w2 */
      cross3(s->v + 3, vj + 3, an);  /* w1 x w2 */
      for (int k = 0; k < 3; ++k) { s->a[k] = l1[k] + l2[k] + aj[k]; s->a[3 + k] = an[k] + aj[3 + k]; }
      const orc_jstate *P = &st[par];
      double t[3], u[3];
      cross3(s->p, P->a + 3, t);
      for (int k = 0; k < 3; ++k) t[k] = P->a[k] - t[k];
      matTvec3(s->R, t, u);
      for (int k = 0; k < 3; ++k) s->a[k] += u[k];
      matTvec3(s->R, P->a + 3, u);
      for (int k = 0; k < 3; ++k) s->a[3 + k] += u[k];
    }
  }
}

/* 6x10 body regressor, row-major B[6][10]: f = B * [m, mc, Ixx,Ixy,Iyy,Ixz,Iyz,Izz] */
static void body_regressor(const double *v, const double *a, double B[6][10]) {
  const double *w = v + 3, *al = a + 3;
  double acc[3], t[3];
  cross3(w, v, t);
  for (int k = 0; k < 3; ++k) acc[k] = a[k] + t[k];
  memset(B, 0, sizeof(double) * 60);
  for (int k = 0; k < 3; ++k) B[k][0] = acc[k];
  /* linear rows, mc columns: skew(alpha) + skew(w)^2,  skew(w)^2 = w w^T - |w|^2 I */
  double ww = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double sk[3][3] = {{0, -al[2], al[1]}, {al[2], 0, -al[0]}, {-al[1], al[0], 0}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) B[i][1 + j] = sk[i][j] + w[i] * w[j] - (i == j ? ww : 0.0);
  /* angular rows, mc columns: -skew(acc) */
  B[3][2] = acc[2];  B[3][3] = -acc[1];
  B[4][1] = -acc[2]; B[4][3] = acc[0];
  B[5][1] = acc[1];  B[5][2] = -acc[0];
  /* angular rows, inertia columns: L(alpha) + skew(w) L(w) */
  double L1[3][6] = {{al[0], al[1], 0, al[2], 0, 0}, {0, al[0], al[1], 0, al[2], 0}, {0, 0, 0, al[0], al[1], al[2]}};
  double L2[3][6] = {{w[0], w[1], 0, w[2], 0, 0}, {0, w[0], w[1], 0, w[2], 0}, {0, 0, 0, w[0], w[1], w[2]}};
  double sw[3][3] = {{0, -w[2], w[1]}, {w[2], 0, -w[0]}, {-w[1], w[0], 0}};
  for (int i = 0; i < 3; ++i)
    for (int c = 0; c < 6; ++c) {
      double s = L1[i][c];
      for (int k = 0; k < 3; ++k) s += sw[i][k] * L2[k][c];
      B[3 + i][4 + c] = s;
    }
}

/* W is (nv x 10(njoints-1)) row-major, fully overwritten */
int orc_joint_torque_regressor(const orc_model *m, const double *q, const double *qd, const double *qdd, double *W) {
  const int n = m->njoints, ncol = 10 * (n - 1);
  orc_jstate *st = (orc_jstate *)malloc(sizeof(orc_jstate) * n);
  if (!st) return -1;
  forward_pass(m, q, qd, qdd, st);
  memset(W, 0, sizeof(double) * (size_t)m->nv * ncol);
  for (int i = n - 1; i >= 1; --i) {
    double B[6][10];
    body_regressor(st[i].v, st[i].a, B);
    for (int j = i; j > 0; j = m->parent[j]) {
      const int iv = m->idx_v[j];
      const double *ax = m->axis + 3 * j;
      double *dst = W + (size_t)iv * ncol + 10 * (i - 1);
      switch (m->jtype[j]) {
        case JT_REVOLUTE: case JT_REVOLUTE_UNB:
          for (int c = 0; c < 10; ++c) dst[c] = ax[0] * B[3][c] + ax[1] * B[4][c] + ax[2] * B[5][c];
          break;
        case JT_PRISMATIC:
          for (int c = 0; c < 10; ++c) dst[c] = ax[0] * B[0][c] + ax[1] * B[1][c] + ax[2] * B[2][c];
          break;
        case JT_FREEFLYER:
          for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 10; ++c) dst[(size_t)r * ncol + c] = B[r][c];
          break;
        default: break;
      }
      if (m->parent[j] > 0) { /* B <- liMi_j o B  (force action column by column) */
        const double *R = st[j].R, *p = st[j].p;
        for (int c = 0; c < 10; ++c) {
          double f[3] = {B[0][c], B[1][c], B[2][c]}, t[3] = {B[3][c], B[4][c], B[5][c]}, f2[3], t2[3], x[3];
          matvec3(R, f, f2); matvec3(R, t, t2); cross3(p, f2, x);
          for (int k = 0; k < 3; ++k) { B[k][c] = f2[k]; B[3 + k][c] = t2[k] + x[k]; }
        }
      }
    }
  }
  free(st);
  return 0;
}

/* inverse dynamics with the model's own inertias; tau has nv entries */
int orc_rnea(const orc_model *m, const double *q, const double *qd, const double *qdd, double *tau) {
  const int n = m->njoints;
  orc_jstate *st = (orc_jstate *)malloc(sizeof(orc_jstate) * n);
  double *f = (double *)calloc((size_t)6 * n, sizeof(double));
  if (!st || !f) { free(st); free(f); return -1; }
  forward_pass(m, q, qd, qdd, st);
  for (int i = 1; i < n; ++i) {
    /* f = I a + v x* (I v), spatial inertia (mass, c, I_C) */
    const double mass = m->mass[i], *c = m->lever + 3 * i, *I = m->inertia + 6 * i;
    const double Ic[9] = {I[0], I[1], I[3], I[1], I[2], I[4], I[3], I[4], I[5]};
    double h[6], fa[6];
    for (int pass = 0; pass < 2; ++pass) {
      const double *mv = pass ? st[i].a : st[i].v;
      double *o = pass ? fa : h;
      double t[3], u[3];
      cross3(mv + 3, c, t);                       /* w x c */
      for (int k = 0; k < 3; ++k) o[k] = mass * (mv[k] + t[k]);
      matvec3(Ic, mv + 3, u);
      cross3(c, o, t);                            /* c x f_lin */
      for (int k = 0; k < 3; ++k) o[3 + k] = u[k] + t[k];
    }
    double t1[3], t2[3], t3[3];
    cross3(st[i].v + 3, h, t1);                   /* w x h_lin */
    cross3(st[i].v + 3, h + 3, t2);               /* w x h_ang */
    cross3(st[i].v, h, t3);                       /* v x h_lin */
    for (int k = 0; k < 3; ++k) { f[6 * i + k] = fa[k] + t1[k]; f[6 * i + 3 + k] = fa[3 + k] + t2[k] + t3[k]; }
  }
  for (int i = n - 1; i >= 1; --i) {
    const int iv = m->idx_v[i];
    const double *ax = m->axis + 3 * i, *fi = f + 6 * i;
    switch (m->jtype[i]) {
      case JT_REVOLUTE: case JT_REVOLUTE_UNB: tau[iv] = ax[0] * fi[3] + ax[1] * fi[4] + ax[2] * fi[5]; break;
      case JT_PRISMATIC: tau[iv] = ax[0] * fi[0] + ax[1] * fi[1] + ax[2] * fi[2]; break;
      case JT_FREEFLYER: for (int k = 0; k < 6; ++k) tau[iv + k] = fi[k]; break;
      default: break;
    }
    const int par = m->parent[i];
    if (par > 0) {
      double f2[3], t2[3], x[3];
      matvec3(st[i].R, fi, f2); matvec3(st[i].R, fi + 3, t2); cross3(st[i].p, f2, x);
      for (int k = 0; k < 3; ++k) { f[6 * par + k] += f2[k]; f[6 * par + 3 + k] += t2[k] + x[k]; }
    }
  }
  free(st); free(f);
  return 0;
}

/*
 * Tall regressor in the reference layout, restating
 * /root/reference/src/figaroh/tools/regressor.py:20-194 (build_regressor_basic).
 *   mode 0 (joint torques): rows j*N+i, j<nv; requires nv == njoints-1.
 *   mode 1 (external wrench): rows j*N+i, j<6; ft_mask bit j enables row block j;
 *     only bodies with body_has_mass[idx] != 0 are copied (regressor.py:36-39,99-138).
 * flags: bit0 friction, bit1 actuator inertia, bit2 joint offset.
 * W is (rows x 14*nb) row-major with FIGAROH's per-body column order
 *   [Ixx Ixy Ixz Iyy Iyz Izz mx my mz m Ia fv fs off].
 */
static const int PIN2FIG[10] = {9, 6, 7, 8, 0, 1, 3, 2, 4, 5}; /* pinocchio col k -> figaroh col */

int orc_build_regressor_basic(const orc_model *m, const double *q, const double *qd, const double *qdd, long N,
                              int mode, int flags, int ft_mask, double *W) {
  const int nb = m->njoints - 1, nv = m->nv, ncol = 14 * nb, npin = 10 * nb;
  const int nrow_blocks = mode == 0 ? nv : 6;
  if (mode == 0 && nv != nb) return -2; /* regressor.py:53 broadcasts 10n into 10nv */
  double *Wt = (double *)malloc(sizeof(double) * (size_t)nv * npin);
  if (!Wt) return -1;
  memset(W, 0, sizeof(double) * (size_t)nrow_blocks * N * ncol);
  for (long i = 0; i < N; ++i) {
    const double *qi = q + (size_t)i * m->nq, *vi = qd + (size_t)i * nv, *ai = qdd + (size_t)i * nv;
    orc_joint_torque_regressor(m, qi, vi, ai, Wt);
    for (int j = 0; j < nrow_blocks; ++j) {
      double *row = W + ((size_t)j * N + i) * ncol;
      if (mode == 0 || (ft_mask >> j) & 1) {
        for (int b = 0; b < nb; ++b) {
          if (mode == 1 && m->mass[b + 1] == 0.0) continue;
          for (int k = 0; k < 10; ++k) row[14 * b + PIN2FIG[k]] = Wt[(size_t)j * npin + 10 * b + k];
        }
      }
      if (mode == 0) {
        if (flags & 1) { row[14 * j + 11] = vi[j]; row[14 * j + 12] = (vi[j] > 0) - (vi[j] < 0); }
        if (flags & 2) row[14 * j + 10] = ai[j];
        if (flags & 4) row[14 * j + 13] = 1.0;
      } else {
        for (int b = 0; b < nb; ++b) { /* regressor.py:142-169: body index reused as velocity index */
          if (flags & 1) { row[14 * b + 11] = vi[b]; row[14 * b + 12] = (vi[b] > 0) - (vi[b] < 0); }
          if (flags & 2) row[14 * b + 10] = ai[b];
          if (flags & 4) row[14 * b + 13] = 1.0;
        }
      }
    }
  }
  free(Wt);
  return 0;
}
