// fg_tileqr.cuh - kernel (1)+(2), third generation.
//
// What changed against fg_blockqr.cuh (profiles/ncu_bq_r01_v2.md: 5 780 warp instructions per
// sample, a third of them in a divergent column-per-lane regressor generator, the rest in 8-row
// panel chains that are paid once per 4 samples):
//
//  * Regressor generation is SAMPLE-PER-THREAD: every lane runs the same straight-line forward
//    pass on its own sample (no divergence, model data read through the constant cache) and
//    writes the entries W[g, body b, :] = s^T B_b, where s is the motion axis of row group g
//    expressed in body b's frame (one 6-vector transform per (g, b) pair instead of carrying ten
//    wrench columns up the chain).  This is what pinocchio::computeJointTorqueRegressor computes
//    (call sites regressor.py:49,93 of the reference), evaluated body by body.
//  * A tile is 32 samples.  Its rows are grouped BY ROW GROUP (joint): row block g holds the 32
//    rows of joint g, so every (row block, 8-column block) pair is either structurally zero - and
//    then neither stored nor touched - or a dense 32 x 8 block in shared memory.
//  * Columns are ordered [private blocks | general blocks | rhs].  The Ia/fv/fs/off columns of
//    joint g are non-zero only in row block g ("private"): they are eliminated first, by six
//    independent warp-local panels.  The general columns keep the reference order, in which row
//    block g starts at the columns of body g (a staircase): panel k only involves the row blocks
//    that are non-zero in it.  A host-side symbolic pass (fg_tq.cu) turns any kinematic tree and
//    column selection into the block lists the kernel walks; nothing is specialised at compile time.
//  * Panels are factored by one warp with one ROW PER LANE (8 reflectors per 32 samples instead
//    of per 4), the compact-WY factor T comes from T^-1 = diag(1/tau) + striu(V^T V), and all
//    trailing work  Y^T = R_kj^T + B_j^T V,  W^T = Y^T T,  B_j^T -= W^T V^T  runs on DMMA
//    (mma.sync.m8n8k4.f64) with operands read from the swizzled shared-memory tile.
//  * R lives in a per-CTA global scratch (a few tens of KB, L2 resident): each 8 x 8 block of it is
//    read and written once per tile, so shared memory is spent on samples in flight instead.
//
// The state this kernel produces is the R factor of [W P | tau] for the column permutation P the
// host chose; fg_api.cu keeps P with the state (fg_state_perm).
#pragma once
#include "fg_device.cuh"
#include "fg_tq.h"

namespace fg {

__device__ __forceinline__ void tq_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Tile block = 8 columns x 32 rows, column-major with the row index XOR-swizzled per column so that
// every access pattern of the kernel (lane = row; lane = (column, row mod 4); lane = (row mod 8,
// column pair)) hits 16 distinct 8-byte banks per half warp.
__host__ __device__ __forceinline__ int tq_sw(int c) {
  const int f = (c + (c >> 2)) & 3;
  return (f << 2) | (f & 1);
}
__host__ __device__ __forceinline__ int tq_idx(int c, int s) { return c * TQ_S + (s ^ tq_sw(c)); }

// index of R block (k, j), j >= k, in the per-CTA block-upper-triangular scratch (64 doubles each, [refl][col])
__host__ __device__ __forceinline__ int tq_rblk(int ncb, int k, int j) { return k * ncb - (k * (k - 1)) / 2 + (j - k); }

// named barriers (ids 1..4; id 0 is __syncthreads): producer arrives, consumers sync
__device__ __forceinline__ void tq_bar_arrive(int id, int count) {
  __threadfence_block();
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void tq_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

__device__ __forceinline__ double tq_allreduce(double x) {
  x += __shfl_xor_sync(0xffffffffu, x, 16);
  x += __shfl_xor_sync(0xffffffffu, x, 8);
  x += __shfl_xor_sync(0xffffffffu, x, 4);
  x += __shfl_xor_sync(0xffffffffu, x, 2);
  x += __shfl_xor_sync(0xffffffffu, x, 1);
  return x;
}

// ---------------------------------------------------------------------------------------------
// regressor entries of (row group g, body): s = oM_body^-1 o A,  value[k] = s^T B[:, k]
// A: motion axis of the row group in the world frame [lin; ang]; oR/op: world pose of the body;
// om/al/ap: angular velocity, angular acceleration and classical linear acceleration (body frame).
__device__ __forceinline__ void tq_emit(const TqParams& P, double* tile, int g, int body, const double* A,
                                        const double* oR, const double* op, const double* om, const double* al,
                                        const double* ap, bool valid, int lane) {
  double t[3], sl[3], sa[3];
  cross3(op, A + 3, t);
  for (int k = 0; k < 3; ++k) t[k] = A[k] - t[k];
  for (int k = 0; k < 3; ++k) {
    sl[k] = oR[k] * t[0] + oR[3 + k] * t[1] + oR[6 + k] * t[2];
    sa[k] = oR[k] * A[3] + oR[3 + k] * A[4] + oR[6 + k] * A[5];
  }
  double val[10];
  val[0] = sl[0] * ap[0] + sl[1] * ap[1] + sl[2] * ap[2];
  const double ww = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
  const double so = sl[0] * om[0] + sl[1] * om[1] + sl[2] * om[2];
  double x1[3], x2[3], h[3];
  cross3(sl, al, x1);
  cross3(ap, sa, x2);
  for (int j = 0; j < 3; ++j) val[1 + j] = x1[j] + so * om[j] - ww * sl[j] + x2[j];
  cross3(sa, om, h);
  val[4] = sa[0] * al[0] + h[0] * om[0];                                   // Ixx
  val[5] = sa[0] * al[1] + sa[1] * al[0] + h[0] * om[1] + h[1] * om[0];   // Ixy
  val[6] = sa[1] * al[1] + h[1] * om[1];                                   // Iyy
  val[7] = sa[0] * al[2] + sa[2] * al[0] + h[0] * om[2] + h[2] * om[0];   // Ixz
  val[8] = sa[1] * al[2] + sa[2] * al[1] + h[1] * om[2] + h[2] * om[1];   // Iyz
  val[9] = sa[2] * al[2] + h[2] * om[2];                                   // Izz
  if (valid) {
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const int slot = P.islot[body][k];
      if (slot >= 0) tile[P.blkoff[g][slot >> 3] + (slot & 7) * TQ_S + (lane ^ P.swz[slot & 7])] = val[k];
    }
  }
}

// Forward pass of one sample per lane; warp w emits the row groups w and w + 4.
static __device__ __noinline__ void tq_generate(const TqParams& P, double* tile, const double* scr, const DevJoint* Js,
                                         const double* __restrict__ q, const double* __restrict__ qd,
                                         const double* __restrict__ qdd, const double* __restrict__ tau,
                                         int64_t ld_tau, int64_t s0, int64_t N, int warp, int lane) {
  const int64_t smp = s0 + lane;
  const bool valid = smp < N;
  const int64_t sc = valid ? smp : N - 1;
  const double* qs = q + sc * P.nq;
  const double* vs = qd + sc * P.nv;
  const double* as = qdd + sc * P.nv;
  const int g0 = warp, g1 = warp + TQ_WARPS;
  double A0[6] = {0, 0, 0, 0, 0, 0}, A1[6] = {0, 0, 0, 0, 0, 0};
  // state of the previously processed joint: world pose, body-frame velocity and acceleration
  double oR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, op[3] = {0, 0, 0};
  double v[6] = {0, 0, 0, 0, 0, 0}, a[6] = {0, 0, 0, 0, 0, 0};
  const unsigned mine = (1u << g0) | (g1 < P.ms ? (1u << g1) : 0u);
  // the (q, dq, ddq) entries of the NEXT joint are requested one joint ahead: their HBM/L2 round trip
  // overlaps the arithmetic of the current joint
  double nq0 = __ldg(qs + Js[P.proc[0]].idx_q), nv0 = __ldg(vs + Js[P.proc[0]].idx_v),
         na0 = __ldg(as + Js[P.proc[0]].idx_v);
  for (int t = 0; t < P.nproc; ++t) {
    const int i = P.proc[t];
    const DevJoint& jt = Js[i];
    const double cq0 = nq0, cv0 = nv0, ca0 = na0;
    if (t + 1 < P.nproc) {
      const DevJoint& jn = Js[P.proc[t + 1]];
      nq0 = __ldg(qs + jn.idx_q); nv0 = __ldg(vs + jn.idx_v); na0 = __ldg(as + jn.idx_v);
    }
    if (jt.parent == 0) {  // restart from the universe (the host only admits chains hanging off it)
      oR[0] = 1; oR[1] = 0; oR[2] = 0; oR[3] = 0; oR[4] = 1; oR[5] = 0; oR[6] = 0; oR[7] = 0; oR[8] = 1;
      op[0] = op[1] = op[2] = 0.0;
      for (int k = 0; k < 6; ++k) v[k] = 0.0;
      a[0] = -P.gravity[0]; a[1] = -P.gravity[1]; a[2] = -P.gravity[2];
      a[3] = a[4] = a[5] = 0.0;
    }
    double Rj[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pj[3] = {0, 0, 0};
    double vj[6] = {0, 0, 0, 0, 0, 0}, aj[6] = {0, 0, 0, 0, 0, 0};
    const int iq = jt.idx_q, iv = jt.idx_v;
    if (jt.jtype == 1) {
      const double c = scr[(2 * t) * TQ_S + lane], s = scr[(2 * t + 1) * TQ_S + lane];
      axis_rot(jt.axis, c, s, Rj);
      const double qd1 = cv0, qdd1 = ca0;
      for (int k = 0; k < 3; ++k) { vj[3 + k] = jt.axis[k] * qd1; aj[3 + k] = jt.axis[k] * qdd1; }
    } else if (jt.jtype == 2) {
      axis_rot(jt.axis, cq0, __ldg(qs + iq + 1), Rj);
      const double qd1 = cv0, qdd1 = ca0;
      for (int k = 0; k < 3; ++k) { vj[3 + k] = jt.axis[k] * qd1; aj[3 + k] = jt.axis[k] * qdd1; }
    } else if (jt.jtype == 3) {
      const double qq = cq0, qd1 = cv0, qdd1 = ca0;
      for (int k = 0; k < 3; ++k) { pj[k] = jt.axis[k] * qq; vj[k] = jt.axis[k] * qd1; aj[k] = jt.axis[k] * qdd1; }
    } else {  // free-flyer
      quat_rot(__ldg(qs + iq + 3), __ldg(qs + iq + 4), __ldg(qs + iq + 5), __ldg(qs + iq + 6), Rj);
      for (int k = 0; k < 3; ++k) pj[k] = __ldg(qs + iq + k);
      for (int k = 0; k < 6; ++k) { vj[k] = __ldg(vs + iv + k); aj[k] = __ldg(as + iv + k); }
    }
    // liMi = placement * Mj
    double R[9], p[3];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c)
        R[3 * r + c] = jt.plcR[3 * r] * Rj[c] + jt.plcR[3 * r + 1] * Rj[3 + c] + jt.plcR[3 * r + 2] * Rj[6 + c];
      p[r] = jt.plcR[3 * r] * pj[0] + jt.plcR[3 * r + 1] * pj[1] + jt.plcR[3 * r + 2] * pj[2] + jt.plcp[r];
    }
    // v = vJ + liMi.actInv(v_parent);  a = v x vJ + S qdd + liMi.actInv(a_parent)
    double vn[6], an[6], tt[3];
    cross3(p, v + 3, tt);
    for (int k = 0; k < 3; ++k) tt[k] = v[k] - tt[k];
    for (int k = 0; k < 3; ++k) {
      vn[k] = vj[k] + R[k] * tt[0] + R[3 + k] * tt[1] + R[6 + k] * tt[2];
      vn[3 + k] = vj[3 + k] + R[k] * v[3] + R[3 + k] * v[4] + R[6 + k] * v[5];
    }
    double l1[3], l2[3], l3[3];
    cross3(vn + 3, vj, l1);
    cross3(vn, vj + 3, l2);
    cross3(vn + 3, vj + 3, l3);
    cross3(p, a + 3, tt);
    for (int k = 0; k < 3; ++k) tt[k] = a[k] - tt[k];
    for (int k = 0; k < 3; ++k) {
      an[k] = l1[k] + l2[k] + aj[k] + R[k] * tt[0] + R[3 + k] * tt[1] + R[6 + k] * tt[2];
      an[3 + k] = l3[k] + aj[3 + k] + R[k] * a[3] + R[3 + k] * a[4] + R[6 + k] * a[5];
    }
    // world pose oMi = oM_parent * liMi
    double nR[9], np[3];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) nR[3 * r + c] = oR[3 * r] * R[c] + oR[3 * r + 1] * R[3 + c] + oR[3 * r + 2] * R[6 + c];
      np[r] = oR[3 * r] * p[0] + oR[3 * r + 1] * p[1] + oR[3 * r + 2] * p[2] + op[r];
    }
    for (int k = 0; k < 9; ++k) oR[k] = nR[k];
    for (int k = 0; k < 3; ++k) op[k] = np[k];
    for (int k = 0; k < 6; ++k) { v[k] = vn[k]; a[k] = an[k]; }
    // world-frame motion axes of the row groups carried by this joint
    if (g0 < P.ms && P.rowjoint[g0] == i) {
      const double* S = P.rowS[g0];
      for (int k = 0; k < 3; ++k) {
        A0[3 + k] = oR[3 * k] * S[3] + oR[3 * k + 1] * S[4] + oR[3 * k + 2] * S[5];
        A0[k] = oR[3 * k] * S[0] + oR[3 * k + 1] * S[1] + oR[3 * k + 2] * S[2];
      }
      cross3(op, A0 + 3, tt);
      for (int k = 0; k < 3; ++k) A0[k] += tt[k];
    }
    if (g1 < P.ms && P.rowjoint[g1] == i) {
      const double* S = P.rowS[g1];
      for (int k = 0; k < 3; ++k) {
        A1[3 + k] = oR[3 * k] * S[3] + oR[3 * k + 1] * S[4] + oR[3 * k + 2] * S[5];
        A1[k] = oR[3 * k] * S[0] + oR[3 * k + 1] * S[1] + oR[3 * k + 2] * S[2];
      }
      cross3(op, A1 + 3, tt);
      for (int k = 0; k < 3; ++k) A1[k] += tt[k];
    }
    const unsigned am = (unsigned)P.anc[i] & mine;
    if (am) {
      double ap[3];
      cross3(v + 3, v, ap);  // classical acceleration a' = a_lin + w x v_lin
      for (int k = 0; k < 3; ++k) ap[k] += a[k];
      if ((am >> g0) & 1u) tq_emit(P, tile, g0, i, A0, oR, op, v + 3, a + 3, ap, valid, lane);
      if (g1 < P.ms && ((am >> g1) & 1u)) tq_emit(P, tile, g1, i, A1, oR, op, v + 3, a + 3, ap, valid, lane);
    }
  }
  // columns read straight off the inputs (unrolled so that the loads of several columns are in flight together)
#pragma unroll 4
  for (int t = warp; t < P.nsimple; t += TQ_WARPS) {
    const TqSimple sc1 = P.simple[t];
    const int j = sc1.slot >> 3, ix = tq_idx(sc1.slot & 7, lane);
    if (sc1.kind == COL_RHS) {
      for (int g = 0; g < P.ms; ++g) {
        const int off = P.blkoff[g][j];
        if (off >= 0) tile[off + ix] = (valid && tau != nullptr) ? __ldg(tau + (int64_t)g * ld_tau + smp) : 0.0;
      }
    } else if (sc1.kind == COL_EXTRA) {
      const double vv = __ldg(vs + sc1.vidx), aa = __ldg(as + sc1.vidx);
      double val = sc1.src == SRC_QDD ? aa : (sc1.src == SRC_QD ? vv : (sc1.src == SRC_SIGN_QD ? sgn(vv) : 1.0));
      if (!valid) val = 0.0;
      if (sc1.grp >= 0) {
        const int off = P.blkoff[sc1.grp][j];
        if (off >= 0) tile[off + ix] = val;
      } else {
        for (int g = 0; g < P.ms; ++g) {
          const int off = P.blkoff[g][j];
          if (off >= 0) tile[off + ix] = val;
        }
      }
    } else if (sc1.kind == COL_COUPLING) {
      // add_coupling_TX40 (regressor.py:216-225): row block 4 <- [a5, v5, sign(v4+v5)], 5 <- [a4, v4, sign(v4+v5)]
      const double v4 = __ldg(vs + 4), v5 = __ldg(vs + 5), a4 = __ldg(as + 4), a5 = __ldg(as + 5);
      const double sg = sgn(v4 + v5);
      for (int g = 0; g < P.ms; ++g) {
        const int off = P.blkoff[g][j];
        const int rb = P.rowblock[g];
        if (off >= 0 && (rb == 4 || rb == 5)) {
          double val = rb == 4 ? (sc1.src == 0 ? a5 : (sc1.src == 1 ? v5 : sg)) : (sc1.src == 0 ? a4 : (sc1.src == 1 ? v4 : sg));
          tile[off + ix] = valid ? val : 0.0;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Householder panel, one warp, one row of every row block per lane:
//   x[gi][c] = entry (row = lane of row block gi, column c); Rkk: the diagonal R block (row stride ldr,
//   shared memory); sc: 16 doubles of shared scratch (tau_j, 1 / (alpha_j - beta_j)).
// On exit x holds the NORMALISED reflectors v = x / (alpha - beta) (LAPACK dlarfg convention), Rkk the
// updated block, sc[0..7] the taus (0 for a skipped / padding column).
template <int NG, int W>
__device__ __forceinline__ void tq_panel_factor(double (&x)[NG][8], const double* Rin, double* Rout, int ldr, double* sc,
                                                int lane) {
  // Row blocks that do not take part and padding columns hold zeros in x, so every loop has compile-time
  // bounds (a zero column gives sigma == 0 and is skipped).  The diagonal block is read from Rin and the
  // updated rows are written to Rout by lane 0: no read-after-write hazard between lanes.
#pragma unroll
  for (int j = 0; j < W; ++j) {
    double tauh = 0.0, inv = 0.0;
    double d[8];
#pragma unroll
    for (int c = j; c < W; ++c) {
      double acc = 0.0;
#pragma unroll
      for (int gi = 0; gi < NG; ++gi) acc = fma(x[gi][j], x[gi][c], acc);
      d[c] = acc;
    }
    // butterfly all-reduce of d[j..W-1], level by level for all values at once: the shuffles of one level
    // are independent, so their latencies overlap (one value after the other costs 5 round trips EACH)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      double t[8];
#pragma unroll
      for (int c = j; c < W; ++c) t[c] = __shfl_xor_sync(0xffffffffu, d[c], off);
#pragma unroll
      for (int c = j; c < W; ++c) d[c] += t[c];
    }
    const double sigma = d[j];
    const double alpha = Rin[j * ldr + j];
    double beta = alpha;
    if (sigma != 0.0) {  // uniform: the butterfly leaves the same bits in every lane
      const double mu = fma(alpha, alpha, sigma);
      const double rs = rsqrt(mu);
      const double nrm = mu * rs;
      beta = alpha >= 0.0 ? -nrm : nrm;
      const double den = alpha - beta;
      inv = __drcp_rn(den);
      tauh = fabs(den) * rs;  // (beta - alpha) / beta
    }
    if (lane == 0) { Rout[j * ldr + j] = beta; sc[j] = tauh; }
#pragma unroll
    for (int c = j + 1; c < W; ++c) {
      const double rho = Rin[j * ldr + c];
      const double w = tauh * fma(d[c], inv, rho);  // tau * v^T [rho; b]  (0 when the reflector is skipped)
      const double coef = w * inv;
#pragma unroll
      for (int gi = 0; gi < NG; ++gi) x[gi][c] = fma(-coef, x[gi][j], x[gi][c]);
      if (lane == 0) Rout[j * ldr + c] = rho - w;
    }
#pragma unroll
    for (int gi = 0; gi < NG; ++gi) x[gi][j] *= inv;  // H = I (dlarfg with xnorm == 0) leaves v = 0
  }
  __syncwarp();
}

// Compact-WY factor of the panel's reflectors: T^-1 = diag(1/tau) + striu(V^T V).  The normalised
// reflectors are already back in the tile (blocks V_g, row groups in gmask); V^T V = sum_g V_g^T V_g is a
// DMMA whose A and B fragments are the same register.  sc[0..7]: the taus; TS (64 doubles, shared)
// receives T row-major.  KW = stored width of the panel block.
template <int KW>
__device__ __forceinline__ void tq_panel_T(const TqParams& P, const double* tile, int k, unsigned gmask,
                                           const double* sc, double* TS, const int (&oy)[TQ_S / 4], int lane) {
  const int cq = lane >> 2, rq = lane & 3;
  double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
  for (unsigned gm = gmask; gm; gm &= gm - 1) {
    const double* Xk = tile + P.blkoff[__ffs(gm) - 1][k];
#pragma unroll
    for (int kk = 0; kk < TQ_S / 4; kk += 2) {
      const double v0 = (KW == 8 || cq < KW) ? Xk[oy[kk]] : 0.0;
      const double v1 = (KW == 8 || cq < KW) ? Xk[oy[kk + 1]] : 0.0;
      tq_dmma(a0, a1, v0, v0);
      tq_dmma(b0, b1, v1, v1);
    }
  }
  a0 += b0;
  a1 += b1;
  // lane holds (V^T V)[cq][2 rq + e]; TS[m][i] = v_i . v_m for i < m, TS[m][m] = tau_m
  if (2 * rq < cq) TS[cq * 8 + 2 * rq] = a0;
  if (2 * rq + 1 < cq) TS[cq * 8 + 2 * rq + 1] = a1;
  if (lane < 8) TS[lane * 8 + lane] = sc[lane];
  __syncwarp();
  // column cq of T (lanes 0..7); S is read in place and T written over it after every lane is done
  double t[8];
  if (lane < 8) {
    const int cq = lane;
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0;
#pragma unroll
    for (int m = 7; m >= 0; --m) {
      const double taum = TS[m * 8 + m];
      const double tm = (m == cq) ? taum : (m < cq ? -taum * acc[m] : 0.0);
      t[m] = tm;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < m) acc[i] = fma(TS[m * 8 + i], tm, acc[i]);
    }
  }
  __syncwarp();
  if (lane < 8) {
#pragma unroll
    for (int m = 0; m < 8; ++m) TS[m * 8 + lane] = t[m];
  }
  __syncwarp();
}

// Per-lane offsets inside a tile block of the three access patterns of the trailing update; they
// depend on the lane only, so a warp computes them once.
struct TqOffs {
  int oy[TQ_S / 4];       // (row 4kk + rq, column cq): A / B fragments of the row contraction
  int oc0[TQ_S / 8], oc1[TQ_S / 8];  // (rows 8nn + 2rq, +1; column cq): accumulator fragments of the update
  int ob0[TQ_S / 8], ob1[TQ_S / 8];  // (row 8nn + cq; columns 2rq, 2rq + 1): V^T as B fragment of the update
};
__device__ __forceinline__ void tq_make_offs(TqOffs& o, int lane) {
  const int cq = lane >> 2, rq = lane & 3;
#pragma unroll
  for (int kk = 0; kk < TQ_S / 4; ++kk) o.oy[kk] = tq_idx(cq, 4 * kk + rq);
#pragma unroll
  for (int nn = 0; nn < TQ_S / 8; ++nn) {
    o.oc0[nn] = tq_idx(cq, 8 * nn + 2 * rq);
    o.oc1[nn] = tq_idx(cq, 8 * nn + 2 * rq + 1);
    o.ob0[nn] = tq_idx(2 * rq, 8 * nn + cq);
    o.ob1[nn] = tq_idx(2 * rq + 1, 8 * nn + cq);
  }
}

// Trailing update of column block j by panel k for the row groups in gmask (one warp):
//   Y^T = R_kj^T + sum_g B_gj^T V_g ;  W^T = Y^T T ;  R_kj^T -= W^T ;  B_gj^T -= W^T V_g^T
// KW = stored width of the panel block (8, or 4 for a private block whose columns 4..7 are padding).
template <int KW>
__device__ __forceinline__ void tq_trailing(const TqParams& P, double* tile, double* Rb, const double* TS, int k, int j,
                                            unsigned gmask, const TqOffs& o, int lane) {
  const int cq = lane >> 2, rq = lane & 3;
  const double t0 = TS[(2 * rq) * 8 + cq], t1 = TS[(2 * rq + 1) * 8 + cq];
  const double r0 = Rb[(2 * rq) * 8 + cq], r1 = Rb[(2 * rq + 1) * 8 + cq];  // L2 round trip hidden by the DMMA chain
  double y0 = 0.0, y1 = 0.0, z0 = 0.0, z1 = 0.0;  // two accumulator pairs: two independent DMMA chains
  for (unsigned gm = gmask; gm; gm &= gm - 1) {
    const int g = __ffs(gm) - 1;
    const double* Bj = tile + P.blkoff[g][j];
    const double* Xk = tile + P.blkoff[g][k];
#pragma unroll
    for (int kk = 0; kk < TQ_S / 4; kk += 2) {
      const double b0 = (KW == 8 || cq < KW) ? Xk[o.oy[kk]] : 0.0;
      const double b1 = (KW == 8 || cq < KW) ? Xk[o.oy[kk + 1]] : 0.0;
      tq_dmma(y0, y1, Bj[o.oy[kk]], b0);
      tq_dmma(z0, z1, Bj[o.oy[kk + 1]], b1);
    }
  }
  y0 += z0 + r0;
  y1 += z1 + r1;
  double w0 = 0.0, w1 = 0.0;
  tq_dmma(w0, w1, y0, t0);
  tq_dmma(w0, w1, y1, t1);
  Rb[(2 * rq) * 8 + cq] = r0 - w0;
  Rb[(2 * rq + 1) * 8 + cq] = r1 - w1;
  const double nw0 = -w0, nw1 = -w1;
  for (unsigned gm = gmask; gm; gm &= gm - 1) {
    const int g = __ffs(gm) - 1;
    double* Bj = tile + P.blkoff[g][j];
    const double* Xk = tile + P.blkoff[g][k];
#pragma unroll
    for (int nn = 0; nn < TQ_S / 8; ++nn) {
      double c0 = Bj[o.oc0[nn]], c1 = Bj[o.oc1[nn]];
      const double b0 = (KW == 8 || 2 * rq < KW) ? Xk[o.ob0[nn]] : 0.0;
      const double b1 = (KW == 8 || 2 * rq + 1 < KW) ? Xk[o.ob1[nn]] : 0.0;
      tq_dmma(c0, c1, nw0, b0);
      tq_dmma(c0, c1, nw1, b1);
      Bj[o.oc0[nn]] = c0;
      Bj[o.oc1[nn]] = c1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <int MAXG>
__global__ void __launch_bounds__(TQ_WARPS * 32, 3)
tq_accumulate_kernel(const __grid_constant__ TqParams P, const double* __restrict__ q, const double* __restrict__ qd,
                     const double* __restrict__ qdd, const double* __restrict__ tau, int64_t ld_tau, int64_t N,
                     double* __restrict__ r_all, int64_t r_doubles, double* __restrict__ parts, int phase) {
  extern __shared__ __align__(16) double tq_smem[];
  double* tile = tq_smem;
  double* scr = tq_smem + P.tile_doubles;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* Rg = r_all + (size_t)blockIdx.x * r_doubles;
  // phase bit 0: start from R = 0 (else continue the factor a previous launch left in the scratch);
  // phase bit 1: write the packed factor to parts[] at the end
  if (phase & 1)
    for (int i = threadIdx.x; i < r_doubles; i += blockDim.x) Rg[i] = 0.0;
  const int ncb = P.ncb;
  // Roles in the general phase: ONE panel warp per CTA (CTAs that share an SM - blockIdx.x = sm + i * sm_count -
  // pick different warps, i.e. different SM sub-partitions) and three workers that own the column blocks
  // round-robin and apply the trailing updates in the shadow of the next panel (look-ahead).
  const int pw = ((int)blockIdx.x / P.sm_count) & (TQ_WARPS - 1);
  const int wi = (warp - pw - 1) & (TQ_WARPS - 1);  // 0..2 for the workers, 3 for the panel warp
  DevJoint* Js = reinterpret_cast<DevJoint*>(scr + 2 * TQ_S * P.nproc);
  TqOffs offs;
  tq_make_offs(offs, lane);
  const int64_t ntiles = (N + TQ_S - 1) / TQ_S;
  for (int64_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x) {
    const int64_t s0 = tl * TQ_S;
    __syncthreads();
    // ---- clear the tile; joint constants and the sines / cosines of the revolute joints into the scratch
    {
      double2* t2 = reinterpret_cast<double2*>(tile);
      const int n2 = P.tile_doubles / 2;
      for (int i = threadIdx.x; i < n2; i += blockDim.x) t2[i] = make_double2(0.0, 0.0);
      for (int i = threadIdx.x; i < P.nj * (int)(sizeof(DevJoint) / 8); i += blockDim.x)
        reinterpret_cast<double*>(Js)[i] = reinterpret_cast<const double*>(P.joints)[i];
      const int64_t smp = s0 + lane;
      const int64_t sc = smp < N ? smp : N - 1;
      for (int t = warp; t < P.nproc; t += TQ_WARPS) {
        const DevJoint& jt = P.joints[P.proc[t]];
        if (jt.jtype == 1) {
          double s, c;
          sincos(__ldg(q + sc * P.nq + jt.idx_q), &s, &c);
          scr[(2 * t) * TQ_S + lane] = c;
          scr[(2 * t + 1) * TQ_S + lane] = s;
        }
      }
    }
    __syncthreads();
    tq_generate(P, tile, scr, Js, q, qd, qdd, tau, ld_tau, s0, N, warp, lane);
    __syncthreads();
    // ---- phase A: private panels, one warp each, no CTA-wide synchronisation
    {
      double* TS = scr + warp * 112;
      double* Ree = TS + 64;
      for (int e = warp; e < P.npriv; e += TQ_WARPS) {
        const unsigned gm = P.gmask[e];
        const int g = __ffs(gm) - 1;
        double* Xe = tile + P.blkoff[g][e];
        double* Rd = Rg + (size_t)tq_rblk(ncb, e, e) * 64;
        if (lane < 16) Ree[lane] = Rd[(lane >> 2) * 8 + (lane & 3)];
        double x[1][8];
#pragma unroll
        for (int c = 0; c < 8; ++c) x[0][c] = c < 4 ? Xe[tq_idx(c, lane)] : 0.0;
        __syncwarp();
        tq_panel_factor<1, 4>(x, Ree, Ree + 16, 4, Ree + 32, lane);
#pragma unroll
        for (int c = 0; c < 4; ++c) Xe[tq_idx(c, lane)] = x[0][c];
        __syncwarp();
        tq_panel_T<4>(P, tile, e, gm, Ree + 32, TS, offs.oy, lane);
        if (lane < 16 && (lane >> 2) <= (lane & 3)) Rd[(lane >> 2) * 8 + (lane & 3)] = Ree[16 + lane];
        __syncwarp();
        for (unsigned tm = P.trmask[e]; tm; tm &= tm - 1) {
          const int j = __ffs(tm) - 1;
          tq_trailing<4>(P, tile, Rg + (size_t)tq_rblk(ncb, e, j) * 64, TS, e, j, gm, offs, lane);
        }
        __syncwarp();
      }
    }
    __syncthreads();
    // ---- phase B: general panels with look-ahead.  Barrier 1 + (k & 1): "panel k is factored" (panel warp
    // arrives, workers wait); barrier 3 + (k & 1): "block k + 1 carries the updates of every panel <= k" (its
    // owner arrives, the panel warp waits).  Scratch: one T per worker, the diagonal block (in / out) and the
    // taus of panel k (double-buffered by the parity of k).
    {
      double* TSw = scr + 64 * wi;       // workers: wi = 0..2
      double* Rkk = scr + 192;           // [192, 256) in, [256, 320) out
      double* scb = scr + 320;           // [320, 336): taus of panel k at scb + 8 * (k & 1)
      if (wi == 3) {
        for (int k = P.npriv; k < ncb; ++k) {
          if (k > P.npriv) tq_bar_sync(3 + ((k - 1) & 1), 64);
          const unsigned gm = P.gmask[k];
          if (gm) {
            double* Rd = Rg + (size_t)tq_rblk(ncb, k, k) * 64;
            Rkk[lane] = Rd[lane];
            Rkk[lane + 32] = Rd[lane + 32];
            int gl[MAXG];
            int ng = 0;
#pragma unroll
            for (int gi = 0; gi < MAXG; ++gi) gl[gi] = 0;
            {
              unsigned m2 = gm;
#pragma unroll
              for (int gi = 0; gi < MAXG; ++gi)
                if (m2) { gl[gi] = P.blkoff[__ffs(m2) - 1][k]; m2 &= m2 - 1; ++ng; }
            }
            double x[MAXG][8];
#pragma unroll
            for (int gi = 0; gi < MAXG; ++gi)
#pragma unroll
              for (int c = 0; c < 8; ++c) x[gi][c] = gi < ng ? tile[gl[gi] + tq_idx(c, lane)] : 0.0;
            __syncwarp();
            tq_panel_factor<MAXG, 8>(x, Rkk, Rkk + 64, 8, scb + 8 * (k & 1), lane);
#pragma unroll
            for (int gi = 0; gi < MAXG; ++gi)
              if (gi < ng) {
#pragma unroll
                for (int c = 0; c < 8; ++c) tile[gl[gi] + tq_idx(c, lane)] = x[gi][c];
              }
            if ((lane >> 3) <= (lane & 7)) Rd[lane] = Rkk[64 + lane];            // rows 0..3 of the upper triangle
            if ((lane >> 3) + 4 <= (lane & 7)) Rd[lane + 32] = Rkk[96 + lane];   // rows 4..7
          }
          tq_bar_arrive(1 + (k & 1), 128);
        }
      } else {
        for (int k = P.npriv; k < ncb; ++k) {
          tq_bar_sync(1 + (k & 1), 128);
          const unsigned gm = P.gmask[k];
          const unsigned tm = P.trmask[k];
          if (gm && tm) {
            tq_panel_T<8>(P, tile, k, gm, scb + 8 * (k & 1), TSw, offs.oy, lane);
            for (int j = k + 1; j < ncb; ++j) {  // block k + 1 first: the next panel waits for it
              if ((j - P.npriv) % 3 != wi || !((tm >> j) & 1u)) continue;
              tq_trailing<8>(P, tile, Rg + (size_t)tq_rblk(ncb, k, j) * 64, TSw, k, j, gm, offs, lane);
              if (j == k + 1) tq_bar_arrive(3 + (k & 1), 64);
            }
          }
          if (k + 1 < ncb && (k + 1 - P.npriv) % 3 == wi && !(gm && tm && ((tm >> (k + 1)) & 1u)))
            tq_bar_arrive(3 + (k & 1), 64);  // nothing to apply to block k + 1, but the hand-shake is unconditional
        }
      }
    }
  }
  __syncthreads();
  if (!(phase & 2)) return;
  // ---- write the CTA's factor: packed upper triangle (p1 state columns) + squared column norms
  const int p1 = P.p1;
  const int ntri = p1 * (p1 + 1) / 2;
  double* out = parts + (size_t)blockIdx.x * part_doubles(p1);
  for (int idx = threadIdx.x; idx < p1 * p1; idx += blockDim.x) {
    const int i = idx / p1, c = idx % p1;
    if (c >= i)
      out[tri_off(i, p1) + c - i] = Rg[(size_t)tq_rblk(ncb, P.sc_blk[i], P.sc_blk[c]) * 64 + P.sc_c[i] * 8 + P.sc_c[c]];
  }
  for (int c = threadIdx.x; c < p1; c += blockDim.x) {
    double sacc = 0.0;
    for (int i = 0; i <= c; ++i) {
      const double r = Rg[(size_t)tq_rblk(ncb, P.sc_blk[i], P.sc_blk[c]) * 64 + P.sc_c[i] * 8 + P.sc_c[c]];
      sacc = fma(r, r, sacc);
    }
    out[ntri + c] = sacc;
  }
}

}  // namespace fg
