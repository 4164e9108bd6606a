// fg_groupqr.cuh - kernel (1)+(2), fourth generation: row-group streams.
//
// What the generation-3 profile showed (profiles/ncu_tq_r01_v3.md): 2 165 warp instructions per
// sample at an issue rate of 24 %, the largest stall reason by far being the CTA barriers between
// the one panel warp and its three workers - the panel chain of a tile that mixes all row groups is
// serial, and every trailing update pays for the fill-in between row groups.
//
// This kernel removes both.  Rows of different row groups are different rows of W, hence
// R = qr([R_0; ...; R_{ms-1}]) with R_g the factor of row group g's own non-zero columns (fg_gq.h).
// A warp works on ONE row group for the whole launch (the number of warps per row group is proportional to
// the group's cost, fg_gq.h) and, per tile of 32 samples:
//   1. runs the forward pass with one sample per lane and writes that row group's entries
//      s_g^T B_b (pinocchio::computeJointTorqueRegressor, reference call sites regressor.py:49,93)
//      into its private 32 x w_g shared-memory tile,
//   2. folds the tile into its private block-triangular R_g (L2-resident scratch) by a blocked
//      Householder QR: 8-column panels with one row per lane, compact-WY T, DMMA trailing updates
//      whose reflector fragments stay in registers for the whole panel.
// No warp ever waits for another one: a CTA is a single warp, there is no barrier in the kernel.
// The per-warp factors are merged by gq_merge_kernel (pairwise tree per row group driven by arrival
// counters, then the ms roots are folded into the full factor in state column order).
#pragma once
#include "fg_tileqr.cuh"
#include "fg_gq.h"

namespace fg {

struct GqOffs {
  int oc0[4], oc1[4];  // (rows 8nn + 2rq, +1; column cq): C fragments of a tile block, also A/B of the row contraction
  int ob0[4], ob1[4];  // (row 8nn + cq; columns 2rq, 2rq + 1): V^T as B fragment of the update
};
__device__ __forceinline__ void gq_make_offs(GqOffs& o, int lane) {
  const int cq = lane >> 2, rq = lane & 3;
#pragma unroll
  for (int nn = 0; nn < 4; ++nn) {
    o.oc0[nn] = tq_idx(cq, 8 * nn + 2 * rq);
    o.oc1[nn] = tq_idx(cq, 8 * nn + 2 * rq + 1);
    o.ob0[nn] = tq_idx(2 * rq, 8 * nn + cq);
    o.ob1[nn] = tq_idx(2 * rq + 1, 8 * nn + cq);
  }
}

template <bool CG>
__device__ __forceinline__ double gq_ld(const double* p) {
  if (CG) return __ldcg(p);
  return *p;
}

// Compact-WY factor from the register fragments of the normalised reflectors:
// T^-1 = diag(1/tau) + striu(V^T V); V^T V is a DMMA whose A and B fragments are the same register
// (the row contraction runs over the rows in fragment order).  sc[0..7]: taus; TS: T row-major.
template <int RPL>
__device__ __forceinline__ void gq_panel_T(const double (&vc0)[RPL][4], const double (&vc1)[RPL][4], const double* sc,
                                           double* TS, int lane) {
  const int cq = lane >> 2, rq = lane & 3;
  double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
  for (int sl = 0; sl < RPL; ++sl)
#pragma unroll
    for (int nn = 0; nn < 4; ++nn) {
      tq_dmma(a0, a1, vc0[sl][nn], vc0[sl][nn]);
      tq_dmma(b0, b1, vc1[sl][nn], vc1[sl][nn]);
    }
  a0 += b0;
  a1 += b1;
  if (2 * rq < cq) TS[cq * 8 + 2 * rq] = a0;
  if (2 * rq + 1 < cq) TS[cq * 8 + 2 * rq + 1] = a1;
  if (lane < 8) TS[lane * 8 + lane] = sc[lane];
  __syncwarp();
  // Column `lane` of T by back substitution (lanes 0..7).  The results go straight into the UPPER part of TS
  // (T[m][lane], m <= lane) while the strictly lower part still holds the dot products the other lanes read;
  // consumers mask the lower part to zero when they load T.
  if (lane < 8) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0;
#pragma unroll
    for (int m = 7; m >= 0; --m) {
      const double taum = TS[m * 8 + m];
      const double tm = (m == lane) ? taum : (m < lane ? -taum * acc[m] : 0.0);
      if (m < lane) TS[m * 8 + lane] = tm;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < m) acc[i] = fma(TS[m * 8 + i], tm, acc[i]);
    }
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// sqrt(x) and 1/sqrt(x) of a normal positive double: MUFU seed (2^-22) + two coupled Newton steps
__device__ __forceinline__ void gq_sqrt_rsqrt(double x, double& sq, double& rs) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  sq = fma(g, r, g);
  rs = 2.0 * fma(h, r, h);
}
__device__ __forceinline__ double gq_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// Householder panel, one warp, one row per lane (x[c] = entry (row = lane, column c)) under the 8 x 8
// diagonal block Rin (row-major, shared memory).  The 32-lane dot products of column j with the columns to
// its right are reduced through shared memory instead of a 5-level shuffle butterfly per value: every lane
// stores its partial products into Pb (8 values x 32 lanes, the panel's own tile block, free while the
// panel sits in registers), lane (value = lane >> 2, part = lane & 3) adds 8 of them and a 2-level
// butterfly over `part` finishes the sum; the 8 results go back through 8 doubles of scratch (sc[8..15]).
// Rotating value c's row by 4c doubles and the read order by part >> 1 keeps both sides bank-conflict free.
// Rin and Rout may be the same 64 doubles (row j is stored only after every lane has read it).
// On exit x holds the normalised reflectors v = x / (alpha - beta) (LAPACK dlarfg convention), Rout the
// updated block (upper triangle), sc[0..7] the taus (0 for a skipped / padding column).
// RPL rows per lane (x[s][c] = entry (row = lane of slab s, column c)): the merge kernel folds 64-row slabs.
// (Measured and rejected: skipping the column steps of a padded last block behind uniform branches, and taking
// sigma by shuffle instead of through the scratch - both break the compiler's interleaving of the unrolled
// column steps and cost more than they save: 4.92 ms and 4.20 ms against 4.15 ms.)
template <int RPL>
__device__ __forceinline__ void gq_panel_factor(double (&x)[RPL][8], const double* Rin, double* Rout, double* sc,
                                                double* Pb, int lane) {
  const int ci = lane >> 2, part = lane & 3;
  int rd[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) rd[k] = ci * 32 + ((8 * part + 2 * ((k + (part >> 1)) & 3) + 4 * ci) & 31);
  double* dbuf = sc + 8;
  double rw[8];  // updated row j - 1 of the diagonal block (lane-uniform), stored once every lane has read it
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int c = j; c < 8; ++c) {
      double pr = x[0][j] * x[0][c];
#pragma unroll
      for (int sl = 1; sl < RPL; ++sl) pr = fma(x[sl][j], x[sl][c], pr);
      Pb[c * 32 + ((lane + 4 * c) & 31)] = pr;
    }
    __syncwarp();
    if (j > 0 && lane == 0) {
#pragma unroll
      for (int c = j - 1; c < 8; ++c) Rout[(j - 1) * 8 + c] = rw[c];
    }
    double s;
    {
      const double2 p0 = *reinterpret_cast<const double2*>(Pb + rd[0]);
      const double2 p1 = *reinterpret_cast<const double2*>(Pb + rd[1]);
      const double2 p2 = *reinterpret_cast<const double2*>(Pb + rd[2]);
      const double2 p3 = *reinterpret_cast<const double2*>(Pb + rd[3]);
      s = ((p0.x + p0.y) + (p1.x + p1.y)) + ((p2.x + p2.y) + (p3.x + p3.y));
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (part == 0) dbuf[ci] = s;
    __syncwarp();
    double d[8];
#pragma unroll
    for (int c = (j & ~1); c < 8; c += 2) {
      const double2 t = *reinterpret_cast<const double2*>(dbuf + c);
      d[c] = t.x;
      d[c + 1] = t.y;
    }
    const double sigma = d[j];
    const double alpha = Rin[j * 8 + j];
    double beta = alpha, tauh = 0.0, inv = 0.0;
    const double mu = fma(alpha, alpha, sigma);
    if (sigma != 0.0 && mu > 1e-280) {  // uniform: every lane read the same d[j]
      double nrm, rs;
      gq_sqrt_rsqrt(mu, nrm, rs);
      beta = alpha >= 0.0 ? -nrm : nrm;
      const double den = alpha - beta;
      inv = gq_rcp(den);
      tauh = fabs(den) * rs;  // (beta - alpha) / beta
    }
    if (lane == 0) sc[j] = tauh;
    rw[j] = beta;
#pragma unroll
    for (int c = j + 1; c < 8; ++c) {
      const double rho = Rin[j * 8 + c];
      const double w = tauh * fma(d[c], inv, rho);  // tau * v^T [rho; b]  (0 when the reflector is skipped)
      const double coef = w * inv;
#pragma unroll
      for (int sl = 0; sl < RPL; ++sl) x[sl][c] = fma(-coef, x[sl][j], x[sl][c]);
      rw[c] = rho - w;
    }
#pragma unroll
    for (int sl = 0; sl < RPL; ++sl) x[sl][j] *= inv;  // H = I (dlarfg with xnorm == 0) leaves v = 0
  }
  __syncwarp();
  if (lane == 0) Rout[63] = rw[7];
  __syncwarp();
}

// Fold a tile of RPL slabs of 32 rows (slab s, block j at tile + (s nb + j) GQ_BLK; nb blocks of 8 columns,
// shared memory, blocks < k0 all zero) into the block-upper-triangular factor Rg (nb (nb + 1) / 2 blocks of
// 8 x 8, row-major, global memory).
// hook(k) is called by every lane once block k of the tile is dead (its panel is factored and applied to all the
// blocks to its right): generation 5 refills that block with the next tile's data while the fold goes on.
struct GqNoHook {
  __device__ __forceinline__ void operator()(int) const {}
};
template <bool CG, int RPL, class Hook>
__device__ __forceinline__ void gq_fold_h(double* tile, int nb, double* Rg, int k0, double* scr, const GqOffs& o,
                                          int lane, Hook& hook) {
  const int cq = lane >> 2, rq = lane & 3;
  double* TS = scr;   // T of the panel: written after the diagonal block has gone back to global memory
  double* Rin = scr;  // diagonal block, updated in place
  double* Rout = scr;
  double* sc = scr + 64;
  double d0, d1;
  {
    const double* Rd = Rg + (size_t)tq_rblk(nb, k0, k0) * 64;
    d0 = gq_ld<CG>(Rd + lane);
    d1 = gq_ld<CG>(Rd + lane + 32);
  }
  for (int k = k0; k < nb; ++k) {
    double* Rd = Rg + (size_t)tq_rblk(nb, k, k) * 64;
    double* Xk = tile + k * GQ_BLK;
    Rin[lane] = d0;
    Rin[lane + 32] = d1;
    double x[RPL][8];
#pragma unroll
    for (int sl = 0; sl < RPL; ++sl)
#pragma unroll
      for (int c = 0; c < 8; ++c) x[sl][c] = Xk[sl * nb * GQ_BLK + tq_idx(c, lane)];
    // first off-diagonal block of this block row: requested before the panel so that its L2 round trip is hidden
    double* Rb = Rd + 64;
    double r0 = 0.0, r1 = 0.0;
    if (k + 1 < nb) { r0 = gq_ld<CG>(Rb + (2 * rq) * 8 + cq); r1 = gq_ld<CG>(Rb + (2 * rq + 1) * 8 + cq); }
    __syncwarp();
    gq_panel_factor<RPL>(x, Rin, Rout, sc, Xk, lane);
    if ((lane >> 3) <= (lane & 7)) Rd[lane] = Rout[lane];            // rows 0..3 of the upper triangle
    if ((lane >> 3) + 4 <= (lane & 7)) Rd[lane + 32] = Rout[lane + 32];  // rows 4..7
    if (k + 1 == nb) {
      __syncwarp();
      hook(k);
      break;
    }
    {  // the next diagonal block is only touched by the next panel: request it now
      const double* Rn = Rg + (size_t)tq_rblk(nb, k + 1, k + 1) * 64;
      d0 = gq_ld<CG>(Rn + lane);
      d1 = gq_ld<CG>(Rn + lane + 32);
    }
#pragma unroll
    for (int sl = 0; sl < RPL; ++sl)
#pragma unroll
      for (int c = 0; c < 8; ++c) Xk[sl * nb * GQ_BLK + tq_idx(c, lane)] = x[sl][c];
    __syncwarp();
    // reflector fragments of this panel, kept in registers for every trailing block
    double vc0[RPL][4], vc1[RPL][4], vb0[RPL][4], vb1[RPL][4];
#pragma unroll
    for (int sl = 0; sl < RPL; ++sl)
#pragma unroll
      for (int nn = 0; nn < 4; ++nn) {
        vc0[sl][nn] = Xk[sl * nb * GQ_BLK + o.oc0[nn]];
        vc1[sl][nn] = Xk[sl * nb * GQ_BLK + o.oc1[nn]];
        vb0[sl][nn] = Xk[sl * nb * GQ_BLK + o.ob0[nn]];
        vb1[sl][nn] = Xk[sl * nb * GQ_BLK + o.ob1[nn]];
      }
    gq_panel_T<RPL>(vc0, vc1, sc, TS, lane);
    const double t0 = 2 * rq <= cq ? TS[(2 * rq) * 8 + cq] : 0.0, t1 = 2 * rq + 1 <= cq ? TS[(2 * rq + 1) * 8 + cq] : 0.0;
    for (int j = k + 1; j < nb; ++j, Rb += 64) {
      double nr0 = 0.0, nr1 = 0.0;
      if (j + 1 < nb) {  // R blocks of one block row are consecutive: request the next one a block ahead
        nr0 = gq_ld<CG>(Rb + 64 + (2 * rq) * 8 + cq);
        nr1 = gq_ld<CG>(Rb + 64 + (2 * rq + 1) * 8 + cq);
      }
      double* Bj = tile + j * GQ_BLK;
      double c0[RPL][4], c1[RPL][4];
#pragma unroll
      for (int sl = 0; sl < RPL; ++sl)
#pragma unroll
        for (int nn = 0; nn < 4; ++nn) {
          c0[sl][nn] = Bj[sl * nb * GQ_BLK + o.oc0[nn]];
          c1[sl][nn] = Bj[sl * nb * GQ_BLK + o.oc1[nn]];
        }
      // Y^T = R_kj^T + B_j^T V   (rows contracted in fragment order: the C fragments are the A operands)
      double y0 = 0.0, y1 = 0.0, z0 = 0.0, z1 = 0.0;
#pragma unroll
      for (int sl = 0; sl < RPL; ++sl)
#pragma unroll
        for (int nn = 0; nn < 4; ++nn) {
          tq_dmma(y0, y1, c0[sl][nn], vc0[sl][nn]);
          tq_dmma(z0, z1, c1[sl][nn], vc1[sl][nn]);
        }
      y0 += z0 + r0;
      y1 += z1 + r1;
      // W^T = Y^T T
      double w0 = 0.0, w1 = 0.0;
      tq_dmma(w0, w1, y0, t0);
      tq_dmma(w0, w1, y1, t1);
      Rb[(2 * rq) * 8 + cq] = r0 - w0;
      Rb[(2 * rq + 1) * 8 + cq] = r1 - w1;
      const double nw0 = -w0, nw1 = -w1;
      // B_j^T -= W^T V^T
#pragma unroll
      for (int sl = 0; sl < RPL; ++sl)
#pragma unroll
        for (int nn = 0; nn < 4; ++nn) {
          tq_dmma(c0[sl][nn], c1[sl][nn], nw0, vb0[sl][nn]);
          tq_dmma(c0[sl][nn], c1[sl][nn], nw1, vb1[sl][nn]);
          Bj[sl * nb * GQ_BLK + o.oc0[nn]] = c0[sl][nn];
          Bj[sl * nb * GQ_BLK + o.oc1[nn]] = c1[sl][nn];
        }
      r0 = nr0;
      r1 = nr1;
    }
    __syncwarp();
    hook(k);
  }
  __syncwarp();
}

template <bool CG, int RPL>
__device__ __forceinline__ void gq_fold(double* tile, int nb, double* Rg, int k0, double* scr, const GqOffs& o,
                                        int lane) {
  GqNoHook nh;
  gq_fold_h<CG, RPL, GqNoHook>(tile, nb, Rg, k0, scr, o, lane, nh);
}

// ---------------------------------------------------------------------------------------------
// regressor entries of (row group g, body): value[k] = s^T B[:, k] with s = [sl; sa] the motion axis of the
// row group expressed in the body's frame (see tq_emit).
// ed = (first local slot << 10) | mask of the generated pinocchio parameters: the generated entries of one
// body occupy consecutive local columns in parameter order.
__device__ __forceinline__ void gq_emit(double* tile, int ed, const double* sl, const double* sa, const double* om,
                                        const double* al, const double* ap, bool valid, int lane) {
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
  int slot = ed >> 10;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    if ((ed >> k) & 1) {  // uniform
      if (valid) tile[(slot >> 3) * GQ_BLK + tq_idx(slot & 7, lane)] = val[k];
      ++slot;
    }
  }
}

// Forward pass of one sample per lane, emitting the entries of row group g.  S: the generator tables in
// shared memory.  The (q, dq, ddq) entries of a joint are requested two joints ahead.
static __device__ __forceinline__ void gq_generate(const GqParams& P, const GqGen& S, int g, double* tile,
                                                   const double* __restrict__ q, const double* __restrict__ qd,
                                                   const double* __restrict__ qdd, const double* __restrict__ tau,
                                                   int64_t ld_tau, int64_t s0, int64_t N, int lane) {
  const int64_t smp = s0 + lane;
  const bool valid = smp < N;
  const int64_t sc = valid ? smp : N - 1;
  const double* qs = q + sc * P.nq;
  const double* vs = qd + sc * P.nv;
  const double* as = qdd + sc * P.nv;
  const GqGroup& G = S.grp[g];
  const int nproc = P.nproc;
  double sx[6] = {0, 0, 0, 0, 0, 0};  // motion axis of the row group in the frame of the joint being walked
  double v[6] = {0, 0, 0, 0, 0, 0}, a[6] = {0, 0, 0, 0, 0, 0};
  // one-joint-ahead input queue (a deeper queue rotated by register moves adds no cover: the move is the first
  // touch of the load, measured)
  double nq0 = __ldg(qs + S.pj[0].idx_q), nv0 = __ldg(vs + S.pj[0].idx_v), na0 = __ldg(as + S.pj[0].idx_v);
  // rhs entry and the row group's own (dq, ddq): requested while the LAST joint is walked, consumed after the walk
  double taug = 0.0, vg = 0.0, ag = 0.0;
  const int vidx_g = S.pj[G.tjoint].idx_v;
  for (int t = 0; t < nproc; ++t) {
    const DevJoint& jt = S.pj[t];
    const double cq0 = nq0, cv0 = nv0, ca0 = na0;
    if (t + 1 < nproc) {
      const DevJoint& jn = S.pj[t + 1];
      nq0 = __ldg(qs + jn.idx_q); nv0 = __ldg(vs + jn.idx_v); na0 = __ldg(as + jn.idx_v);
    }
    if (jt.parent == 0) {  // restart from the universe (the host only admits chains hanging off it)
      for (int k = 0; k < 6; ++k) v[k] = 0.0;
      a[0] = -P.gravity[0]; a[1] = -P.gravity[1]; a[2] = -P.gravity[2];
      a[3] = a[4] = a[5] = 0.0;
    }
    double Rj[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pj[3] = {0, 0, 0};
    double vj[6] = {0, 0, 0, 0, 0, 0}, aj[6] = {0, 0, 0, 0, 0, 0};
    const int iq = jt.idx_q, iv = jt.idx_v;
    const double ax[3] = {jt.axis[0], jt.axis[1], jt.axis[2]};
    if (jt.jtype == 1) {
      double s, c;
      sincos(cq0, &s, &c);
      axis_rot(ax, c, s, Rj);
      for (int k = 0; k < 3; ++k) { vj[3 + k] = ax[k] * cv0; aj[3 + k] = ax[k] * ca0; }
    } else if (jt.jtype == 2) {
      axis_rot(ax, cq0, __ldg(qs + iq + 1), Rj);
      for (int k = 0; k < 3; ++k) { vj[3 + k] = ax[k] * cv0; aj[3 + k] = ax[k] * ca0; }
    } else if (jt.jtype == 3) {
      for (int k = 0; k < 3; ++k) { pj[k] = ax[k] * cq0; vj[k] = ax[k] * cv0; aj[k] = ax[k] * ca0; }
    } else {  // free-flyer
      quat_rot(__ldg(qs + iq + 3), __ldg(qs + iq + 4), __ldg(qs + iq + 5), __ldg(qs + iq + 6), Rj);
      for (int k = 0; k < 3; ++k) pj[k] = __ldg(qs + iq + k);
      for (int k = 0; k < 6; ++k) { vj[k] = __ldg(vs + iv + k); aj[k] = __ldg(as + iv + k); }
    }
    // liMi = placement * Mj
    double R[9], p[3];
    for (int r = 0; r < 3; ++r) {
      const double m0 = jt.plcR[3 * r], m1 = jt.plcR[3 * r + 1], m2 = jt.plcR[3 * r + 2];
      for (int c = 0; c < 3; ++c) R[3 * r + c] = m0 * Rj[c] + m1 * Rj[3 + c] + m2 * Rj[6 + c];
      p[r] = m0 * pj[0] + m1 * pj[1] + m2 * pj[2] + jt.plcp[r];
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
    for (int k = 0; k < 6; ++k) { v[k] = vn[k]; a[k] = an[k]; }
    // the row group's motion axis: its own joint's axis at that joint, transported down the chain afterwards
    // (liMi.actInv: the same transform the velocity takes) - the product of these transports is oM_body^-1 oM_joint
    if (t == nproc - 1) {
      if (valid && tau != nullptr) taug = __ldg(tau + (int64_t)g * ld_tau + smp);
      vg = __ldg(vs + vidx_g);
      ag = __ldg(as + vidx_g);
    }
    if (t == G.tjoint) {
      for (int k = 0; k < 6; ++k) sx[k] = G.S[k];
    } else if (t > G.tjoint) {
      cross3(p, sx + 3, tt);
      for (int k = 0; k < 3; ++k) tt[k] = sx[k] - tt[k];
      double sn[6];
      for (int k = 0; k < 3; ++k) {
        sn[k] = R[k] * tt[0] + R[3 + k] * tt[1] + R[6 + k] * tt[2];
        sn[3 + k] = R[k] * sx[3] + R[3 + k] * sx[4] + R[6 + k] * sx[5];
      }
      for (int k = 0; k < 6; ++k) sx[k] = sn[k];
    }
    const int ed = S.emit[g][t];
    if (ed) {
      double ap[3];
      cross3(v + 3, v, ap);  // classical acceleration a' = a_lin + w x v_lin
      for (int k = 0; k < 3; ++k) ap[k] += a[k];
      gq_emit(tile, ed, sx, sx + 3, v + 3, a + 3, ap, valid, lane);
    }
  }
  // columns read straight off the inputs
  for (int t = 0; t < G.nsimple; ++t) {
    const GqSimple s1 = S.simple[g][t];
    const int ix = (s1.slot >> 3) * GQ_BLK + tq_idx(s1.slot & 7, lane);
    double val = 0.0;
    if (s1.kind == COL_RHS) {
      val = taug;
    } else if (s1.kind == COL_EXTRA) {
      const bool own = s1.vidx == vidx_g;  // uniform
      const double vv = own ? vg : __ldg(vs + s1.vidx), aa = own ? ag : __ldg(as + s1.vidx);
      val = s1.src == SRC_QDD ? aa : (s1.src == SRC_QD ? vv : (s1.src == SRC_SIGN_QD ? sgn(vv) : 1.0));
    } else {
      // add_coupling_TX40 (regressor.py:216-225): row block 4 <- [a5, v5, sign(v4+v5)], 5 <- [a4, v4, sign(v4+v5)]
      const double v4 = __ldg(vs + 4), v5 = __ldg(vs + 5), a4 = __ldg(as + 4), a5 = __ldg(as + 5);
      const double sg = sgn(v4 + v5);
      val = G.rowblock == 4 ? (s1.src == 0 ? a5 : (s1.src == 1 ? v5 : sg)) : (s1.src == 0 ? a4 : (s1.src == 1 ? v4 : sg));
    }
    tile[ix] = valid ? val : 0.0;
  }
}

#ifndef FG_GQ_DEVICE_ONLY  // fg_rq.cu reuses the device functions above, not the generation-4 kernels
// ---------------------------------------------------------------------------------------------
// One CTA per SM, blockDim.x / 32 independent warps; the only CTA-wide synchronisation is the one after the
// generator tables have been copied to shared memory.  Warp gw = blockIdx.x * nw + warp works on row group
// tab_g[gw] as its instance tab_i[gw]: it folds the tiles inst, inst + ninst, ... into its own factor.  The
// shared-memory tile of a warp has exactly the width of its row group.
__global__ void __launch_bounds__(GQ_MAXNW * 32, 1)
gq_accumulate_kernel(const __grid_constant__ GqParams P, const unsigned char* __restrict__ tab_g,
                     const unsigned short* __restrict__ tab_i, const double* __restrict__ q,
                     const double* __restrict__ qd, const double* __restrict__ qdd, const double* __restrict__ tau,
                     int64_t ld_tau, int64_t N, double* __restrict__ r_all, int first) {
  extern __shared__ __align__(16) double gq_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  GqGen* Sg = reinterpret_cast<GqGen*>(gq_smem);
  {
    const int* src = reinterpret_cast<const int*>(&P.gen);
    int* dst = reinterpret_cast<int*>(Sg);
    for (int i = threadIdx.x; i < (int)(sizeof(GqGen) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const GqGen& S = *Sg;
  const int gw = blockIdx.x * nw + warp;
  const int g = tab_g[gw], inst = tab_i[gw];
  int boff = 0;
  for (int w = 0; w < warp; ++w) boff += S.grp[tab_g[blockIdx.x * nw + w]].nb;
  double* tile = gq_smem + GQ_GEN_DOUBLES + (size_t)boff * GQ_BLK + (size_t)warp * GQ_SCRATCH;
  const GqGroup& G = S.grp[g];
  const int nb = G.nb, ninst = G.ninst;
  double* scr = tile + nb * GQ_BLK;
  double* Rg = r_all + (size_t)G.r_off + (size_t)inst * G.r_size;
  if (first) {
    for (int i = lane; i < G.r_size; i += 32) Rg[i] = 0.0;
    __syncwarp();
  }
  GqOffs offs;
  gq_make_offs(offs, lane);
  const int64_t ntiles = (N + GQ_S - 1) / GQ_S;
  for (int64_t tl = inst; tl < ntiles; tl += ninst) {
    if (tl + ninst < ntiles) {  // pull the next tile's inputs into L2 while this one is folded
      const int64_t sn = (tl + ninst) * GQ_S + lane;
      if (sn < N) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q + sn * P.nq));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(qd + sn * P.nv));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(qdd + sn * P.nv));
      }
    }
    {
      double2* t2 = reinterpret_cast<double2*>(tile);
      for (int i = lane; i < nb * (GQ_BLK / 2); i += 32) t2[i] = make_double2(0.0, 0.0);
    }
    __syncwarp();
    gq_generate(P, S, g, tile, q, qd, qdd, tau, ld_tau, tl * GQ_S, N, lane);
    __syncwarp();
    gq_fold<false, 1>(tile, nb, Rg, 0, scr, offs, lane);
  }
}

// ---------------------------------------------------------------------------------------------
// Merge of the per-warp factors.  One warp per (row group g, pair of instances).  A warp folds instance
// i + stride into instance i, then climbs: at every internal node the first of the two children to arrive
// leaves, the second one does the node's merge (always "right child into left child", so the result does not
// depend on who came last).
struct GqMergeArgs {
  int leaves[GQ_MAXG];    // instances of row group g that take part
  int blk0[GQ_MAXG + 1];  // first block of row group g (one block = one warp = one pair of instances)
};

__device__ __forceinline__ void gq_merge_into(const GqParams& P, double* tile, int nb, double* Rdst, const double* Rsrc,
                                              double* scr, const GqOffs& o, int lane) {
  // 64-row slabs (two rows per lane): 8 block rows of the source per fold, i.e. ONE fold for nb <= 8
  for (int t0 = 0; t0 < nb; t0 += 8) {
    {
      double2* t2 = reinterpret_cast<double2*>(tile);
      for (int i = lane; i < 2 * nb * (GQ_BLK / 2); i += 32) t2[i] = make_double2(0.0, 0.0);
    }
    __syncwarp();
    for (int m = 0; m < 8 && t0 + m < nb; ++m) {
      const int rb = t0 + m;
      double* slab = tile + (size_t)(m >> 2) * nb * GQ_BLK;
      const int mm = m & 3;
      // the blocks of one block row are consecutive: request all of them before the first one is stored
      const double* B = Rsrc + (size_t)tq_rblk(nb, rb, rb) * 64;
      for (int j0 = 0; rb + j0 < nb; j0 += 8) {
        double v0[8], v1[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          if (rb + j0 + jj < nb) { v0[jj] = __ldcg(B + (j0 + jj) * 64 + lane); v1[jj] = __ldcg(B + (j0 + jj) * 64 + 32 + lane); }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          if (rb + j0 + jj < nb) {
            slab[(rb + j0 + jj) * GQ_BLK + tq_idx(lane & 7, 8 * mm + (lane >> 3))] = v0[jj];
            slab[(rb + j0 + jj) * GQ_BLK + tq_idx(lane & 7, 8 * mm + 4 + (lane >> 3))] = v1[jj];
          }
      }
    }
    __syncwarp();
    gq_fold<true, 2>(tile, nb, Rdst, t0, scr, o, lane);
  }
}

__global__ void __launch_bounds__(32)
gq_merge_kernel(const __grid_constant__ GqParams P, const __grid_constant__ GqMergeArgs M, double* __restrict__ r_all,
                unsigned* __restrict__ counters, double* __restrict__ rfull, double* __restrict__ part_out) {
  extern __shared__ __align__(16) double gq_smem[];
  const int lane = threadIdx.x;
  int g = 0;
  while (g + 1 < P.ms && (int)blockIdx.x >= M.blk0[g + 1]) ++g;
  int i = 2 * ((int)blockIdx.x - M.blk0[g]);
  const int leaves = M.leaves[g];
  const int mb = P.maxnb > P.nbf ? P.maxnb : P.nbf;
  double* tile = gq_smem;
  double* scr = gq_smem + 2 * mb * GQ_BLK;
  GqOffs offs;
  gq_make_offs(offs, lane);
  const GqGroup& G = P.gen.grp[g];
  const int nb = G.nb;
  double* Rgrp = r_all + (size_t)G.r_off;
  for (int stride = 1;; stride *= 2) {
    if (i + stride < leaves)
      gq_merge_into(P, tile, nb, Rgrp + (size_t)i * G.r_size, Rgrp + (size_t)(i + stride) * G.r_size, scr, offs, lane);
    const int s2 = 2 * stride;
    if (s2 >= leaves) break;  // i == 0: the root of this row group is complete
    const int ip = i & ~(2 * s2 - 1);
    const int sib = (i == ip) ? ip + s2 : ip;
    if (sib < leaves) {
      __threadfence();
      __syncwarp();
      unsigned old = 0;
      if (lane == 0) old = atomicAdd(&counters[(size_t)g * P.maxinst + ip + s2], 1u);
      old = __shfl_sync(0xffffffffu, old, 0);
      if (old == 0) return;
      __threadfence();
    }
    i = ip;
  }
  // ---- root of row group g done.  The roots are folded into the full factor in the fixed order ms-1, ..., 0
  // (deterministic result) by whichever warp holds the BUSY bit of `fin`: a finishing root publishes its bit and,
  // if nobody is busy, becomes the worker; the worker folds the next root in order as soon as its bit is there
  // and otherwise releases BUSY (compare-and-swap against the value it examined, so a bit that arrives in
  // between makes the release fail and is picked up).  No warp ever spins on another one.  The small row
  // groups finish their trees first, so their folds overlap the trees of the large ones.
  unsigned* fin = counters + (size_t)GQ_MAXG * P.maxinst;
  unsigned* nfolded = fin + 1;
  const unsigned BUSY = 0x80000000u;
  {
    __threadfence();
    __syncwarp();
    unsigned old = 0;
    if (lane == 0) old = atomicOr(fin, (1u << g) | BUSY);
    old = __shfl_sync(0xffffffffu, old, 0);
    if (old & BUSY) return;
  }
  const int nbf = P.nbf;
  for (;;) {
    unsigned cur = 0, done = 0;
    if (lane == 0) { cur = atomicOr(fin, 0u); done = atomicAdd(nfolded, 0u); }
    cur = __shfl_sync(0xffffffffu, cur, 0);
    done = __shfl_sync(0xffffffffu, done, 0);
    __threadfence();
    const int gg = P.ms - 1 - (int)done;
    if (gg < 0) break;  // every root is in: write the result
    if (!((cur >> gg) & 1u)) {
      unsigned seen = 0;
      if (lane == 0) seen = atomicCAS(fin, cur, cur & ~BUSY);
      seen = __shfl_sync(0xffffffffu, seen, 0);
      if (seen == cur) return;
      continue;
    }
    if (done == 0) {
      for (int idx = lane; idx < nbf * (nbf + 1) / 2 * 64; idx += 32) rfull[idx] = 0.0;
      __syncwarp();
    }
    const GqGroup& H = P.gen.grp[gg];
    const double* Rsrc = r_all + H.r_off;
    for (int t0 = 0; t0 < H.nb; t0 += 8) {  // 64-row slabs of the root
      {
        double2* t2 = reinterpret_cast<double2*>(tile);
        for (int idx = lane; idx < 2 * nbf * (GQ_BLK / 2); idx += 32) t2[idx] = make_double2(0.0, 0.0);
      }
      __syncwarp();
      int kmin = nbf - 1;
      for (int m = 0; m < 8 && t0 + m < H.nb; ++m) {
        const int rb = t0 + m;
        double* slab = tile + (size_t)(m >> 2) * nbf * GQ_BLK;
        const int mm = m & 3;
        const double* B = Rsrc + (size_t)tq_rblk(H.nb, rb, rb) * 64;
        for (int j0 = 0; rb + j0 < H.nb; j0 += 8) {
          double v0[8], v1[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            if (rb + j0 + jj < H.nb) { v0[jj] = __ldcg(B + (j0 + jj) * 64 + lane); v1[jj] = __ldcg(B + (j0 + jj) * 64 + 32 + lane); }
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int l = 8 * (rb + j0 + jj) + (lane & 7);
            if (rb + j0 + jj < H.nb && l < H.w) {
              const int scol = P.gcol[gg][l];
              slab[(scol >> 3) * GQ_BLK + tq_idx(scol & 7, 8 * mm + (lane >> 3))] = v0[jj];
              slab[(scol >> 3) * GQ_BLK + tq_idx(scol & 7, 8 * mm + 4 + (lane >> 3))] = v1[jj];
              kmin = min(kmin, scol >> 3);
            }
          }
        }
      }
      kmin = __reduce_min_sync(0xffffffffu, kmin);
      __syncwarp();
      gq_fold<true, 2>(tile, nbf, rfull, kmin, scr, offs, lane);
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicAdd(nfolded, 1u);
    __syncwarp();
  }
  __syncwarp();
  // ---- packed upper triangle + squared column norms
  const int p1 = P.p1;
  const int ntri = p1 * (p1 + 1) / 2;
  for (int idx = lane; idx < p1 * p1; idx += 32) {
    const int r = idx / p1, c = idx % p1;
    if (c >= r) part_out[tri_off(r, p1) + c - r] = __ldcg(rfull + (size_t)tq_rblk(nbf, r >> 3, c >> 3) * 64 + (r & 7) * 8 + (c & 7));
  }
  for (int c = lane; c < p1; c += 32) {
    double sacc = 0.0;
    for (int r = 0; r <= c; ++r) {
      const double x = __ldcg(rfull + (size_t)tq_rblk(nbf, r >> 3, c >> 3) * 64 + (r & 7) * 8 + (c & 7));
      sacc = fma(x, x, sacc);
    }
    part_out[ntri + c] = sacc;
  }
}
#endif  // FG_GQ_DEVICE_ONLY

}  // namespace fg
