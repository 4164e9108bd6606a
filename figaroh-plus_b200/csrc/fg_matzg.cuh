// fg_matzg.cuh - kernel (1) alone for SERIAL CHAINS of single-dof joints (UR10, TX40 without its coupling columns,
// any 6/7-axis arm): W written row-block major, every 32-row x ncols region of a (tile, row block) leaving as whole
// rows through the copy engine.
//
// Why a second materialise kernel: fg_matz.cuh emits (row block, body) patches in the order of the tree walk, so the
// rows of a tile are completed over nb patch passes and warps x SMs x ms x 32 rows of partially written lines must
// survive in L2 (UR10: 129 KB per warp; 0.99 ms = 4.07 TB/s with six warps per SM and evict-last stores, while a pure
// store stream reaches 7.5 TB/s on this GPU: tools/probe_write_bw.py).  For a chain the per-body kinematics are small
// enough to keep: the walk runs ONCE per tile (one sample per lane) and leaves (omega, alpha, a', joint angle, qd,
// qdd) of every body in shared memory (13 doubles per body and lane); then, row block by row block, the lane
// transports the joint's motion axis down the chain, forms the entries s^T B_b of
// pinocchio::computeJointTorqueRegressor (regressor.py:49-70) and the Ia / fv / fs / off columns, writes ITS row of
// the block (ncols doubles, 16-byte stores) into a shared-memory stage and sends it with one cp.async.bulk: no
// cross-lane dependency at all, complete 672-byte rows at DRAM, 22 KB of stage per warp instead of 129 KB in L2.
#pragma once
#include "fg_matz.cuh"
#include "fg_matzg_params.h"

namespace fg {

constexpr int MZG_KIN = 13;   // per body and lane: omega 3, alpha 3, a' 3, (c, s) of the joint angle (or q, -), qd, qdd

__global__ void __launch_bounds__(MATZG_WARPS * 32, 1)
matzg_kernel(const __grid_constant__ MatzgParams P, const MatzTables* __restrict__ tab, const double* __restrict__ q,
             const double* __restrict__ qd, const double* __restrict__ qdd, int64_t N, double* __restrict__ W, int64_t ldw) {
  extern __shared__ __align__(16) double mzg_smem[];
  MatzTables* T = reinterpret_cast<MatzTables*>(mzg_smem);
  {
    const int* src = reinterpret_cast<const int*>(tab);
    int* dst = reinterpret_cast<int*>(T);
    for (int i = threadIdx.x; i < (int)(sizeof(MatzTables) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  constexpr int TAB_DOUBLES = (int)((sizeof(MatzTables) + 15) / 16) * 2;
  double* kin = mzg_smem + TAB_DOUBLES + (size_t)warp * P.smem_per_warp + lane;   // kin[(b * MZG_KIN + k) * 32]
  double* stage = mzg_smem + TAB_DOUBLES + (size_t)warp * P.smem_per_warp + P.nb * MZG_KIN * 32 + (size_t)lane * P.pitch;
  const int nb = P.nb, ncols = P.ncols;
  const int64_t ntiles = (N + 31) / 32;
  // the stage starts as zeros: the columns of the bodies above a row block's joint stay zero from pass to pass
  for (int c = 0; c < P.pitch; ++c) stage[c] = 0.0;
  for (int64_t tl = (int64_t)blockIdx.x * nw + warp; tl < ntiles; tl += (int64_t)gridDim.x * nw) {
    const int64_t s0 = tl * 32, smp = s0 + lane;
    const bool valid = smp < N;
    const int64_t sc = valid ? smp : N - 1;
    const double* qs = q + sc * P.nq;
    const double* vs = qd + sc * P.nv;
    const double* as = qdd + sc * P.nv;
    {  // the next tile of this warp towards L2
      const int64_t sn = (tl + (int64_t)gridDim.x * nw) * 32 + lane;
      if (sn < N) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q + sn * P.nq));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(qd + sn * P.nv));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(qdd + sn * P.nv));
      }
    }
    // ---- forward pass down the chain (same recursion as rq_walk_tile; joints 1 .. nb in walk order)
    {
      double v[6] = {0, 0, 0, 0, 0, 0}, a[6] = {-P.gravity[0], -P.gravity[1], -P.gravity[2], 0, 0, 0};
      double nq0, nq1 = 0.0, nv0, na0;
      {
        const DevJoint& j0 = T->pj[0];
        nq0 = __ldg(qs + j0.idx_q); nv0 = __ldg(vs + j0.idx_v); na0 = __ldg(as + j0.idx_v);
        if (j0.jtype == 2) nq1 = __ldg(qs + j0.idx_q + 1);
      }
      for (int t = 0; t < nb; ++t) {
        const DevJoint& jt = T->pj[t];
        const double cq0 = nq0, cq1 = nq1, cv0 = nv0, ca0 = na0;
        if (t + 1 < nb) {
          const DevJoint& jn = T->pj[t + 1];
          nq0 = __ldg(qs + jn.idx_q); nv0 = __ldg(vs + jn.idx_v); na0 = __ldg(as + jn.idx_v);
          if (jn.jtype == 2) nq1 = __ldg(qs + jn.idx_q + 1);
        }
        const double ax[3] = {jt.axis[0], jt.axis[1], jt.axis[2]};
        double Rj[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pj[3] = {0, 0, 0}, vj[6] = {0, 0, 0, 0, 0, 0}, aj[6] = {0, 0, 0, 0, 0, 0};
        double k0 = cq0, k1 = 0.0;   // what phase 2 rebuilds the joint transform from
        if (jt.jtype == 1) {
          double s, c;
          sincos(cq0, &s, &c);
          k0 = c; k1 = s;
          axis_rot(ax, c, s, Rj);
          for (int k = 0; k < 3; ++k) { vj[3 + k] = ax[k] * cv0; aj[3 + k] = ax[k] * ca0; }
        } else if (jt.jtype == 2) {
          k0 = cq0; k1 = cq1;
          axis_rot(ax, cq0, cq1, Rj);
          for (int k = 0; k < 3; ++k) { vj[3 + k] = ax[k] * cv0; aj[3 + k] = ax[k] * ca0; }
        } else {
          for (int k = 0; k < 3; ++k) { pj[k] = ax[k] * cq0; vj[k] = ax[k] * cv0; aj[k] = ax[k] * ca0; }
        }
        double R[9], p[3];
        for (int r = 0; r < 3; ++r) {
          const double m0 = jt.plcR[3 * r], m1 = jt.plcR[3 * r + 1], m2 = jt.plcR[3 * r + 2];
          for (int c = 0; c < 3; ++c) R[3 * r + c] = m0 * Rj[c] + m1 * Rj[3 + c] + m2 * Rj[6 + c];
          p[r] = m0 * pj[0] + m1 * pj[1] + m2 * pj[2] + jt.plcp[r];
        }
        double vn[6], an[6], tl3[3], ta3[3], l1[3], l2[3], l3[3];
        rq_actinv(R, p, v, v + 3, tl3, ta3);
        for (int k = 0; k < 3; ++k) { vn[k] = vj[k] + tl3[k]; vn[3 + k] = vj[3 + k] + ta3[k]; }
        cross3(vn + 3, vj, l1);
        cross3(vn, vj + 3, l2);
        cross3(vn + 3, vj + 3, l3);
        rq_actinv(R, p, a, a + 3, tl3, ta3);
        for (int k = 0; k < 3; ++k) { an[k] = l1[k] + l2[k] + aj[k] + tl3[k]; an[3 + k] = l3[k] + aj[3 + k] + ta3[k]; }
        for (int k = 0; k < 6; ++k) { v[k] = vn[k]; a[k] = an[k]; }
        double ap[3];
        cross3(v + 3, v, ap);
        double* K = kin + (size_t)t * MZG_KIN * 32;
        for (int k = 0; k < 3; ++k) { K[k * 32] = v[3 + k]; K[(3 + k) * 32] = a[3 + k]; K[(6 + k) * 32] = ap[k] + a[k]; }
        K[9 * 32] = k0; K[10 * 32] = k1; K[11 * 32] = cv0; K[12 * 32] = ca0;
      }
    }
    // ---- row block by row block: this lane's row of the block, then out through the copy engine
    for (int g = 0; g < P.ms; ++g) {
      const int tg = T->rowjoint[g] - 1;   // walk index of the row block's joint (chain: joint j at walk index j - 1)
      double sl[3] = {T->S[g][0], T->S[g][1], T->S[g][2]}, sa[3] = {T->S[g][3], T->S[g][4], T->S[g][5]};
      for (int t = tg; t < nb; ++t) {
        const double* K = kin + (size_t)t * MZG_KIN * 32;
        if (t > tg) {  // the axis into the frame of body t: joint transform rebuilt from the stored angle
          const DevJoint& jt = T->pj[t];
          const double ax[3] = {jt.axis[0], jt.axis[1], jt.axis[2]};
          double Rj[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pj[3] = {0, 0, 0};
          const double k0 = K[9 * 32], k1 = K[10 * 32];
          if (jt.jtype == 3) { for (int k = 0; k < 3; ++k) pj[k] = ax[k] * k0; }
          else axis_rot(ax, k0, k1, Rj);
          double R[9], p[3];
          for (int r = 0; r < 3; ++r) {
            const double m0 = jt.plcR[3 * r], m1 = jt.plcR[3 * r + 1], m2 = jt.plcR[3 * r + 2];
            for (int c = 0; c < 3; ++c) R[3 * r + c] = m0 * Rj[c] + m1 * Rj[3 + c] + m2 * Rj[6 + c];
            p[r] = m0 * pj[0] + m1 * pj[1] + m2 * pj[2] + jt.plcp[r];
          }
          double nl[3], na[3];
          rq_actinv(R, p, sl, sa, nl, na);
          for (int k = 0; k < 3; ++k) { sl[k] = nl[k]; sa[k] = na[k]; }
        }
        const double om[3] = {K[0], K[32], K[64]}, al[3] = {K[96], K[128], K[160]}, ap[3] = {K[192], K[224], K[256]};
        // pinocchio order m, mx, my, mz, Ixx, Ixy, Iyy, Ixz, Iyz, Izz -> FIGAROH Ixx Ixy Ixz Iyy Iyz Izz mx my mz m
        double val[10];
        val[9] = sl[0] * ap[0] + sl[1] * ap[1] + sl[2] * ap[2];
        const double ww = om[0] * om[0] + om[1] * om[1] + om[2] * om[2];
        const double so = sl[0] * om[0] + sl[1] * om[1] + sl[2] * om[2];
        double x1[3], x2[3], h[3];
        cross3(sl, al, x1);
        cross3(ap, sa, x2);
        for (int j = 0; j < 3; ++j) val[6 + j] = x1[j] + so * om[j] - ww * sl[j] + x2[j];
        cross3(sa, om, h);
        val[0] = sa[0] * al[0] + h[0] * om[0];
        val[1] = sa[0] * al[1] + sa[1] * al[0] + h[0] * om[1] + h[1] * om[0];
        val[3] = sa[1] * al[1] + h[1] * om[1];
        val[2] = sa[0] * al[2] + sa[2] * al[0] + h[0] * om[2] + h[2] * om[0];
        val[4] = sa[1] * al[2] + sa[2] * al[1] + h[1] * om[2] + h[2] * om[1];
        val[5] = sa[2] * al[2] + h[2] * om[2];
        double e10 = 0.0, e11 = 0.0, e12 = 0.0, e13 = 0.0;
        if (T->extra_grp[t + 1] == g) {  // Ia, fv, fs, off of the row block's own joint (regressor.py:55-70)
          const double vv = K[11 * 32], aa = K[12 * 32];
          if (P.flags & 2) e10 = aa;
          if (P.flags & 1) { e11 = vv; e12 = sgn(vv); }
          if (P.flags & 4) e13 = 1.0;
        }
        if (t == tg) {  // first write of the pass (its values are in registers by now): the previous block must have left
          if (lane == 0) rq_bulk_wait_read<0>();
          __syncwarp();
          if (tg >= 1) {  // the body above the joint had entries in the pass before: zero them
            double2* z = reinterpret_cast<double2*>(stage + 14 * (tg - 1));
#pragma unroll
            for (int c2 = 0; c2 < 7; ++c2) z[c2] = make_double2(0.0, 0.0);
          }
        }
        double2* row = reinterpret_cast<double2*>(stage + 14 * t);
#pragma unroll
        for (int f = 0; f < 5; ++f) row[f] = make_double2(val[2 * f], val[2 * f + 1]);
        row[5] = make_double2(e10, e11);
        row[6] = make_double2(e12, e13);
      }
      // ONE request per (warp, row block): 32 x ncols doubles, contiguous in W and in the stage (one request per lane
      // and row was request-bound: 0.92 ms; two per lane 1.14 ms)
      rq_fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const int nrows = (int)(N - s0 < 32 ? N - s0 : 32);
        rq_bulk_s2g(W + ((int64_t)T->rowblock[g] * N + s0) * ldw, stage - (size_t)lane * P.pitch, (uint32_t)(nrows * ncols) * 8);
        rq_bulk_commit();
      }
    }
  }
  if (lane == 0) rq_bulk_wait_read<0>();  // the copies read this CTA's shared memory
}

}  // namespace fg
