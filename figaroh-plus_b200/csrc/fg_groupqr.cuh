// fg_groupqr.cuh - the row-group TSQR stream: Householder panel, compact-WY factor, DMMA trailing update and
// the fold of a 32-row (or 64-row) tile into a block-triangular factor, plus the merge of the per-stream factors.
//
// Rows of different row groups are different rows of W, hence R = qr([R_0; ...; R_{ms-1}]) with R_g the factor of
// row group g's own non-zero columns (fg_gq.h).  A warp works on ONE row group for the whole launch and folds
// 32-sample tiles of that group's entries (pinocchio::computeJointTorqueRegressor rows, reference call sites
// regressor.py:49,93) into its private block-triangular R_g (L2-resident scratch) by a blocked Householder QR:
// 8-column panels with one row per lane, compact-WY T, DMMA trailing updates whose reflector fragments stay in
// registers for the whole panel.  No warp ever waits for another one.  The tiles come from the producers of
// fg_ringqr.cuh (generation 5); generation 4 generated them inside the stream (one forward pass per row group
// and tile - 39 % of the kernel, profiles/ncu_gq_r01_v4i.md).
// The per-warp factors are merged by gq_merge_kernel (pairwise tree per row group driven by arrival counters,
// then the ms roots are folded into the full factor in state column order).
#pragma once
#include "fg_tile.cuh"
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
// NC: real columns of the block (the columns NC..7 are zero padding: the last block of a row group; their column
// steps are skipped - 21 of UR10's 248 column steps per tile are padding)
template <int RPL, int NC = 8>
__device__ __forceinline__ void gq_panel_factor(double (&x)[RPL][8], const double* Rin, double* Rout, double* sc,
                                                double* Pb, int lane) {
  const int ci = lane >> 2, part = lane & 3;
  int rd[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) rd[k] = ci * 32 + ((8 * part + 2 * ((k + (part >> 1)) & 3) + 4 * ci) & 31);
  double* dbuf = sc + 8;
  double rw[8];  // updated row j - 1 of the diagonal block (lane-uniform), stored once every lane has read it
#pragma unroll
  for (int c = NC; c < 8; ++c) rw[c] = 0.0;
  if (NC < 8 && lane >= NC && lane < 8) sc[lane] = 0.0;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
#pragma unroll
    for (int c = j; c < NC; ++c) {
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
    for (int c = (j & ~1); c < NC; c += 2) {
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
    for (int c = j + 1; c < NC; ++c) {
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
  if (lane == 0) {
#pragma unroll
    for (int c = NC - 1; c < 8; ++c) Rout[(NC - 1) * 8 + c] = rw[c];
  }
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
    // (measured and rejected in round 2: a second instantiation gq_panel_factor<RPL, 4> for last blocks with <= 4 real
    // columns - 21 of UR10's 248 column steps per tile are padding - 4.51 ms against 3.44 ms: the extra code evicts
    // the panel from the instruction cache, as the round-1 attempt with uniform branches did)
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

#ifndef FG_GQ_DEVICE_ONLY  // fg_rq.cu reuses the device functions above, not the merge kernel
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
