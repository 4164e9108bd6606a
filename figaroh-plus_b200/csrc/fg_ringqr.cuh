// fg_ringqr.cuh - kernel (1)+(2), fifth generation: regressor producers and TSQR consumers in one persistent
// kernel, connected by an L2-resident ring of regressor tiles (design: fg_rq.h).
//
//   producer warp : one sample per lane; depth-first walk of the kinematic tree (descend / ascend program
//                   built by the host), motion axes of all live row groups kept in shared memory and
//                   transported frame to frame, entries s_g^T B_b of pinocchio::computeJointTorqueRegressor
//                   (reference call sites regressor.py:49,93) written for EVERY row group into the ring slot of
//                   the tile, in the consumers' tile layout (a warp store = one 256-byte row of a column).
//   consumer warp : one row group for the whole launch (as in generation 4); pulls its blocks of the tile
//                   from the ring with cp.async.bulk (completion on a per-warp mbarrier) and folds them into
//                   its private block-triangular factor (gq_fold_h).  Block k of the NEXT tile is requested as
//                   soon as panel k of the current one has retired.
#pragma once
#define FG_GQ_DEVICE_ONLY
#include "fg_groupqr.cuh"
#include "fg_rq.h"
#include "fg_async.cuh"

namespace fg {

constexpr unsigned RQ_SPIN_LIMIT = 1u << 22;  // bounded waits: ~1 s, then the error flag is raised
constexpr int RQ_ERR_CONSUMER = 1, RQ_ERR_PRODUCER = 2, RQ_ERR_MBAR = 3;

// lane 0 waits until *flag >= want (monotonic counters); false = gave up (error flag raised by the caller)
// (first version: __nanosleep(20..200) - at those values the warp does not sleep at all; the consumers' polls after
// a fold were 25 % of all executed instructions, each with an L1 invalidate: profiles/ncu_rq_ur10_r02.md)
__device__ __forceinline__ bool rq_spin_ge(const unsigned* flag, unsigned want, const unsigned* err, int sleep_ns) {
  for (unsigned it = 0; it < RQ_SPIN_LIMIT; ++it) {
    if (rq_ld_acquire(flag) >= want) return true;
    if ((it & 255u) == 255u && rq_ld_acquire(err) != 0u) return false;
    __nanosleep((unsigned)sleep_ns);
  }
  return false;
}

// ---------------------------------------------------------------------------------------------
// spatial transforms with M = (R, p) the child-in-parent placement
//   actInv (into the child frame):  ang' = R^T ang,  lin' = R^T (lin - p x ang)
//   act    (back to the parent):    ang' = R ang,    lin' = R lin + p x (R ang)
__device__ __forceinline__ void rq_actinv(const double* R, const double* p, const double* lin, const double* ang,
                                          double* lo, double* ao) {
  double t[3];
  cross3(p, ang, t);
  for (int k = 0; k < 3; ++k) t[k] = lin[k] - t[k];
  for (int k = 0; k < 3; ++k) {
    lo[k] = R[k] * t[0] + R[3 + k] * t[1] + R[6 + k] * t[2];
    ao[k] = R[k] * ang[0] + R[3 + k] * ang[1] + R[6 + k] * ang[2];
  }
}
__device__ __forceinline__ void rq_act(const double* R, const double* p, const double* lin, const double* ang,
                                       double* lo, double* ao) {
  double ra[3], rl[3], x[3];
  for (int k = 0; k < 3; ++k) {
    ra[k] = R[3 * k] * ang[0] + R[3 * k + 1] * ang[1] + R[3 * k + 2] * ang[2];
    rl[k] = R[3 * k] * lin[0] + R[3 * k + 1] * lin[1] + R[3 * k + 2] * lin[2];
  }
  cross3(p, ra, x);
  for (int k = 0; k < 3; ++k) { ao[k] = ra[k]; lo[k] = rl[k] + x[k]; }
}

// entries of (row group, body) into the ring slot (same arithmetic as gq_emit; padding samples store zeros)
__device__ __forceinline__ void rq_emit(double* gslot, int ed, const double* sl, const double* sa, const double* om,
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
  // column `slot` of the tile starts at slot * 32 doubles ((slot >> 3) * GQ_BLK + (slot & 7) * 32); the row
  // swizzle tq_sw(c) of column c = slot & 7 takes the values 0, 5, 8, 13, 5, 8, 13, 0
  int slot = ed >> 10;
  double* dst = gslot + slot * 32;
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    if ((ed >> k) & 1) {  // uniform
      dst[lane ^ tq_sw(slot & 7)] = valid ? val[k] : 0.0;  // (measured and rejected: L2 evict-last on the ring stores)
      ++slot;
      dst += 32;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// The producers' depth-first walk over the kinematic tree for one tile (32 samples, one per lane).
// ps: this warp's shared memory: axes [ms][6][32], saved (v, a) [nvslots][12][32], saved joint transforms
// [nmslots][12][32].  What happens to the entries s_g^T B_b is the Sink's business:
//   sink.body_begin(op, cv0, ca0)                     before the entries of the body of a DESCEND step
//   sink.emit(g, t, sl, sa, om, al, ap)               row group g (axis [sl; sa] in the body frame), walk index t
//   sink.body_end(op)                                 after them
struct RqWalkCtx {
  const DevJoint* pj;   // joints in walk order
  const RqOp* ops;
  int nops, ms, nq, nv, nvslots;
  double gravity[3];
};

// SIN: the inputs were staged (q, qd, qdd point to a tile in shared memory, s0 = 0, N = rows of the tile)
template <bool SIN>
__device__ __forceinline__ double rq_ldin(const double* p) {
  if (SIN) return *p;
  return __ldg(p);
}
template <bool SIN, class Sink>
static __device__ __forceinline__ void rq_walk_tile(const RqWalkCtx& C, const double* Sbase, int Sstride /* doubles */,
                                                    double* ps, const double* __restrict__ q,
                                                    const double* __restrict__ qd, const double* __restrict__ qdd,
                                                    int64_t s0, int64_t N, int lane, Sink& sink) {
  const int ms = C.ms;
  const int64_t smp = s0 + lane;
  const bool valid = smp < N;
  const int64_t sc = valid ? smp : N - 1;
  const double* qs = q + sc * C.nq;
  const double* vs = qd + sc * C.nv;
  const double* as = qdd + sc * C.nv;
  double* AX = ps + lane;                         // AX[(g * 6 + k) * 32]
  double* VS = ps + ms * 6 * 32 + lane;           // VS[(slot * 12 + k) * 32]
  double* MS = VS + C.nvslots * 12 * 32;          // MS[(slot * 12 + k) * 32]
  double v[6] = {0, 0, 0, 0, 0, 0}, a[6] = {0, 0, 0, 0, 0, 0};
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
  // inputs of the first joint; afterwards requested one descend ahead
  double nq0 = 0.0, nv0 = 0.0, na0 = 0.0;
  {
    const DevJoint& j0 = C.pj[C.ops[0].t];
    nq0 = rq_ldin<SIN>(qs + j0.idx_q); nv0 = rq_ldin<SIN>(vs + j0.idx_v); na0 = rq_ldin<SIN>(as + j0.idx_v);
  }
  const int nops = C.nops;
  for (int o = 0; o < nops; ++o) {
    const RqOp op = C.ops[o];
    if (op.kind == 0) {
      const DevJoint& jt = C.pj[op.t];
      if (op.vsave != 255) {
        for (int k = 0; k < 6; ++k) { VS[(op.vsave * 12 + k) * 32] = v[k]; VS[(op.vsave * 12 + 6 + k) * 32] = a[k]; }
      }
      if (op.from_universe) {
        for (int k = 0; k < 6; ++k) v[k] = 0.0;
        a[0] = -C.gravity[0]; a[1] = -C.gravity[1]; a[2] = -C.gravity[2];
        a[3] = a[4] = a[5] = 0.0;
      }
      const double cq0 = nq0, cv0 = nv0, ca0 = na0;
      if (op.pad != 255) {  // the next descend's joint
        const DevJoint& jn = C.pj[op.pad];
        nq0 = rq_ldin<SIN>(qs + jn.idx_q); nv0 = rq_ldin<SIN>(vs + jn.idx_v); na0 = rq_ldin<SIN>(as + jn.idx_v);
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
        axis_rot(ax, cq0, rq_ldin<SIN>(qs + iq + 1), Rj);
        for (int k = 0; k < 3; ++k) { vj[3 + k] = ax[k] * cv0; aj[3 + k] = ax[k] * ca0; }
      } else if (jt.jtype == 3) {
        for (int k = 0; k < 3; ++k) { pj[k] = ax[k] * cq0; vj[k] = ax[k] * cv0; aj[k] = ax[k] * ca0; }
      } else {  // free-flyer
        quat_rot(rq_ldin<SIN>(qs + iq + 3), rq_ldin<SIN>(qs + iq + 4), rq_ldin<SIN>(qs + iq + 5), rq_ldin<SIN>(qs + iq + 6), Rj);
        for (int k = 0; k < 3; ++k) pj[k] = rq_ldin<SIN>(qs + iq + k);
        for (int k = 0; k < 6; ++k) { vj[k] = rq_ldin<SIN>(vs + iv + k); aj[k] = rq_ldin<SIN>(as + iv + k); }
      }
      // liMi = placement * Mj
      for (int r = 0; r < 3; ++r) {
        const double m0 = jt.plcR[3 * r], m1 = jt.plcR[3 * r + 1], m2 = jt.plcR[3 * r + 2];
        for (int c = 0; c < 3; ++c) R[3 * r + c] = m0 * Rj[c] + m1 * Rj[3 + c] + m2 * Rj[6 + c];
        p[r] = m0 * pj[0] + m1 * pj[1] + m2 * pj[2] + jt.plcp[r];
      }
      if (op.msave != 255) {
        for (int k = 0; k < 9; ++k) MS[(op.msave * 12 + k) * 32] = R[k];
        for (int k = 0; k < 3; ++k) MS[(op.msave * 12 + 9 + k) * 32] = p[k];
      }
      // v = vJ + liMi.actInv(v_parent);  a = v x vJ + S qdd + liMi.actInv(a_parent)
      double vn[6], an[6], tl[3], ta[3];
      rq_actinv(R, p, v, v + 3, tl, ta);
      for (int k = 0; k < 3; ++k) { vn[k] = vj[k] + tl[k]; vn[3 + k] = vj[3 + k] + ta[k]; }
      double l1[3], l2[3], l3[3];
      cross3(vn + 3, vj, l1);
      cross3(vn, vj + 3, l2);
      cross3(vn + 3, vj + 3, l3);
      rq_actinv(R, p, a, a + 3, tl, ta);
      for (int k = 0; k < 3; ++k) {
        an[k] = l1[k] + l2[k] + aj[k] + tl[k];
        an[3 + k] = l3[k] + aj[3 + k] + ta[k];
      }
      for (int k = 0; k < 6; ++k) { v[k] = vn[k]; a[k] = an[k]; }
      double ap[3];
      cross3(v + 3, v, ap);  // classical acceleration a' = a_lin + w x v_lin
      for (int k = 0; k < 3; ++k) ap[k] += a[k];
      sink.body_begin(op, cv0, ca0);
      // row groups whose axis is live: transport into this frame and emit; then the groups that start here
      for (unsigned m = op.live; m; m &= m - 1) {
        const int g = __ffs(m) - 1;
        double sl[3], sa[3], nl[3], na[3];
        for (int k = 0; k < 3; ++k) { sl[k] = AX[(g * 6 + k) * 32]; sa[k] = AX[(g * 6 + 3 + k) * 32]; }
        rq_actinv(R, p, sl, sa, nl, na);
        for (int k = 0; k < 3; ++k) { AX[(g * 6 + k) * 32] = nl[k]; AX[(g * 6 + 3 + k) * 32] = na[k]; }
        sink.emit(g, op.t, nl, na, v + 3, a + 3, ap);
      }
      for (unsigned m = op.own; m; m &= m - 1) {
        const int g = __ffs(m) - 1;
        double sl[3], sa[3];
        for (int k = 0; k < 3; ++k) { sl[k] = Sbase[g * Sstride + k]; sa[k] = Sbase[g * Sstride + 3 + k]; }
        for (int k = 0; k < 3; ++k) { AX[(g * 6 + k) * 32] = sl[k]; AX[(g * 6 + 3 + k) * 32] = sa[k]; }
        sink.emit(g, op.t, sl, sa, v + 3, a + 3, ap);
      }
      sink.body_end(op);
    } else {
      for (int k = 0; k < 9; ++k) R[k] = MS[(op.mslot * 12 + k) * 32];
      for (int k = 0; k < 3; ++k) p[k] = MS[(op.mslot * 12 + 9 + k) * 32];
      for (unsigned m = op.live; m; m &= m - 1) {
        const int g = __ffs(m) - 1;
        double sl[3], sa[3], nl[3], na[3];
        for (int k = 0; k < 3; ++k) { sl[k] = AX[(g * 6 + k) * 32]; sa[k] = AX[(g * 6 + 3 + k) * 32]; }
        rq_act(R, p, sl, sa, nl, na);
        for (int k = 0; k < 3; ++k) { AX[(g * 6 + k) * 32] = nl[k]; AX[(g * 6 + 3 + k) * 32] = na[k]; }
      }
      if (op.vload != 255) {
        for (int k = 0; k < 6; ++k) { v[k] = VS[(op.vload * 12 + k) * 32]; a[k] = VS[(op.vload * 12 + 6 + k) * 32]; }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// One tile (32 samples, one per lane) of EVERY row group into a ring slot.
struct RqRingSink {
  const RqParams* P;
  const GqGen* S;
  const RqStepTab* ST;
  double* slot;
  bool valid;
  int lane;
  // the joint's own Ia / fv / fs / off columns (regressor.py:55-70): its (qd, qdd) are in registers here; at the end
  // of the tile each of them cost a dependent load (a fifth of the producers' time, profiles/ncu_rq_ur10_r02.md)
  __device__ __forceinline__ void body_begin(const RqOp& op, double cv0, double ca0) {
    const int n = ST->n[op.t];
    for (int k = 0; k < n; ++k) {
      const int e = ST->col[op.t][k];
      const int col = e >> 2, src = e & 3;
      const double val = src == SRC_QDD ? ca0 : (src == SRC_QD ? cv0 : (src == SRC_SIGN_QD ? sgn(cv0) : 1.0));
      slot[col * 32 + (lane ^ tq_sw(col & 7))] = valid ? val : 0.0;
    }
  }
  __device__ __forceinline__ void body_end(const RqOp&) {}
  __device__ __forceinline__ void emit(int g, int t, const double* sl, const double* sa, const double* om, const double* al,
                                       const double* ap) {
    const int ed = S->emit[g][t];
    if (ed) rq_emit(slot + P->goff[g], ed, sl, sa, om, al, ap, valid, lane);
  }
};

static __device__ __forceinline__ void rq_produce_tile(const RqParams& P, const GqGen& S, const RqStepTab& ST,
                                                       const RqOp* ops, double* ps,
                                                       double* slot, const double* __restrict__ q,
                                                       const double* __restrict__ qd, const double* __restrict__ qdd,
                                                       const double* __restrict__ tau, int64_t ld_tau, int64_t s0,
                                                       int64_t N, int lane) {
  const int ms = P.gq.ms;
  const int64_t smp = s0 + lane;
  const bool valid = smp < N;
  const int64_t sc = valid ? smp : N - 1;
  const double* vs = qd + sc * P.gq.nv;
  const double* as = qdd + sc * P.gq.nv;
  RqWalkCtx C;
  C.pj = S.pj; C.ops = ops; C.nops = P.nops; C.ms = ms; C.nq = P.gq.nq; C.nv = P.gq.nv; C.nvslots = P.nvslots;
  for (int k = 0; k < 3; ++k) C.gravity[k] = P.gq.gravity[k];
  static_assert(sizeof(GqGroup) % 8 == 0, "GqGroup must be a whole number of doubles");
  RqRingSink sink{&P, &S, &ST, slot, valid, lane};
  rq_walk_tile<false>(C, S.grp[0].S, (int)(sizeof(GqGroup) / 8), ps, q, qd, qdd, s0, N, lane, sink);
  // columns read straight off the inputs that no single-dof joint's walk step took (rhs, coupling, free-flyer dofs)
  // (measured and rejected: L2 evict-first loads of the inputs, 3.57 ms against 3.47; rings of 768 / 1024 slots)
  // (measured and rejected: issuing the loads of eight columns together before consuming them - 4.49 ms against 3.48)
  for (int g = 0; g < ms; ++g) {
    const GqGroup& G = S.grp[g];
    double* gs = slot + P.goff[g];
    for (int t = 0; t < G.nsimple; ++t) {
      const GqSimple s1 = S.simple[g][t];
      const int ix = s1.slot * 32 + (lane ^ tq_sw(s1.slot & 7));
      double val = 0.0;
      if (s1.kind == COL_RHS) {
        if (valid && tau != nullptr) val = __ldg(tau + (int64_t)g * ld_tau + smp);
      } else if (s1.kind == COL_EXTRA) {
        const double vv = __ldg(vs + s1.vidx), aa = __ldg(as + s1.vidx);
        val = s1.src == SRC_QDD ? aa : (s1.src == SRC_QD ? vv : (s1.src == SRC_SIGN_QD ? sgn(vv) : 1.0));
      } else {
        const double v4 = __ldg(vs + 4), v5 = __ldg(vs + 5), a4 = __ldg(as + 4), a5 = __ldg(as + 5);
        const double sg = sgn(v4 + v5);
        val = G.rowblock == 4 ? (s1.src == 0 ? a5 : (s1.src == 1 ? v5 : sg)) : (s1.src == 0 ? a4 : (s1.src == 1 ? v4 : sg));
      }
      gs[ix] = valid ? val : 0.0;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// consumer-side refill of the shared-memory tile with the next tile's blocks (called from gq_fold_h)
struct RqRefill {
  bool has_next;
  bool armed;
  int issued, nb, lane;
  const unsigned* ready;
  unsigned want;
  const double* src;
  double* dst;
  uint64_t* bar;
  __device__ __forceinline__ void arm() {
    armed = true;
    rq_mbar_expect_tx(bar, (uint32_t)nb * GQ_BLK * 8);
  }
  __device__ __forceinline__ void issue_upto(int k) {
    while (issued <= k) {
      rq_bulk_g2s(dst + issued * GQ_BLK, src + issued * GQ_BLK, GQ_BLK * 8, bar);
      ++issued;
    }
  }
  __device__ __forceinline__ void operator()(int k) {
    if (!has_next) return;  // uniform
    rq_fence_proxy_async_smem();  // every lane: its generic-proxy writes to block k precede the bulk copy
    __syncwarp();
    if (lane == 0) {
      if (!armed && rq_ld_acquire(ready) >= want) {
        rq_fence_proxy_async();
        arm();
      }
      if (armed) issue_upto(k);
    }
  }
};

// ---------------------------------------------------------------------------------------------
// ctl: ready[ring] | done[ring] | err
// MAXNW: warps per CTA the kernel is compiled for (16: 128 registers per thread, 20: 96)
// PROF: count the cycles of every fold / produced tile / wait into prof (calibration launches only)
template <int MAXNW, bool PROF>
__global__ void __launch_bounds__(MAXNW * 32, 1)
rq_accumulate_kernel(const __grid_constant__ RqParams P, const unsigned char* __restrict__ tab_g,
                     const unsigned short* __restrict__ tab_i, const double* __restrict__ q,
                     const double* __restrict__ qd, const double* __restrict__ qdd, const double* __restrict__ tau,
                     int64_t ld_tau, int64_t N, double* __restrict__ r_all, double* __restrict__ ring,
                     unsigned* __restrict__ ctl, int first, unsigned long long* __restrict__ prof) {
  extern __shared__ __align__(16) double gq_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  GqGen* Sg = reinterpret_cast<GqGen*>(gq_smem);
  RqOp* ops = reinterpret_cast<RqOp*>(gq_smem + GQ_GEN_DOUBLES);
  constexpr int OPS_DOUBLES = (int)(sizeof(RqOp) * RQ_MAXOPS / 8) + RQ_STEP_DOUBLES;
  RqStepTab* STs = reinterpret_cast<RqStepTab*>(gq_smem + GQ_GEN_DOUBLES + (int)(sizeof(RqOp) * RQ_MAXOPS / 8));
  {
    const int* ss = reinterpret_cast<const int*>(&P.step);
    int* ds = reinterpret_cast<int*>(STs);
    for (int i = threadIdx.x; i < (int)(sizeof(RqStepTab) / 4); i += blockDim.x) ds[i] = ss[i];
    const int* src = reinterpret_cast<const int*>(&P.gq.gen);
    int* dst = reinterpret_cast<int*>(Sg);
    for (int i = threadIdx.x; i < (int)(sizeof(GqGen) / 4); i += blockDim.x) dst[i] = src[i];
    const int* so = reinterpret_cast<const int*>(P.ops);
    int* dop = reinterpret_cast<int*>(ops);
    for (int i = threadIdx.x; i < (int)(sizeof(RqOp) * RQ_MAXOPS / 4); i += blockDim.x) dop[i] = so[i];
  }
  const int gw = blockIdx.x * nw + warp;
  const int role = tab_g[gw], inst = tab_i[gw];
  // this warp's shared memory: consumers [tile nb*GQ_BLK | scratch GQ_SCRATCH | mbarrier 2], producers [prod_smem]
  int off = GQ_GEN_DOUBLES + OPS_DOUBLES;
  for (int w = 0; w < warp; ++w) {
    const int rw = tab_g[blockIdx.x * nw + w];
    off += rw == 255 ? P.prod_smem_doubles : P.gq.gen.grp[rw].nb * GQ_BLK + GQ_SCRATCH + 2;
  }
  double* mine = gq_smem + off;
  unsigned* ready = ctl;
  unsigned* done = ctl + P.ring;
  unsigned* err = ctl + 2 * P.ring;
  uint64_t* bar = nullptr;
  if (role != 255) {
    bar = reinterpret_cast<uint64_t*>(mine + P.gq.gen.grp[role].nb * GQ_BLK + GQ_SCRATCH);
    if (lane == 0) {
      rq_mbar_init(bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      rq_fence_proxy_async();
    }
  }
  __syncthreads();
  const GqGen& S = *Sg;
  const int64_t ntiles = (N + GQ_S - 1) / GQ_S;
  const int ms = P.gq.ms;

  long long pf_work = 0, pf_wait = 0, pf_t = 0;
  unsigned pf_n = 0;
  if (role == 255) {
    // ------------------------------------------------------------------ producer
    for (int64_t tl = inst; tl < ntiles; tl += P.nprod) {
      if (PROF) pf_t = clock64();
      const int s = (int)(tl % P.ring);
      const unsigned round = (unsigned)(tl / P.ring);
      if (round > 0) {  // every row group must have copied the tile that used this slot before
        unsigned ok = 1;
        if (lane == 0) ok = rq_spin_ge(done + s, (unsigned)ms * round, err, P.sleep_ns) ? 1u : 0u;
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (!ok) {
          if (lane == 0) atomicCAS(err, 0u, (unsigned)RQ_ERR_PRODUCER);
          return;
        }
      }
      if (PROF) { const long long t = clock64(); pf_wait += t - pf_t; pf_t = t; }
      if (tl + P.nprod < ntiles) {  // pull the next tile's inputs towards L2
        const int64_t sn = (tl + P.nprod) * GQ_S + lane;
        if (sn < N) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(q + sn * P.gq.nq));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(qd + sn * P.gq.nv));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(qdd + sn * P.gq.nv));
        }
        if (tau != nullptr && lane < 2 * ms) {  // the 256 bytes of every row group's rhs: two lines
          const int64_t st = (tl + P.nprod) * GQ_S + 16 * (lane & 1);
          if (st < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(tau + (int64_t)(lane >> 1) * ld_tau + st));
        }
      }
      rq_produce_tile(P, S, *STs, ops, mine, ring + (size_t)s * P.slot_doubles, q, qd, qdd, tau, ld_tau, tl * GQ_S, N, lane);
      __threadfence();
      __syncwarp();
      if (lane == 0) rq_st_release(ready + s, (unsigned)(tl + 1));
      if (PROF) { pf_work += clock64() - pf_t; ++pf_n; }
    }
    if (PROF && lane == 0 && pf_n) {
      atomicAdd(prof + 24, (unsigned long long)pf_work);
      atomicAdd(prof + 25, (unsigned long long)pf_n);
      atomicAdd(prof + 26, (unsigned long long)pf_wait);
    }
    return;
  }

  // -------------------------------------------------------------------- consumer of row group `role`
  const int g = role;
  const GqGroup& G = S.grp[g];
  const int nb = G.nb, ninst = G.ninst;
  double* tile = mine;
  double* scr = tile + nb * GQ_BLK;
  double* Rg = r_all + (size_t)G.r_off + (size_t)inst * G.r_size;
  if (first) {
    for (int i = lane; i < G.r_size; i += 32) Rg[i] = 0.0;
    __syncwarp();
  }
  GqOffs offs;
  gq_make_offs(offs, lane);
  const int goff = P.goff[g];
  uint32_t parity = 0;
  int64_t tl = inst;
  if (tl < ntiles) {  // first tile: wait, then request every block
    const int s = (int)(tl % P.ring);
    unsigned ok = 1;
    if (lane == 0) {
      ok = rq_spin_ge(ready + s, (unsigned)(tl + 1), err, P.sleep_ns) ? 1u : 0u;
      if (ok) {
        rq_fence_proxy_async();
        rq_mbar_expect_tx(bar, (uint32_t)nb * GQ_BLK * 8);
        const double* src = ring + (size_t)s * P.slot_doubles + goff;
        for (int b = 0; b < nb; ++b) rq_bulk_g2s(tile + b * GQ_BLK, src + b * GQ_BLK, GQ_BLK * 8, bar);
      }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (!ok) {
      if (lane == 0) atomicCAS(err, 0u, (unsigned)RQ_ERR_CONSUMER);
      return;
    }
  }
  for (; tl < ntiles; tl += ninst) {
    if (PROF) pf_t = clock64();
    {  // the tile's bytes have landed
      unsigned it = 0;
      while (!rq_mbar_try_wait(bar, parity)) {
        if (++it > RQ_SPIN_LIMIT) {
          if (lane == 0) atomicCAS(err, 0u, (unsigned)RQ_ERR_MBAR);
          return;
        }
      }
      parity ^= 1u;
    }
    if (PROF) { const long long t = clock64(); pf_wait += t - pf_t; pf_t = t; }
    const int s = (int)(tl % P.ring);
    if (lane == 0) atomicAdd(done + s, 1u);  // the copy has read the slot: it may be refilled
    const int64_t nx = tl + ninst;
    RqRefill rf;
    rf.has_next = nx < ntiles;
    rf.armed = false;
    rf.issued = 0;
    rf.nb = nb;
    rf.lane = lane;
    const int sn = (int)(nx % P.ring);
    rf.ready = ready + sn;
    rf.want = (unsigned)(nx + 1);
    rf.src = ring + (size_t)sn * P.slot_doubles + goff;
    rf.dst = tile;
    rf.bar = bar;
    gq_fold_h<false, 1, RqRefill>(tile, nb, Rg, 0, scr, offs, lane, rf);
    if (PROF) { const long long t = clock64(); pf_work += t - pf_t; pf_t = t; ++pf_n; }
    if (rf.has_next) {
      unsigned ok = 1;
      if (lane == 0) {
        if (!rf.armed) {
          ok = rq_spin_ge(rf.ready, rf.want, err, P.sleep_ns) ? 1u : 0u;
          if (ok) {
            rq_fence_proxy_async();
            rf.arm();
          }
        }
        if (ok) rf.issue_upto(nb - 1);
      }
      ok = __shfl_sync(0xffffffffu, ok, 0);
      if (!ok) {
        if (lane == 0) atomicCAS(err, 0u, (unsigned)RQ_ERR_CONSUMER);
        return;
      }
      if (PROF) pf_wait += clock64() - pf_t;
    }
  }
  if (PROF && lane == 0 && pf_n) {
    atomicAdd(prof + g, (unsigned long long)pf_work);
    atomicAdd(prof + 8 + g, (unsigned long long)pf_n);
    atomicAdd(prof + 16 + g, (unsigned long long)pf_wait);
  }
}

}  // namespace fg
