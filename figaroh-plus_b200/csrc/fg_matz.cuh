// fg_matz.cuh - kernel (1) alone, second version: the tall regressor W written to HBM in the reference layout
// (row j*N+i, column order of Robot.get_standard_parameters; build_regressor_basic, regressor.py:20-194).
//
// What the first version showed (profiles/launches_r01_v4i.md: 4.03 GB in 4.36 ms = 0.93 TB/s, 14 % of the measured
// HBM copy bandwidth): four of 32 lanes ran the forward pass.  This one reuses the producers' walk of the ring
// kernel (rq_walk_tile: one sample per lane, the tree walked ONCE for all row blocks) and turns every (row block,
// body) pair into one 32-row x 14-column patch of W: the ten inertial entries s_g^T B_b and the body's
// Ia / fv / fs / off columns are staged in shared memory ([column][row], padded) and written out with lanes running
// along the 14 contiguous columns of a row (112-byte segments); pairs outside the row block's subtree are written as
// zero patches.  All 16 warps of every SM generate and store; the kernel is bound by the HBM write stream.
#pragma once
#include "fg_ringqr.cuh"
#include "fg_matz_params.h"

namespace fg {

struct MatzTables {          // one device blob, staged into shared memory by every CTA
  DevJoint pj[kMaxJoints];                // joints in walk order
  RqOp ops[2 * kMaxJoints];
  double S[kMaxRowGroups][6];             // motion axis of every row block
  int rowblock[kMaxRowGroups];
  int rowjoint[kMaxRowGroups];
  int body_of_t[kMaxJoints];              // walk index -> joint index (= body b, 1-based)
  int extra_grp[kMaxJoints];              // row group whose Ia / fv / fs / off columns of body b are non-zero; -1 all, -2 none
  unsigned char inertial[kMaxJoints];     // body has inertial columns (external-wrench mode skips massless bodies)
};

constexpr int MATZ_STAGE = 16 * 33;       // one patch: [column][row] padded to 33 rows; columns 0, 1 = the fs / off
                                          // columns of the body to the left (sector-aligned patches, see below)

// Patches and 32-byte sectors.  A body's 14 columns are 112 bytes: every other body starts in the middle of a sector,
// and a partially written sector costs a DRAM read-modify-write (first version of this kernel: 1.04 GB read for
// 0.19 GB of inputs, profiles/ncu_matz_r02_v1.md).  When the rows are sector aligned (W 32-byte aligned, ldw % 4 == 0)
// the patches are cut on sector boundaries instead: body k (0-based) even -> its first 12 columns (96 bytes), body k
// odd -> the fs / off columns of body k - 1 and its own 14 (128 bytes); an even last body writes all 14.
struct MatzSink {
  const MatzParams* P;
  const MatzTables* T;
  double* stage;
  double* W;
  int64_t ldw, N, s0;
  const double* vs;
  const double* as;
  bool valid, aligned;
  int lane, nrows;
  unsigned done;   // row groups that got a patch for the current body
  int body, mode;  // mode 0: 14 columns, 1: first 12, 2: 2 + 14
  int eg, egp;     // extra_grp of the body and of the body to its left
  double vv, aa, pv;

  template <int NC, int SC0>
  __device__ __forceinline__ void store_cols(double* Wp, bool zero) {
    if (NC == 16) {
      double* p = Wp + (int64_t)(lane >> 4) * ldw + (lane & 15);
      const double* sp = stage + (SC0 + (lane & 15)) * 33 + (lane >> 4);
      const int64_t step = 2 * ldw;
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        if (2 * it + (lane >> 4) < nrows) *p = zero ? 0.0 : sp[2 * it];
        p += step;
      }
    } else {
#pragma unroll
      for (int it = 0; it < NC; ++it) {
        const int e = it * 32 + lane;
        const int row = e / NC, col = e - NC * row;
        if (row < nrows) Wp[(int64_t)row * ldw + col] = zero ? 0.0 : stage[(SC0 + col) * 33 + row];
      }
    }
  }
  // one (row group, body) patch; val: the ten inertial entries in FIGAROH order or nullptr (outside the subtree)
#ifndef MATZ_NOINLINE_PATCH
  __device__ __forceinline__ void patch(int g, const double* val) {
#else
  __device__ __noinline__ void patch(int g, const double* val) {
#endif
    const bool own = eg == -1 || eg == g;                   // uniform
    const bool prv = mode == 2 && (egp == -1 || egp == g);  // uniform
    const bool any = val != nullptr || own || prv;
    if (any) {
      const bool inert = T->inertial[body] != 0;
#pragma unroll
      for (int f = 0; f < 10; ++f) stage[(2 + f) * 33 + lane] = (val != nullptr && valid && inert) ? val[f] : 0.0;
      // Ia, fv, fs, off of body b in row group g (regressor.py:55-70 joint torques, :142-169 external wrench: v[i, b])
      double e10 = 0.0, e11 = 0.0, e12 = 0.0, e13 = 0.0;
      if (valid && own) {
        if (P->flags & 2) e10 = aa;                       // FG_FLAG_ACTUATOR_INERTIA
        if (P->flags & 1) { e11 = vv; e12 = sgn(vv); }    // FG_FLAG_FRICTION
        if (P->flags & 4) e13 = 1.0;                      // FG_FLAG_JOINT_OFFSET
      }
      stage[12 * 33 + lane] = e10;
      stage[13 * 33 + lane] = e11;
      stage[14 * 33 + lane] = e12;
      stage[15 * 33 + lane] = e13;
      if (mode == 2) {
        stage[lane] = (valid && prv && (P->flags & 1)) ? sgn(pv) : 0.0;
        stage[33 + lane] = (valid && prv && (P->flags & 4)) ? 1.0 : 0.0;
      }
      __syncwarp();
    }
    double* Wp = W + ((int64_t)T->rowblock[g] * N + s0) * ldw + 14 * (body - 1);
    if (mode == 2) store_cols<16, 0>(Wp - 2, !any);
    else if (mode == 1) store_cols<12, 2>(Wp, !any);
    else store_cols<14, 2>(Wp, !any);
    if (any) __syncwarp();
    done |= 1u << g;
  }
  __device__ __forceinline__ void body_begin(const RqOp& op, double, double) {
    done = 0u;
    body = T->body_of_t[op.t];
    const int k = body - 1;
    mode = !aligned ? 0 : ((k & 1) ? 2 : (k + 1 < P->nj - 1 ? 1 : 0));
    eg = T->extra_grp[body];
    egp = k >= 1 ? T->extra_grp[body - 1] : -2;
    vv = aa = pv = 0.0;
    if (eg != -2) { vv = __ldg(vs + k); aa = __ldg(as + k); }
    if (mode == 2 && egp != -2) pv = __ldg(vs + k - 1);
  }
  __device__ __forceinline__ void emit(int g, int, const double* sl, const double* sa, const double* om, const double* al,
                                       const double* ap) {
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
    val[0] = sa[0] * al[0] + h[0] * om[0];                                   // Ixx
    val[1] = sa[0] * al[1] + sa[1] * al[0] + h[0] * om[1] + h[1] * om[0];   // Ixy
    val[3] = sa[1] * al[1] + h[1] * om[1];                                   // Iyy
    val[2] = sa[0] * al[2] + sa[2] * al[0] + h[0] * om[2] + h[2] * om[0];   // Ixz
    val[4] = sa[1] * al[2] + sa[2] * al[1] + h[1] * om[2] + h[2] * om[1];   // Iyz
    val[5] = sa[2] * al[2] + h[2] * om[2];                                   // Izz
    patch(g, val);
  }
  __device__ __forceinline__ void body_end(const RqOp&) {
    // row groups outside whose subtree this body lies: zero inertial entries; the body's own columns only where the
    // reference writes them
    for (int g = 0; g < P->ms; ++g)
      if (!((done >> g) & 1u)) patch(g, nullptr);
  }
};

#ifndef MATZ_WARPS
#define MATZ_WARPS 8   // 205 registers, no spills (A/B on B200: 16 warps at 128 registers with spills 2.26 ms, 12: 1.63, 8: 1.09)
#endif
__global__ void __launch_bounds__(MATZ_WARPS * 32, 1)
matz_kernel(const __grid_constant__ MatzParams P, const MatzTables* __restrict__ tab, const double* __restrict__ q,
            const double* __restrict__ qd, const double* __restrict__ qdd, int64_t N, double* __restrict__ W, int64_t ldw) {
  extern __shared__ __align__(16) double mz_smem[];
  MatzTables* T = reinterpret_cast<MatzTables*>(mz_smem);
  {
    const int* src = reinterpret_cast<const int*>(tab);
    int* dst = reinterpret_cast<int*>(T);
    for (int i = threadIdx.x; i < (int)(sizeof(MatzTables) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  constexpr int TAB_DOUBLES = (int)((sizeof(MatzTables) + 15) / 16) * 2;
  double* ps = mz_smem + TAB_DOUBLES + (size_t)warp * P.smem_per_warp;
  double* stage = ps + (P.ms * 6 + 12 * (P.nvslots + P.nmslots)) * 32;
  RqWalkCtx C;
  C.pj = T->pj; C.ops = T->ops; C.nops = P.nops; C.ms = P.ms; C.nq = P.nq; C.nv = P.nv; C.nvslots = P.nvslots;
  for (int k = 0; k < 3; ++k) C.gravity[k] = P.gravity[k];
  const int64_t ntiles = (N + 31) / 32;
  const int nb = P.nj - 1;
  for (int64_t tl = (int64_t)blockIdx.x * nw + warp; tl < ntiles; tl += (int64_t)gridDim.x * nw) {
    const int64_t s0 = tl * 32, smp = s0 + lane;
    const bool valid = smp < N;
    const int64_t sc = valid ? smp : N - 1;
    MatzSink sink;
    sink.P = &P; sink.T = T; sink.stage = stage; sink.W = W; sink.ldw = ldw; sink.N = N; sink.s0 = s0;
    sink.vs = qd + sc * P.nv; sink.as = qdd + sc * P.nv; sink.valid = valid; sink.lane = lane; sink.done = 0u; sink.body = 1;
    sink.aligned = (ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 31) == 0;
    sink.nrows = (int)(N - s0 < 32 ? N - s0 : 32);
    sink.mode = 0; sink.eg = sink.egp = -2; sink.vv = sink.aa = sink.pv = 0.0;
    rq_walk_tile(C, &T->S[0][0], 6, ps, q, qd, qdd, s0, N, lane, sink);
    if (P.coupling) {
      // add_coupling_TX40 (regressor.py:216-225): row block 4 <- [a5, v5, sign(v4+v5)], 5 <- [a4, v4, sign(v4+v5)]
      const double v4 = __ldg(sink.vs + 4), v5 = __ldg(sink.vs + 5), a4 = __ldg(sink.as + 4), a5 = __ldg(sink.as + 5);
      const double sg = sgn(v4 + v5);
      if (valid) {
        for (int g = 0; g < P.ms; ++g) {
          const int rb = T->rowblock[g];
          double* dst = W + ((int64_t)rb * N + smp) * ldw + 14 * nb;
          dst[0] = rb == 4 ? a5 : (rb == 5 ? a4 : 0.0);
          dst[1] = rb == 4 ? v5 : (rb == 5 ? v4 : 0.0);
          dst[2] = (rb == 4 || rb == 5) ? sg : 0.0;
        }
      }
    }
  }
}

}  // namespace fg
