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

#ifndef MATZ_NBUF
#define MATZ_NBUF 2                        // stage buffers per warp (rotating)
#endif
constexpr int MATZ_ROW = 18;               // doubles per staged row: 16 columns + 2 of padding (144 bytes: the
                                           // 16-byte column pairs of the 32 rows fall on 8 bank groups x 4 rows, the
                                           // optimum for a 512-byte warp store; rows stay contiguous for bulk copies)
constexpr int MATZ_STAGE = MATZ_NBUF * 32 * MATZ_ROW + MATZ_ROW;   // buffers + one row of zeros

// Patches and 32-byte sectors.  A body's 14 columns are 112 bytes: every other body starts in the middle of a sector,
// and a partially written sector costs a DRAM read-modify-write (first version of this kernel: 1.04 GB read for
// 0.19 GB of inputs, profiles/ncu_matz_r02_v1.md).  When the rows are sector aligned (W 32-byte aligned, ldw % 4 == 0)
// the patches are cut on sector boundaries instead: body k (0-based) even -> its first 12 columns (96 bytes), body k
// odd -> the fs / off columns of body k - 1 and its own 14 (128 bytes); an even last body writes all 14.
// A staged row holds 8 column pairs: (fs, off of the body to the left) | Ixx Ixy | Ixz Iyy | Iyz Izz | mx my | mz m |
// Ia fv | fs off.  Every lane stages its own sample's row with 16-byte stores; the patch leaves either as 16-byte
// global stores (8 lanes per row, 4 rows per instruction) or, with MATZ_BULK, as one cp.async.bulk per row issued by
// the lane that staged it (no cross-lane dependency at all).
struct MatzSink {
  const MatzParams* P;
  const MatzTables* T;
  double* stage;   // MATZ_NBUF buffers of 32 rows, then the zero row
  double* W;
  int64_t ldw, N, s0;
  const double* vs;
  const double* as;
  bool valid, aligned, aligned16;
  int lane, nrows, buf;
  unsigned done;   // row groups that got a patch for the current body
  int body, mode;  // mode 0: 14 columns, 1: first 12, 2: 2 + 14
  int eg, egp;     // extra_grp of the body and of the body to its left
  double vv, aa, pv;

  // NC2 column pairs starting at pair C20 of the staged rows -> Wp (first column of the patch, row 0 of the tile)
  template <int NC2, int C20>
  __device__ __forceinline__ void store_pairs(const double* sb, double* Wp, bool zero) {
#ifdef MATZ_BULK
    if (lane < nrows) {
      const double* src = zero ? stage + MATZ_NBUF * 32 * MATZ_ROW : sb + lane * MATZ_ROW + 2 * C20;
      rq_bulk_s2g(Wp + (int64_t)lane * ldw, src, NC2 * 16);
    }
    rq_bulk_commit();
#else
    const int r0 = lane >> 3, c2 = lane & 7;
    double2* p = reinterpret_cast<double2*>(Wp + (int64_t)r0 * ldw) + c2;
    const double2* sp = reinterpret_cast<const double2*>(sb + r0 * MATZ_ROW) + C20 + c2;
    const int64_t step = 2 * ldw;  // 4 rows, in double2 units
    if (c2 < NC2) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        // L2 evict-last: the row is completed by the patches of the bodies that follow; lines that leave L2 half
        // written cost DRAM efficiency (A/B, 6 warps: 0.99 ms with the hint, 1.09 ms without)
#ifndef MATZ_NO_EVICT_LAST
        if (4 * it + r0 < nrows) rq_st_keep(p, zero ? make_double2(0.0, 0.0) : sp[it * (4 * MATZ_ROW / 2)]);
#else
        if (4 * it + r0 < nrows) *p = zero ? make_double2(0.0, 0.0) : sp[it * (4 * MATZ_ROW / 2)];
#endif
        p += step;
      }
    }
#endif
  }
  // rows that are not 16-byte aligned (odd ldw: the TX40 coupling columns): 14 columns with 8-byte stores
  __device__ __forceinline__ void store_scalar(const double* sb, double* Wp, bool zero) {
#pragma unroll 2
    for (int it = 0; it < 14; ++it) {
      const int e = it * 32 + lane;
      const int row = e / 14, col = e - 14 * row;
      if (row < nrows) Wp[(int64_t)row * ldw + col] = zero ? 0.0 : sb[row * MATZ_ROW + 2 + col];
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
    double* sb = stage + buf * (32 * MATZ_ROW);
    if (any) {
#ifdef MATZ_BULK
      rq_bulk_wait_read<MATZ_NBUF - 1>();  // this lane's copy out of the buffer it is about to overwrite has been read
#endif
      double2* row = reinterpret_cast<double2*>(sb + lane * MATZ_ROW);
      const bool inert = val != nullptr && valid && T->inertial[body] != 0;
#pragma unroll
      for (int f = 0; f < 5; ++f) row[1 + f] = inert ? make_double2(val[2 * f], val[2 * f + 1]) : make_double2(0.0, 0.0);
      // Ia, fv, fs, off of body b in row group g (regressor.py:55-70 joint torques, :142-169 external wrench: v[i, b])
      double e10 = 0.0, e11 = 0.0, e12 = 0.0, e13 = 0.0;
      if (valid && own) {
        if (P->flags & 2) e10 = aa;                       // FG_FLAG_ACTUATOR_INERTIA
        if (P->flags & 1) { e11 = vv; e12 = sgn(vv); }    // FG_FLAG_FRICTION
        if (P->flags & 4) e13 = 1.0;                      // FG_FLAG_JOINT_OFFSET
      }
      row[6] = make_double2(e10, e11);
      row[7] = make_double2(e12, e13);
      if (mode == 2)
        row[0] = make_double2((valid && prv && (P->flags & 1)) ? sgn(pv) : 0.0, (valid && prv && (P->flags & 4)) ? 1.0 : 0.0);
#ifdef MATZ_BULK
      rq_fence_proxy_async_smem();
#else
      __syncwarp();
#endif
      buf = buf + 1 == MATZ_NBUF ? 0 : buf + 1;
    }
    double* Wp = W + ((int64_t)T->rowblock[g] * N + s0) * ldw + 14 * (body - 1);
    if (mode == 2) store_pairs<8, 0>(sb, Wp - 2, !any);
    else if (mode == 1) store_pairs<6, 1>(sb, Wp, !any);
    else if (aligned16) store_pairs<7, 1>(sb, Wp, !any);
    else store_scalar(sb, Wp, !any);
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
    if (eg != -2) { vv = vs[k]; aa = as[k]; }
    if (mode == 2 && egp != -2) pv = vs[k - 1];
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

#ifndef MATZG_WARPS
#define MATZG_WARPS 5   // fg_matzg.cuh: 42 KB of kinematics + stage per warp
#endif
#ifndef MATZ_WARPS
#define MATZ_WARPS 6   // spill-free (A/B on B200, UR10 1 M samples: 16 warps at 128 registers with spills 2.26 ms, 12: 1.63, 8: 1.09;
                       // with evict-last stores 4: 1.27, 5: 1.23, 6: 0.99, 7: 1.04, 8: 1.04, 10: 1.18 - the rows of a tile are
                       // completed over six patch passes: warps x 148 SMs x 129 KB of partially written rows must fit L2)
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
  for (int i = lane; i < MATZ_ROW; i += 32) stage[MATZ_NBUF * 32 * MATZ_ROW + i] = 0.0;
#ifdef MATZ_BULK
  rq_fence_proxy_async_smem();
#endif
  __syncwarp();
  RqWalkCtx C;
  C.pj = T->pj; C.ops = T->ops; C.nops = P.nops; C.ms = P.ms; C.nq = P.nq; C.nv = P.nv; C.nvslots = P.nvslots;
  for (int k = 0; k < 3; ++k) C.gravity[k] = P.gravity[k];
  const int64_t ntiles = (N + 31) / 32;
  const int nb = P.nj - 1;
  // inputs of a tile, staged by the copy engine while the previous tile's patches drain (the per-lane loads at a
  // 48-byte stride were the walk's top stall under the write stream: profiles/ncu_matz_r02_v2.md)
  const bool stin = P.stage_inputs != 0;
  double* in_q = stage + MATZ_STAGE + (MATZ_STAGE & 1);
  double* in_v = in_q + 32 * P.nq;
  double* in_a = in_v + 32 * P.nv;
  uint64_t* bar = reinterpret_cast<uint64_t*>(in_a + 32 * P.nv);
  const bool in16 = ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(qd) | reinterpret_cast<uintptr_t>(qdd)) & 15) == 0;
  if (stin && lane == 0) {
    rq_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  uint32_t parity = 0;
  // a full tile goes through cp.async.bulk (sizes are multiples of 256 bytes); the ragged last one is copied by hand
  auto request = [&](int64_t tl) {
    const int64_t s0 = tl * 32;
    if (s0 + 32 <= N && in16) {
      if (lane == 0) {
        rq_mbar_expect_tx(bar, (uint32_t)(32 * 8 * (P.nq + 2 * P.nv)));
        rq_bulk_g2s(in_q, q + s0 * P.nq, 32 * 8 * P.nq, bar);
        rq_bulk_g2s(in_v, qd + s0 * P.nv, 32 * 8 * P.nv, bar);
        rq_bulk_g2s(in_a, qdd + s0 * P.nv, 32 * 8 * P.nv, bar);
      }
    }
  };
  int buf = 0;
  int64_t tl = (int64_t)blockIdx.x * nw + warp;
  if (stin && tl < ntiles) request(tl);
  for (; tl < ntiles; tl += (int64_t)gridDim.x * nw) {
    const int64_t s0 = tl * 32, smp = s0 + lane;
    const bool valid = smp < N;
    const int nrows = (int)(N - s0 < 32 ? N - s0 : 32);
    if (!stin) {
    } else if (nrows == 32 && in16) {
      while (!rq_mbar_try_wait(bar, parity)) {}
      parity ^= 1u;
    } else {
      for (int i = lane; i < nrows * P.nq; i += 32) in_q[i] = __ldg(q + s0 * P.nq + i);
      for (int i = lane; i < nrows * P.nv; i += 32) { in_v[i] = __ldg(qd + s0 * P.nv + i); in_a[i] = __ldg(qdd + s0 * P.nv + i); }
      __syncwarp();
    }
    const int64_t sc = stin ? (valid ? lane : nrows - 1) : (valid ? smp : N - 1);
    const double* src_v = stin ? in_v : qd;
    const double* src_a = stin ? in_a : qdd;
    MatzSink sink;
    sink.P = &P; sink.T = T; sink.stage = stage; sink.W = W; sink.ldw = ldw; sink.N = N; sink.s0 = s0;
    sink.vs = src_v + sc * P.nv; sink.as = src_a + sc * P.nv; sink.valid = valid; sink.lane = lane; sink.done = 0u; sink.body = 1;
    sink.aligned = (ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 31) == 0;
    sink.aligned16 = (ldw & 1) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0;
    sink.buf = buf;  // the rotation runs across tiles: one __syncwarp lies between two uses of a buffer
    sink.nrows = nrows;
    sink.mode = 0; sink.eg = sink.egp = -2; sink.vv = sink.aa = sink.pv = 0.0;
    if (stin) rq_walk_tile<true>(C, &T->S[0][0], 6, ps, in_q, in_v, in_a, 0, nrows, lane, sink);
    else rq_walk_tile<false>(C, &T->S[0][0], 6, ps, q, qd, qdd, s0, N, lane, sink);
    buf = sink.buf;
    if (P.coupling) {
      // add_coupling_TX40 (regressor.py:216-225): row block 4 <- [a5, v5, sign(v4+v5)], 5 <- [a4, v4, sign(v4+v5)]
      const double v4 = sink.vs[4], v5 = sink.vs[5], a4 = sink.as[4], a5 = sink.as[5];
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
    if (stin) {  // every lane has read its inputs: the next tile may land
      rq_fence_proxy_async_smem();
      __syncwarp();
      const int64_t nx = tl + (int64_t)gridDim.x * nw;
      if (nx < ntiles) request(nx);
    }
  }
#ifdef MATZ_BULK
  rq_bulk_wait_read<0>();  // the copies read this CTA's shared memory
#endif
}

}  // namespace fg
