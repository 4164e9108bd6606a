// fg_wide.cuh - kernel (1)+(2) for WIDE problems: any kinematic tree, any number of row groups, up to 767
// columns (Talos with free-flyer in external-wrench mode: 6 rows x 330 columns per sample, regressor.py:89-192;
// the full 24-row-block TIAGo tree; the drop-in double_QR on a dense matrix that is already in HBM).
//
// The row-group streams of fg_ringqr.cuh keep one 32 x w_g tile per WARP in shared memory; at w_g = 331 that
// is 85 KB per warp and two warps per SM.  Here the whole CTA works on ONE tile that lives in registers:
//   * 256 threads, thread t owns columns t, t + 256, ...  (NSLOT of them) of a TR-row tile: B[TR][NSLOT];
//   * column step c: the owner of column c takes the Householder reflector of [R_cc; B[:, c]] (LAPACK dlarfg
//     convention) from its own registers, publishes (v, tau, 1/(alpha-beta)) through shared memory, ONE named
//     barrier, then every thread applies it to its own columns: the dot products and the rank-1 update never
//     leave the thread - no shuffles, no reductions, pure DFMA;
//   * the CTA's triangular factor lives in an L2-resident global scratch (its part slot): row c is read and
//     written exactly once per tile, requested one step ahead;
//   * the regressor entries of a tile are generated column by column (gen_column, fg_device.cuh: body column
//     in the body frame, carried up the tree to the row group's joint) from per-sample forward kinematics that a
//     thread-per-sample pre-kernel leaves in a global buffer (a chunk of samples at a time); the states of the
//     NEXT tile arrive in shared memory by one bulk copy (TMA engine, mbarrier) while the current tile is folded.
#pragma once
#include "fg_device.cuh"
#include "fg_async.cuh"

namespace fg {

constexpr int WD_BAR = 1;         // named barrier of the column owners

// NT: column-owner threads of the CTA (256 with up to three columns each, or 512 with one or two: twice the
// warps for the same tile, which is what a latency-bound column step wants)
template <int NT>
__device__ __forceinline__ void wd_bar() { asm volatile("bar.sync %0, %1;" ::"n"(WD_BAR), "n"(NT) : "memory"); }

// sqrt(x), 1/sqrt(x) and 1/x from a MUFU seed and two Newton steps (full double accuracy for normal x; the
// sequences sqrt() / division expand to are three times as long and sit on the critical path of every column)
__device__ __forceinline__ void wd_sqrt_rsqrt(double x, double& sq, double& rs) {
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
__device__ __forceinline__ double wd_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}

// What the owner of a column publishes for the other threads: xb[0..TR) = x (the column below R's diagonal, not
// normalised), xb[TRP] = alpha (the diagonal entry of R), xb[TRP + 1] = sigma = |x|^2.  EVERY thread derives the
// reflector scalars from (alpha, sigma) itself (LAPACK dlarfg convention: beta = -sign(alpha) sqrt(alpha^2 + sigma),
// v = x / (alpha - beta), tau = (beta - alpha) / beta) while its dot products are in flight, so the square root and
// the reciprocal are not on the owner's serial path (profiles/ncu_wide_r02_v2.md: 48 % of the samples waited at the
// barrier for the owner when it computed them).
template <int TR, int NSLOT>
__device__ __forceinline__ void wide_publish(const double (&B)[TR][NSLOT], int sc, double alpha, double* xb) {
  constexpr int TRP = ((TR + 7) / 8) * 8;
  double sg[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int s = 0; s < NSLOT; ++s)
    if (s == sc) {
#pragma unroll
      for (int r = 0; r < TR; ++r) {
        const double x = B[r][s];
        xb[r] = x;
        sg[r & 3] = fma(x, x, sg[r & 3]);
      }
    }
  xb[TRP] = alpha;
  xb[TRP + 1] = (sg[0] + sg[1]) + (sg[2] + sg[3]);
}

// Apply the published reflector to the thread's column of slot S_ (col > c): dot product, new R entry, rank-1
// update.  PUB: this is the owner's pass over the NEXT column - the updated column and its squared norm go to xn.
template <int TR, int NSLOT, int S_, bool PUB>
__device__ __forceinline__ void wide_apply_slot(double (&B)[TR][NSLOT], const double* xb, double inv, double tauh,
                                                double rho, double* Rdst, bool active, double* xn, double alpha_n) {
  constexpr int TRP = ((TR + 7) / 8) * 8;
  double gco = 0.0;
  if (tauh != 0.0) {  // uniform
    double d[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int r0 = 0; r0 < TR; r0 += 4) {
      const double2 t0 = *reinterpret_cast<const double2*>(xb + r0);
      const double2 t1 = *reinterpret_cast<const double2*>(xb + r0 + 2);
      if (r0 < TR) d[0] = fma(t0.x, B[r0 < TR ? r0 : 0][S_], d[0]);
      if (r0 + 1 < TR) d[1] = fma(t0.y, B[r0 + 1 < TR ? r0 + 1 : 0][S_], d[1]);
      if (r0 + 2 < TR) d[2] = fma(t1.x, B[r0 + 2 < TR ? r0 + 2 : 0][S_], d[2]);
      if (r0 + 3 < TR) d[3] = fma(t1.y, B[r0 + 3 < TR ? r0 + 3 : 0][S_], d[3]);
    }
    const double dd = (d[0] + d[1]) + (d[2] + d[3]);
    const double sdot = fma(dd, inv, rho);  // w^T [rho; b],  w = [1; x * inv]
    if (active) {
      __stcg(Rdst, fma(-tauh, sdot, rho));
      gco = tauh * sdot * inv;
    }
  }
  double sg[4] = {0.0, 0.0, 0.0, 0.0};
  if (tauh != 0.0 || PUB) {
#pragma unroll
    for (int r0 = 0; r0 < TR; r0 += 4) {
      const double2 t0 = *reinterpret_cast<const double2*>(xb + r0);
      const double2 t1 = *reinterpret_cast<const double2*>(xb + r0 + 2);
      const double tv[4] = {t0.x, t0.y, t1.x, t1.y};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (r0 + k < TR) {
          const double nb = fma(-gco, tv[k], B[r0 + k < TR ? r0 + k : 0][S_]);
          B[r0 + k < TR ? r0 + k : 0][S_] = nb;
          if (PUB) {
            xn[r0 + k] = nb;
            sg[k] = fma(nb, nb, sg[k]);
          }
        }
    }
  }
  if (PUB) {
    xn[TRP] = alpha_n;
    xn[TRP + 1] = (sg[0] + sg[1]) + (sg[2] + sg[3]);
  }
}

// all live slots of the thread; `pub_slot`: the slot of column c + 1 if this thread owns it (else -1)
template <int NT, int TR, int NSLOT, int S_>
__device__ __forceinline__ void wide_apply_slots(double (&B)[TR][NSLOT], const double* xb, double inv, double tauh,
                                                 const double (&rho)[NSLOT], const double (&rho_n)[NSLOT], double* Rrow,
                                                 int c, int p1, int tid, int pub_slot, double* xn) {
  if constexpr (S_ < NSLOT) {
    const int col = tid + NT * S_;
    // warp-uniform: a slot whose 32 columns of this warp are all <= c has retired
    const bool warp_live = (col | 31) > c;
    if (warp_live) {
      if (pub_slot == S_)
        wide_apply_slot<TR, NSLOT, S_, true>(B, xb, inv, tauh, rho[S_], Rrow + col - c, col > c && col < p1, xn, rho_n[S_]);
      else
        wide_apply_slot<TR, NSLOT, S_, false>(B, xb, inv, tauh, rho[S_], Rrow + col - c, col > c && col < p1, nullptr, 0.0);
    }
    wide_apply_slots<NT, TR, NSLOT, S_ + 1>(B, xb, inv, tauh, rho, rho_n, Rrow, c, p1, tid, pub_slot, xn);
  }
}

// Fold the register tile B (TR rows; thread `tid` owns columns tid + 256 s) into the packed upper-triangular
// factor Rt (p1 columns, global memory, private to this CTA).  Columns < c_begin of the tile are zero.
// xs: 2 * (ceil8(TR) + 4) doubles of shared memory.
// Step c applies reflector c (published in xs[c & 1]) to every live column.  LOOK-AHEAD: while it updates its
// column c + 1, the owner of that column writes the updated values and their squared norm into the other buffer,
// so the one barrier per step waits for nothing but the updates themselves.
template <int NT, int TR, int NSLOT>
__device__ __forceinline__ void wide_fold(double (&B)[TR][NSLOT], double* __restrict__ Rt, int p1, int c_begin,
                                          double* xs, int tid) {
  constexpr int TRP = ((TR + 7) / 8) * 8;
  constexpr int XS = TRP + 4;
  double rho[NSLOT];
  {
    const int roff = tri_off(c_begin, p1);
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      const int col = tid + NT * s;
      rho[s] = (col >= c_begin && col < p1) ? __ldcg(Rt + roff + col - c_begin) : 0.0;
    }
    if (tid == (c_begin & (NT - 1))) {
      const int sc = c_begin / NT;
      double alpha = 0.0;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s)
        if (s == sc) alpha = rho[s];
      wide_publish<TR, NSLOT>(B, sc, alpha, xs + (c_begin & 1) * XS);
    }
  }
  // R rows are requested TWO steps ahead: the owner of column c + 1 needs R[c+1][c+1] while step c runs, and an
  // L2 round trip (330 cycles idle, more under load) must not sit on the path to the barrier
  double rho_n[NSLOT];
  {
    const int rn = tri_off(c_begin, p1) + (p1 - c_begin);
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      const int col = tid + NT * s;
      rho_n[s] = (c_begin + 1 < p1 && col > c_begin && col < p1) ? __ldcg(Rt + rn + col - c_begin - 1) : 0.0;
    }
  }
  wd_bar<NT>();
  for (int c = c_begin; c < p1; ++c) {
    double rho_nn[NSLOT];
    const int roff = tri_off(c, p1);
    const int rn = roff + (p1 - c);       // tri_off(c + 1, p1)
    const int rnn = rn + (p1 - c - 1);    // tri_off(c + 2, p1)
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      const int col = tid + NT * s;
      rho_nn[s] = (c + 2 < p1 && col > c + 1 && col < p1) ? __ldcg(Rt + rnn + col - c - 2) : 0.0;
    }
    const double* xb = xs + (c & 1) * XS;
    const double alpha = xb[TRP], sigma = xb[TRP + 1];
    double inv = 0.0, tauh = 0.0;
    const double mu = fma(alpha, alpha, sigma);
    if (sigma != 0.0 && mu > 1e-280) {  // uniform; else H = I (dlarfg with xnorm == 0)
      double nrm, rs;
      wd_sqrt_rsqrt(mu, nrm, rs);
      const double beta = alpha >= 0.0 ? -nrm : nrm;
      const double den = alpha - beta;
      inv = wd_rcp(den);
      tauh = fabs(den) * rs;  // (beta - alpha) / beta
      if (tid == (c & (NT - 1))) __stcg(Rt + roff, beta);
    }
    const int cn = c + 1;
    const int pub_slot = (cn < p1 && tid == (cn & (NT - 1))) ? cn / NT : -1;
    wide_apply_slots<NT, TR, NSLOT, 0>(B, xb, inv, tauh, rho, rho_n, Rt + roff, c, p1, tid, pub_slot, xs + (cn & 1) * XS);
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) { rho[s] = rho_n[s]; rho_n[s] = rho_nn[s]; }
    wd_bar<NT>();
  }
}

// ------------------------------------------------------------------ forward kinematics of a chunk of samples
// One thread per sample: the per-joint forward state [nj][FW_STRIDE] (placement in the parent frame, spatial
// velocity, spatial and classical acceleration - pinocchio's JointTorqueRegressorForwardStep) goes to a global
// buffer, sample-major, so the states of the consecutive samples of a tile are ONE contiguous run that the fold
// kernel pulls into shared memory with a single bulk copy.
__global__ void __launch_bounds__(128)
wide_forward_kernel(DevPlanPtrs pp, const double* __restrict__ q, const double* __restrict__ qd,
                    const double* __restrict__ qdd, int64_t n, double* __restrict__ fw_out) {
  extern __shared__ __align__(16) unsigned char wd_smem[];
  const Plan P0 = *pp.plan;
  Plan* Ps = reinterpret_cast<Plan*>(wd_smem);
  double* J = reinterpret_cast<double*>(wd_smem + align16((int)sizeof(Plan)));
  if (threadIdx.x == 0) *Ps = P0;
  for (int i = threadIdx.x; i < P0.nj * (int)(sizeof(DevJoint) / 8); i += blockDim.x) J[i] = reinterpret_cast<const double*>(pp.joints)[i];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  forward_sample(*Ps, reinterpret_cast<const DevJoint*>(J), q + i * P0.nq, qd + i * P0.nv, qdd + i * P0.nv,
                 fw_out + (size_t)i * P0.nj * FW_STRIDE, nullptr, nullptr, true);
}

// ------------------------------------------------------------------ fused: samples -> factor
// grid: persistent; CTA b folds tiles b, b + gridDim.x, ... of the chunk into its own part slot parts[b] (the
// factor stays there between the chunks of one call: `first` zeroes it).  256 threads = column owners.
// shared memory: plan tables | 2 x fw tile [S][nj][27] (bulk-copied, double-buffered) | xs | 2 mbarriers
struct WideLayout {
  int off_plan, off_joints, off_rows, off_fw[2], off_xs, off_bar, total;
};
__host__ __device__ inline WideLayout wide_layout(int nj, int ms, int S, int TR) {
  WideLayout L;
  int o = 0;
  L.off_plan = o;   o = align16(o + (int)sizeof(Plan));
  L.off_joints = o; o = align16(o + nj * (int)sizeof(DevJoint));
  L.off_rows = o;   o = align16(o + ms * (int)sizeof(RowDesc));
  for (int b = 0; b < 2; ++b) { L.off_fw[b] = o; o = align16(o + S * nj * FW_STRIDE * 8); }
  L.off_xs = o; o = align16(o + 2 * (((TR + 7) / 8) * 8 + 4) * 8);
  L.off_bar = o; o = align16(o + 16);
  L.total = o;
  return L;
}

template <int NT, int MSMAX, int S, int NSLOT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
wide_accumulate_samples_kernel(DevPlanPtrs pp, const double* __restrict__ qd, const double* __restrict__ qdd,
                               const double* __restrict__ tau, int64_t ld_tau, int64_t n,
                               const double* __restrict__ fw_all, double* __restrict__ parts, int first, int last) {
  constexpr int TR = MSMAX * S;
  extern __shared__ __align__(16) unsigned char wd_smem[];
  const Plan P0 = *pp.plan;
  const WideLayout L = wide_layout(P0.nj, P0.ms, S, TR);
  const int tid = threadIdx.x;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wd_smem + L.off_bar);
  {
    if (tid == 0) {
      *reinterpret_cast<Plan*>(wd_smem + L.off_plan) = P0;
      rq_mbar_init(bars, 1);
      rq_mbar_init(bars + 1, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      rq_fence_proxy_async();
    }
    double* J = reinterpret_cast<double*>(wd_smem + L.off_joints);
    for (int i = tid; i < P0.nj * (int)(sizeof(DevJoint) / 8); i += blockDim.x) J[i] = reinterpret_cast<const double*>(pp.joints)[i];
    double* Rw = reinterpret_cast<double*>(wd_smem + L.off_rows);
    for (int i = tid; i < P0.ms * (int)(sizeof(RowDesc) / 8); i += blockDim.x) Rw[i] = reinterpret_cast<const double*>(pp.rows)[i];
  }
  const int p1 = P0.p1, ntri = p1 * (p1 + 1) / 2;
  double* Rt = parts + (size_t)blockIdx.x * part_doubles(p1);
  if (first)
    for (int i = tid; i < ntri + p1; i += blockDim.x) Rt[i] = 0.0;
  __syncthreads();
  const Plan& P = *reinterpret_cast<const Plan*>(wd_smem + L.off_plan);
  const DevJoint* J = reinterpret_cast<const DevJoint*>(wd_smem + L.off_joints);
  const RowDesc* rows = reinterpret_cast<const RowDesc*>(wd_smem + L.off_rows);
  double* xs = reinterpret_cast<double*>(wd_smem + L.off_xs);
  const int64_t ntiles = (n + S - 1) / S;
  const int fw_sample = P.nj * FW_STRIDE;  // doubles per sample

  ColDesc cd[NSLOT];
  uint64_t anc[NSLOT];
  double csq[NSLOT];
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) {
    const int col = tid + NT * s;
    cd[s].kind = COL_ZERO; cd[s].body = 0; cd[s].k = 0; cd[s].vidx = 0; cd[s].grp = -2;
    anc[s] = 0;
    csq[s] = 0.0;
    if (col < p1) {
      cd[s] = pp.cols[col];
      anc[s] = col_ancmask(cd[s]);
    }
  }
  auto request = [&](int64_t tile, int buf) {  // thread 0: forward states of the tile's samples -> shared memory
    const int64_t s0 = tile * S;
    const int ns = (int)((n - s0) < S ? (n - s0) : S);
    const uint32_t bytes = (uint32_t)ns * fw_sample * 8;
    rq_mbar_expect_tx(bars + buf, bytes);
    rq_bulk_g2s(wd_smem + L.off_fw[buf], fw_all + (size_t)s0 * fw_sample, bytes, bars + buf);
  };
  int64_t tile = blockIdx.x;
  if (tid == 0 && tile < ntiles) request(tile, 0);
  int buf = 0;
  uint32_t par[2] = {0u, 0u};
  for (; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    // every thread finished reading the other buffer (generation of the previous tile) before the last barrier of
    // its fold: it can be refilled now
    if (tid == 0 && tile + gridDim.x < ntiles) request(tile + gridDim.x, buf ^ 1);
    while (!rq_mbar_try_wait(bars + buf, par[buf])) {}
    par[buf] ^= 1u;
    const double* fw = reinterpret_cast<const double*>(wd_smem + L.off_fw[buf]);
    const int64_t s0 = tile * S;
    double B[TR][NSLOT];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      double out[TR];
      gen_column<MSMAX, S>(P, J, rows, cd[s], anc[s], fw, qd + s0 * P.nv, qdd + s0 * P.nv, tau, ld_tau, s0, n, out);
#pragma unroll
      for (int r = 0; r < TR; ++r) {
        B[r][s] = out[r];
        csq[s] = fma(out[r], out[r], csq[s]);
      }
    }
    wide_fold<NT, TR, NSLOT>(B, Rt, p1, 0, xs, tid);
  }
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) {
    const int col = tid + NT * s;
    if (col < p1) Rt[ntri + col] += csq[s];
  }
  (void)last;
}

// ------------------------------------------------------------------ dense: rows x p matrix in HBM -> factor
template <int NT, int TR, int NSLOT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
wide_accumulate_dense_kernel(const double* __restrict__ W, int64_t ldw, const double* __restrict__ tau,
                             int64_t tau_stride, int64_t nrows, int p1, double* __restrict__ parts) {
  extern __shared__ __align__(16) unsigned char wd_smem[];
  double* xs = reinterpret_cast<double*>(wd_smem);
  const int tid = threadIdx.x;
  const int ntri = p1 * (p1 + 1) / 2, p = p1 - 1;
  double* Rt = parts + (size_t)blockIdx.x * part_doubles(p1);
  for (int i = tid; i < ntri; i += blockDim.x) Rt[i] = 0.0;
  __syncthreads();
  double csq[NSLOT];
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) csq[s] = 0.0;
  const int64_t ntiles = (nrows + TR - 1) / TR;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t r0 = tile * TR;
    double B[TR][NSLOT];
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      const int64_t row = r0 + r;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int col = tid + NT * s;
        double val = 0.0;
        if (row < nrows) {
          if (col < p) val = __ldcs(W + row * ldw + col);
          else if (col == p && tau != nullptr) val = __ldcs(tau + row * tau_stride);
        }
        B[r][s] = val;
        csq[s] = fma(val, val, csq[s]);
      }
    }
    wide_fold<NT, TR, NSLOT>(B, Rt, p1, 0, xs, tid);
  }
#pragma unroll
  for (int s = 0; s < NSLOT; ++s) {
    const int col = tid + NT * s;
    if (col < p1) Rt[ntri + col] = csq[s];
  }
}

// ------------------------------------------------------------------ pairwise merge of parts (one tree level)
// CTA b folds part (2b + 1) * stride into part 2b * stride (packed triangles + norms), if it exists.
template <int NT, int TR, int NSLOT>
__global__ void __launch_bounds__(NT, 1)
wide_merge_pairs_kernel(double* __restrict__ parts, int nparts, int stride, int p1) {
  extern __shared__ __align__(16) unsigned char wd_smem[];
  double* xs = reinterpret_cast<double*>(wd_smem);
  const int tid = threadIdx.x;
  const int a = 2 * blockIdx.x * stride, b = a + stride;
  if (b >= nparts) return;
  const int ntri = p1 * (p1 + 1) / 2;
  double* pa = parts + (size_t)a * part_doubles(p1);
  const double* pb = parts + (size_t)b * part_doubles(p1);
  for (int r0 = 0; r0 < p1; r0 += TR) {
    double B[TR][NSLOT];
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      const int row = r0 + r;
#pragma unroll
      for (int s = 0; s < NSLOT; ++s) {
        const int col = tid + NT * s;
        B[r][s] = (row < p1 && col >= row && col < p1) ? __ldcg(pb + tri_off(row, p1) + col - row) : 0.0;
      }
    }
    wide_fold<NT, TR, NSLOT>(B, pa, p1, r0, xs, tid);
  }
  for (int i = tid; i < p1; i += blockDim.x) pa[ntri + i] += pb[ntri + i];
}

}  // namespace fg
