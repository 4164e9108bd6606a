// fg_blockqr.cuh - kernel (1)+(2), second generation: the regressor tile lives in registers in
// the fp64 tensor-core fragment layout and is folded into R with a blocked (compact-WY)
// Householder update whose trailing work runs on DMMA (mma.sync.m8n8k4.f64).
//
// Geometry (one warp = one tile of S = 4 samples, ms <= 2*NRB row groups, P = 8*NCB columns):
//   lane l: cq = l >> 2 (column inside an 8-column block), rq = l & 3 (sample inside the tile)
//   Bt[cb][rb][e] = W[sample rq, row group 2*rb + e, column 8*cb + cq]
// i.e. the TRANSPOSED tile (columns x rows) stored as the C fragment of an 8x8 DMMA block with
// M = column, N = row (row n = 2*rq + e of row block rb).  In that layout one register is at the
// same time
//   * an A operand (M = column, K = row)      -> contraction over rows:  Y^T = R_kj^T + B_j^T Xh
//   * a  B operand (K = row,    N = reflector) -> the panel's reflectors Xh need no data movement
// and the 8x8 results (Y^T, W^T) come out as C fragments that are again valid A operands
// (K = reflector), so  W^T = Y^T T  and  B_j^T -= W^T Xh^T  chain through registers.  Only
// Xh^T (8x8 per row block, once per panel) is transposed through shared memory.
//
// Panel factorisation (8 reflectors, LAPACK dlarfg convention) keeps the reductions inside the
// 4 lanes that share a column (two xor-shuffles); T comes from T^-1 = diag(1/tau) + striu(Xh^T Xh)
// by an 8-step column back substitution done redundantly by the 4 lanes of each column.
//
// R is shared by WPR warps of a CTA and handed from tile to tile one block row at a time
// (ticket order, deterministic): tile t may touch block row k as soon as tile t-1 released it.
#pragma once
#include "fg_device.cuh"

namespace fg {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

constexpr int BQ_S = 4;         // samples per tile
constexpr int BQ_VROW = 10;     // row stride (doubles) of the transposed-reflector scratch
__host__ __device__ constexpr int BQ_XB(int nrb) { return nrb * 8 + 4; }  // pivot column + 4 partial norms

// R is stored as NCB block rows; block row k keeps only its upper-triangular part's blocks k..NCB-1:
// an 8 x 8(NCB-k) row-major panel (+2 doubles of row padding so that the 4 row pairs a warp
// touches in one access fall into distinct banks).
__host__ __device__ constexpr int bq_ld(int ncb, int k) { return 8 * (ncb - k) + 2; }
__host__ __device__ constexpr int bq_roff(int ncb, int k) {  // doubles before block row k
  return 8 * (k * (8 * ncb + 2) - 8 * (k * (k - 1) / 2));
}

struct BqLayout {
  int off_plan, off_joints, off_rows, off_cols, off_nactk, off_flags, off_R, off_warp0;
  int r_stride;     // bytes per shared R
  int warp_stride;  // bytes per warp
  int w_fw, w_vs, w_as, w_x, w_vt, w_s;
  int ldr;          // leading dimension of R (doubles)
  int total;
};

__host__ __device__ inline BqLayout bq_layout(int nj, int ms, int nv, int ncb, int nrb, int nwarps, int nR) {
  BqLayout L;
  const int P = 8 * ncb;
  int o = 0;
  L.off_plan = o;   o = align16(o + (int)sizeof(Plan));
  L.off_joints = o; o = align16(o + nj * (int)sizeof(DevJoint));
  L.off_rows = o;   o = align16(o + ms * (int)sizeof(RowDesc));
  L.off_cols = o;   o = align16(o + P * (int)sizeof(ColDesc));
  L.off_nactk = o;  o = align16(o + ncb * (int)sizeof(int));
  L.off_flags = o;  o = align16(o + nR * ncb * (int)sizeof(int));
  L.ldr = 0;  // block row k is an 8 x 8(ncb-k) row-major panel with leading dimension bq_ld(ncb, k)
  L.r_stride = align16(bq_roff(ncb, ncb) * 8);
  L.off_R = o;      o += nR * L.r_stride;
  L.off_warp0 = o;
  int w = 0;
  L.w_fw = w; w = align16(w + BQ_S * nj * FW_STRIDE * 8);
  L.w_vs = w; w = align16(w + BQ_S * nv * 8);
  L.w_as = w; w = align16(w + BQ_S * nv * 8);
  L.w_x = w;  w = align16(w + 2 * BQ_XB(nrb) * 8);       // double-buffered pivot column
  L.w_vt = w; w = align16(w + nrb * 8 * BQ_VROW * 8);    // -Xh^T per row block
  L.w_s = w;  w = align16(w + 64 * 8);                   // S[m][i] = xh_i . xh_m (i < m), S[m][m] = tau_m
  L.warp_stride = w;
  L.total = o + nwarps * w;
  return L;
}

// ---------------------------------------------------------------------------------------------
// Fold one register tile into the shared R.  na_of_k[k] = number of leading row blocks that can be
// non-zero in panel k (block-triangular structure of the regressor; NRB for dense tiles).
//
// The panel loop is NOT unrolled: after panel k the column blocks are shifted down one slot
// (Bt[j] <- Bt[j+1], ~120 register moves), so the current panel is always Bt[0] and one compact
// copy of the panel code (a few hundred instructions) serves all NCB panels.  A fully unrolled
// version measured 56% of its issue slots lost to instruction-cache misses (profiles/).
template <int NCB, int NRB>
__device__ __forceinline__ void bq_fold_tile(double (&Bt)[NCB][NRB][2], double* __restrict__ R,
                                             volatile int* flags, int ticket, const int* na_of_k, double* xbuf,
                                             double* vtbuf, double* sbuf, int lane) {
  const int cq = lane >> 2, rq = lane & 3;
#pragma unroll 1
  for (int k = 0; k < NCB; ++k) {
    const int na = na_of_k[k];
    // ---- acquire block row k of R
    if (lane == 0) {
      while (flags[k] != ticket) {}
      __threadfence_block();
    }
    __syncwarp();
    const int ldr = bq_ld(NCB, k);
    if (na > 0) {
      double* Rk = R + bq_roff(NCB, k);  // block row k; its diagonal block comes first
      // ---- panel: 8 Householder reflectors.  The pivot columns stay UNnormalised (x, not x/den)
      // until the panel is done; 1/den of the lane's own column is kept in `myinv`.  Row blocks
      // beyond `na` hold exact zeros, so the link arithmetic simply runs over all of them.
      double myinv = 0.0;
      double* Rrow = Rk;
#pragma unroll 1
      for (int j = 0; j < 8; ++j, Rrow += ldr) {
        const double alpha = Rrow[j];
        const double rho = Rrow[cq];
        double* xb = xbuf + (j & 1) * BQ_XB(NRB);
        double ss0 = 0.0, ss1 = 0.0;
#pragma unroll
        for (int rb = 0; rb < NRB; ++rb) {
          ss0 = fma(Bt[0][rb][0], Bt[0][rb][0], ss0);
          ss1 = fma(Bt[0][rb][1], Bt[0][rb][1], ss1);
        }
        if (cq == j) {  // publish the pivot column and the 4 partial sums of squares
#pragma unroll
          for (int rb = 0; rb < NRB; ++rb)
            *reinterpret_cast<double2*>(xb + rb * 8 + 2 * rq) = make_double2(Bt[0][rb][0], Bt[0][rb][1]);
          xb[NRB * 8 + rq] = ss0 + ss1;
        }
        __syncwarp();
        const double2 sp0 = *reinterpret_cast<const double2*>(xb + NRB * 8);
        const double2 sp1 = *reinterpret_cast<const double2*>(xb + NRB * 8 + 2);
        const double sigma = (sp0.x + sp0.y) + (sp1.x + sp1.y);
        if (sigma == 0.0) {  // nothing to annihilate: H = I (dlarfg with xnorm == 0), uniform branch
          if (rq == 0 && cq <= j) sbuf[j * 8 + cq] = 0.0;
        } else {
          double xv[NRB][2];
          double d0 = 0.0, d1 = 0.0;
#pragma unroll
          for (int rb = 0; rb < NRB; ++rb) {
            const double2 t2 = *reinterpret_cast<const double2*>(xb + rb * 8 + 2 * rq);
            xv[rb][0] = t2.x; xv[rb][1] = t2.y;
            d0 = fma(t2.x, Bt[0][rb][0], d0);
            d1 = fma(t2.y, Bt[0][rb][1], d1);
          }
          double d = d0 + d1;
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          d += __shfl_xor_sync(0xffffffffu, d, 2);
          const double mu = fma(alpha, alpha, sigma);
          const double rs = rsqrt(mu);
          const double nrm = mu * rs;
          const double beta = alpha >= 0.0 ? -nrm : nrm;
          const double den = alpha - beta;
          const double inv = __drcp_rn(den);
          const double tauh = fabs(den) * rs;   // (beta - alpha) / beta
          const double di = d * inv;
          const double w = tauh * (di + rho);   // tau * v^T [rho; b],  v = [1; x * inv]
          const double coef = cq > j ? -(w * inv) : 0.0;
#pragma unroll
          for (int rb = 0; rb < NRB; ++rb) {
            Bt[0][rb][0] = fma(coef, xv[rb][0], Bt[0][rb][0]);
            Bt[0][rb][1] = fma(coef, xv[rb][1], Bt[0][rb][1]);
          }
          if (rq == 0) {
            if (cq >= j) Rrow[cq] = cq == j ? beta : rho - w;
            if (cq <= j) sbuf[j * 8 + cq] = cq == j ? tauh : di * myinv;  // tau_j | xh_cq . xh_j
          }
          if (cq == j) myinv = inv;
        }
      }
      if (k + 1 < NCB) {
        // ---- normalised reflectors Xh (B operand of the row contraction) and -Xh^T in B-operand
        // layout (row = cq, reflector pair = rq) for the update; S becomes visible to all lanes
        double xh[NRB][2];
#pragma unroll
        for (int rb = 0; rb < NRB; ++rb) {
          xh[rb][0] = Bt[0][rb][0] * myinv;
          xh[rb][1] = Bt[0][rb][1] * myinv;
          if (rb < na) {
            vtbuf[rb * 8 * BQ_VROW + (2 * rq) * BQ_VROW + cq] = -xh[rb][0];
            vtbuf[rb * 8 * BQ_VROW + (2 * rq + 1) * BQ_VROW + cq] = -xh[rb][1];
          }
        }
        __syncwarp();
        double vt[NRB][2];
#pragma unroll
        for (int rb = 0; rb < NRB; ++rb)
          if (rb < na) {
            const double2 t2 = *reinterpret_cast<const double2*>(vtbuf + rb * 8 * BQ_VROW + cq * BQ_VROW + 2 * rq);
            vt[rb][0] = t2.x; vt[rb][1] = t2.y;
          }
        // ---- column cq of T:  (diag(1/tau) + striu(S)) t = e_cq
        double acc[8], t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.0;
#pragma unroll
        for (int m = 7; m >= 0; --m) {
          const double taum = sbuf[m * 8 + m];
          const double tm = (m == cq) ? taum : (m < cq ? -taum * acc[m] : 0.0);
          t[m] = tm;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i < m) acc[i] = fma(sbuf[m * 8 + i], tm, acc[i]);
        }
        const double ts0 = rq == 0 ? t[0] : (rq == 1 ? t[2] : (rq == 2 ? t[4] : t[6]));
        const double ts1 = rq == 0 ? t[1] : (rq == 1 ? t[3] : (rq == 2 ? t[5] : t[7]));
        // ---- trailing column blocks (slot jb holds column block k + jb)
        const int nrem = NCB - k;
#pragma unroll
        for (int jb = 1; jb < NCB; ++jb) {
          if (jb < nrem) {
            double* Rj = Rk + (size_t)(2 * rq) * ldr + 8 * jb + cq;
            const double r0 = Rj[0], r1 = Rj[ldr];
            double y0 = r0, y1 = r1;
#pragma unroll
            for (int rb = 0; rb < NRB; ++rb)
              if (rb < na) {
                dmma884(y0, y1, Bt[jb][rb][0], xh[rb][0]);
                dmma884(y0, y1, Bt[jb][rb][1], xh[rb][1]);
              }
            double w0 = 0.0, w1 = 0.0;
            dmma884(w0, w1, y0, ts0);
            dmma884(w0, w1, y1, ts1);
            Rj[0] = r0 - w0;
            Rj[ldr] = r1 - w1;
#pragma unroll
            for (int rb = 0; rb < NRB; ++rb)
              if (rb < na) {
                dmma884(Bt[jb][rb][0], Bt[jb][rb][1], w0, vt[rb][0]);
                dmma884(Bt[jb][rb][0], Bt[jb][rb][1], w1, vt[rb][1]);
              }
          }
        }
      }
    }
    // ---- release block row k
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      flags[k] = ticket + 1;
    }
    // ---- shift the column blocks down one slot
#pragma unroll
    for (int jb = 0; jb + 1 < NCB; ++jb)
#pragma unroll
      for (int rb = 0; rb < NRB; ++rb) { Bt[jb][rb][0] = Bt[jb + 1][rb][0]; Bt[jb][rb][1] = Bt[jb + 1][rb][1]; }
  }
}

// one regressor column block entry set for this lane; kept out of line so that the kernel holds a
// single copy of the (large) column generator
template <int MS>
__device__ __noinline__ void bq_gen_block(const Plan& P, const DevJoint* J, const RowDesc* rows, const ColDesc* cols,
                                          int col, const double* fw, const double* vs, const double* as,
                                          const double* __restrict__ tau, int64_t ld_tau, int64_t sample, int64_t N,
                                          double* out) {
  const ColDesc cd = cols[col];
  double o[MS];
  gen_column<MS, 1>(P, J, rows, cd, col_ancmask(cd), fw, vs, as, tau, ld_tau, sample, N, o);
#pragma unroll
  for (int g = 0; g < MS; ++g) out[g] = o[g];
}

// ---------------------------------------------------------------------------------------------
// Fused kernel: persistent CTAs, each warp generates tiles of 4 samples and folds them into one
// of the CTA's shared R factors.  parts[blockIdx.x * nR + r] receives the packed factor + norms.
template <int NCB, int NRB, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
bq_accumulate_samples_kernel(DevPlanPtrs pp, const double* __restrict__ q, const double* __restrict__ qd,
                             const double* __restrict__ qdd, const double* __restrict__ tau, int64_t ld_tau,
                             int64_t N, double* __restrict__ parts, int wpr) {
  constexpr int MS = 2 * NRB;
  constexpr int P = 8 * NCB;
  extern __shared__ __align__(16) unsigned char smem[];
  const Plan P0 = *pp.plan;
  const int nwarps = blockDim.x / 32;
  const int nR = nwarps / wpr;
  const BqLayout L = bq_layout(P0.nj, P0.ms, P0.nv, NCB, NRB, nwarps, nR);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  // ---- CTA setup
  if (threadIdx.x == 0) *reinterpret_cast<Plan*>(smem + L.off_plan) = P0;
  {
    DevJoint* J = reinterpret_cast<DevJoint*>(smem + L.off_joints);
    for (int i = threadIdx.x; i < P0.nj * (int)(sizeof(DevJoint) / 8); i += blockDim.x)
      reinterpret_cast<double*>(J)[i] = reinterpret_cast<const double*>(pp.joints)[i];
    RowDesc* Rw = reinterpret_cast<RowDesc*>(smem + L.off_rows);
    for (int i = threadIdx.x; i < P0.ms * (int)(sizeof(RowDesc) / 8); i += blockDim.x)
      reinterpret_cast<double*>(Rw)[i] = reinterpret_cast<const double*>(pp.rows)[i];
    ColDesc* C = reinterpret_cast<ColDesc*>(smem + L.off_cols);
    for (int c = threadIdx.x; c < P; c += blockDim.x) {
      ColDesc cd;
      cd.kind = COL_ZERO; cd.body = 0; cd.k = 0; cd.vidx = 0; cd.grp = -2; cd.anc_lo = 0; cd.anc_hi = 0; cd.pad = 0;
      if (c < P0.p1) cd = pp.cols[c];
      C[c] = cd;
    }
    int* nk = reinterpret_cast<int*>(smem + L.off_nactk);
    for (int k = threadIdx.x; k < NCB; k += blockDim.x) {
      int na = 0;
      for (int rb = 0; rb < NRB; ++rb) {
        const int g = 2 * rb;
        if (g < P0.ms && pp.nact[g] <= 8 * k + 7) na = rb + 1;
      }
      nk[k] = na;
    }
    int* fl = reinterpret_cast<int*>(smem + L.off_flags);
    for (int i = threadIdx.x; i < nR * NCB; i += blockDim.x) fl[i] = 0;
    double* Rall = reinterpret_cast<double*>(smem + L.off_R);
    for (int i = threadIdx.x; i < nR * (L.r_stride / 8); i += blockDim.x) Rall[i] = 0.0;
  }
  __syncthreads();
  const Plan& Pl = *reinterpret_cast<const Plan*>(smem + L.off_plan);
  const DevJoint* J = reinterpret_cast<const DevJoint*>(smem + L.off_joints);
  const RowDesc* rows = reinterpret_cast<const RowDesc*>(smem + L.off_rows);
  const ColDesc* cols = reinterpret_cast<const ColDesc*>(smem + L.off_cols);
  const int* nactk = reinterpret_cast<const int*>(smem + L.off_nactk);
  const int rg = warp / wpr, wi = warp % wpr;
  double* R = reinterpret_cast<double*>(smem + L.off_R + (size_t)rg * L.r_stride);
  volatile int* flags = reinterpret_cast<volatile int*>(smem + L.off_flags) + rg * NCB;
  unsigned char* wb = smem + L.off_warp0 + (size_t)warp * L.warp_stride;
  double* fw = reinterpret_cast<double*>(wb + L.w_fw);
  double* vs = reinterpret_cast<double*>(wb + L.w_vs);
  double* as = reinterpret_cast<double*>(wb + L.w_as);
  double* xbuf = reinterpret_cast<double*>(wb + L.w_x);
  double* vtbuf = reinterpret_cast<double*>(wb + L.w_vt);
  double* sbuf = reinterpret_cast<double*>(wb + L.w_s);
  const int cq = lane >> 2, rq = lane & 3;

  const int64_t ntiles = (N + BQ_S - 1) / BQ_S;
  const int64_t nrg_total = (int64_t)gridDim.x * nR;
  const int64_t rgid = (int64_t)blockIdx.x * nR + rg;
  for (int ticket = wi;; ticket += wpr) {
    const int64_t tile = (int64_t)ticket * nrg_total + rgid;
    if (tile >= ntiles) break;
    const int64_t s0 = tile * BQ_S;
    // ---- forward kinematics of the tile's 4 samples (lanes 0..3)
    if (lane < BQ_S) {
      const int64_t smp = s0 + lane;
      const bool valid = smp < N;
      const int64_t sc = valid ? smp : 0;
      forward_sample(Pl, J, q + sc * Pl.nq, qd + sc * Pl.nv, qdd + sc * Pl.nv, fw + (size_t)lane * Pl.nj * FW_STRIDE,
                     vs + lane * Pl.nv, as + lane * Pl.nv, valid);
    }
    __syncwarp();
    // ---- regressor entries straight into the fragment layout
    double Bt[NCB][NRB][2];
#pragma unroll
    for (int cb = 0; cb < NCB; ++cb) {
      double out[MS];
      bq_gen_block<MS>(Pl, J, rows, cols, 8 * cb + cq, fw + (size_t)rq * Pl.nj * FW_STRIDE, vs + rq * Pl.nv,
                       as + rq * Pl.nv, tau, ld_tau, s0 + rq, N, out);
#pragma unroll
      for (int rb = 0; rb < NRB; ++rb) { Bt[cb][rb][0] = out[2 * rb]; Bt[cb][rb][1] = out[2 * rb + 1]; }
    }
    __syncwarp();
    bq_fold_tile<NCB, NRB>(Bt, R, flags, ticket, nactk, xbuf, vtbuf, sbuf, lane);
  }
  __syncthreads();
  // ---- write the CTA's factors: packed upper triangle (p1 columns) + squared column norms
  const int p1 = Pl.p1;
  const int ntri = p1 * (p1 + 1) / 2;
  for (int r = 0; r < nR; ++r) {
    const double* Rr = reinterpret_cast<const double*>(smem + L.off_R + (size_t)r * L.r_stride);
    double* out = parts + ((size_t)blockIdx.x * nR + r) * part_doubles(p1);
    for (int idx = threadIdx.x; idx < p1 * p1; idx += blockDim.x) {
      const int i = idx / p1, c = idx % p1;
      if (c >= i) out[tri_off(i, p1) + c - i] = Rr[bq_roff(NCB, i >> 3) + (i & 7) * bq_ld(NCB, i >> 3) + c - 8 * (i >> 3)];
    }
    for (int c = threadIdx.x; c < p1; c += blockDim.x) {
      double sacc = 0.0;
      for (int i = 0; i <= c; ++i) {
        const double r = Rr[bq_roff(NCB, i >> 3) + (i & 7) * bq_ld(NCB, i >> 3) + c - 8 * (i >> 3)];
        sacc = fma(r, r, sacc);
      }
      out[ntri + c] = sacc;
    }
  }
}

}  // namespace fg
