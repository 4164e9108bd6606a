// fg_kernels.cuh - the __global__ kernels of the identification path (sm_100a, fp64).
//
//   regressor_materialize_kernel : kernel (1) alone, W written to HBM in the reference layout
//   accumulate_dense_kernel      : TSQR fold of a dense matrix already in HBM
//   merge_pairs_kernel           : pairwise merge of partial R factors (tree reduction step)
//   small_qr_kernel / solve      : kernel (3), the N-independent stage on the (p+1)^2 factor
#pragma once
#include "fg_device.cuh"

namespace fg {

// ------------------------------------------------------------------ shared memory carving
struct SmemLayout {
  int off_plan, off_joints, off_rows, off_nact, off_warp0;  // byte offsets
  int warp_stride;                                          // bytes per warp
  int w_R, w_fw, w_vs, w_as, w_xs;                          // byte offsets inside a warp block
  int total;
};


__host__ __device__ inline SmemLayout make_layout(int nj, int ms, int p1, int nv, int S, int TR, int nwarps,
                                                  bool R_in_smem, bool need_fw) {
  SmemLayout L;
  int o = 0;
  L.off_plan = o;   o = align16(o + (int)sizeof(Plan));
  L.off_joints = o; o = align16(o + nj * (int)sizeof(DevJoint));
  L.off_rows = o;   o = align16(o + ms * (int)sizeof(RowDesc));
  L.off_nact = o;   o = align16(o + (kMaxRowGroups + 1) * (int)sizeof(int));
  L.off_warp0 = o;
  int w = 0;
  L.w_R = w;  w = align16(w + (R_in_smem ? (p1 * (p1 + 1) / 2) * 8 : 0));
  L.w_fw = w; w = align16(w + (need_fw ? S * nj * FW_STRIDE * 8 : 0));
  L.w_vs = w; w = align16(w + (need_fw ? S * nv * 8 : 0));
  L.w_as = w; w = align16(w + (need_fw ? S * nv * 8 : 0));
  L.w_xs = w; w = align16(w + 2 * (TR + 2) * 8);
  L.warp_stride = w;
  L.total = o + nwarps * w;
  return L;
}


__device__ inline void load_plan_to_smem(unsigned char* smem, const SmemLayout& L, const DevPlanPtrs& pp) {
  const Plan P = *pp.plan;
  if (threadIdx.x == 0) *reinterpret_cast<Plan*>(smem + L.off_plan) = P;
  DevJoint* J = reinterpret_cast<DevJoint*>(smem + L.off_joints);
  for (int i = threadIdx.x; i < P.nj * (int)(sizeof(DevJoint) / 8); i += blockDim.x)
    reinterpret_cast<double*>(J)[i] = reinterpret_cast<const double*>(pp.joints)[i];
  RowDesc* Rw = reinterpret_cast<RowDesc*>(smem + L.off_rows);
  for (int i = threadIdx.x; i < P.ms * (int)(sizeof(RowDesc) / 8); i += blockDim.x)
    reinterpret_cast<double*>(Rw)[i] = reinterpret_cast<const double*>(pp.rows)[i];
  int* na = reinterpret_cast<int*>(smem + L.off_nact);
  for (int i = threadIdx.x; i <= P.ms; i += blockDim.x) na[i] = pp.nact[i];
  __syncthreads();
}

// ------------------------------------------------------------------ kernel (1): materialise W
// One warp handles S consecutive samples at a time: lanes 0..S-1 run the forward pass, then
// every lane generates its own columns (lane + 32 t) and stores them; for a fixed row the 32
// lanes write 32 consecutive doubles, so the stores are fully coalesced 256-byte segments.
template <int MSMAX, int S>
__global__ void __launch_bounds__(256) regressor_materialize_kernel(DevPlanPtrs pp, const double* __restrict__ q,
                                                                    const double* __restrict__ qd,
                                                                    const double* __restrict__ qdd, int64_t N,
                                                                    double* __restrict__ W, int64_t ldw) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Plan P0 = *pp.plan;
  const int nwarps = blockDim.x / 32;
  const SmemLayout L = make_layout(P0.nj, P0.ms, P0.p1, P0.nv, S, MSMAX * S, nwarps, false, true);
  load_plan_to_smem(smem, L, pp);
  const Plan& P = *reinterpret_cast<const Plan*>(smem + L.off_plan);
  const DevJoint* J = reinterpret_cast<const DevJoint*>(smem + L.off_joints);
  const RowDesc* rows = reinterpret_cast<const RowDesc*>(smem + L.off_rows);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  unsigned char* wb = smem + L.off_warp0 + (size_t)warp * L.warp_stride;
  double* fw = reinterpret_cast<double*>(wb + L.w_fw);
  double* vs = reinterpret_cast<double*>(wb + L.w_vs);
  double* as = reinterpret_cast<double*>(wb + L.w_as);
  const int64_t ntiles = (N + S - 1) / S;
  const int nslot = (P.ncols + 31) / 32;
  for (int64_t tile = (int64_t)blockIdx.x * nwarps + warp; tile < ntiles; tile += (int64_t)gridDim.x * nwarps) {
    const int64_t s0 = tile * S;
    if (lane < S) {
      const int64_t smp = s0 + lane;
      const bool valid = smp < N;
      const int64_t sc = valid ? smp : 0;
      forward_sample(P, J, q + sc * P.nq, qd + sc * P.nv, qdd + sc * P.nv, fw + (size_t)lane * P.nj * FW_STRIDE,
                     vs + lane * P.nv, as + lane * P.nv, valid);
    }
    __syncwarp();
    for (int t = 0; t < nslot; ++t) {
      const int col = lane + 32 * t;
      ColDesc cd;
      cd.kind = COL_ZERO; cd.body = 0; cd.k = 0; cd.vidx = 0; cd.grp = -2;
      uint64_t anc = 0;
      if (col < P.ncols) {
        cd = pp.cols[col];
        anc = col_ancmask(cd);
      }
      double out[MSMAX * S];
      gen_column<MSMAX, S>(P, J, rows, cd, anc, fw, vs, as, nullptr, 0, s0, N, out);
      if (col < P.ncols) {
        const int fc = pp.fullcol[col];
#pragma unroll
        for (int g = 0; g < MSMAX; ++g) {
          if (g < P.ms) {
            const int64_t rb = rows[g].rowblock;
#pragma unroll
            for (int s = 0; s < S; ++s)
              if (s0 + s < N) W[(rb * N + s0 + s) * ldw + fc] = out[g * S + s];
          }
        }
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ partial factors in HBM
// A "part" is one packed upper-triangular factor (p1 columns) plus p1 squared column norms.

// Fold the rows of packed factor `src` (p1 columns) into `dst` (packed).  One warp.
template <int NSLOT, int TR>
__device__ __forceinline__ void fold_factor(const double* src, double* dst, int p1, double* xs, int lane) {
  for (int r0 = 0; r0 < p1; r0 += TR) {
    double B[TR][NSLOT];
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      const int row = r0 + r;
#pragma unroll
      for (int t = 0; t < NSLOT; ++t) {
        const int col = lane + 32 * t;
        B[r][t] = (row < p1 && col >= row && col < p1) ? src[tri_off(row, p1) + col - row] : 0.0;
      }
    }
    fold_tile<NSLOT, TR, TR>(B, dst, p1, xs, nullptr, r0, lane);
    __syncwarp();
  }
}

// End-of-kernel reduction of the per-warp factors of one CTA, then store to part[blockIdx.x].
template <int NSLOT, int TR>
__device__ __forceinline__ void cta_merge_and_store(double* myR, double* (*warpR)(int, void*), void* ctx,
                                                    const double (&csq)[NSLOT], double* csq_smem /*[nwarps][32*NSLOT]*/,
                                                    int p1, double* xs, double* parts, int lane, int warp, int nwarps) {
  // squared column norms: fixed-order sum over warps
#pragma unroll
  for (int t = 0; t < NSLOT; ++t) csq_smem[(size_t)warp * 32 * NSLOT + 32 * t + lane] = csq[t];
  __syncthreads();
  for (int step = 1; step < nwarps; step *= 2) {
    if ((warp % (2 * step)) == 0 && warp + step < nwarps) fold_factor<NSLOT, TR>(warpR(warp + step, ctx), myR, p1, xs, lane);
    __syncthreads();
  }
  if (warp == 0) {
    double* out = parts + (size_t)blockIdx.x * part_doubles(p1);
    const int ntri = p1 * (p1 + 1) / 2;
    for (int i = lane; i < ntri; i += 32) out[i] = myR[i];
#pragma unroll
    for (int t = 0; t < NSLOT; ++t) {
      const int col = lane + 32 * t;
      if (col < p1) {
        double sacc = 0.0;
        for (int w = 0; w < nwarps; ++w) sacc += csq_smem[(size_t)w * 32 * NSLOT + 32 * t + lane];
        out[ntri + col] = sacc;
      }
    }
  }
}

struct WarpRCtx {
  unsigned char* base;  // smem or global base of warp 0's factor
  size_t stride;        // bytes between consecutive warps' factors
};
__device__ inline double* warpR_fn(int w, void* ctx) {
  WarpRCtx* c = reinterpret_cast<WarpRCtx*>(ctx);
  return reinterpret_cast<double*>(c->base + (size_t)w * c->stride);
}

// ------------------------------------------------------------------ dense TSQR fold
// W is rows x p (ld = ldw) in HBM, tau optional.  Each warp streams tiles of TR rows:
// B[r][t] = W[row0 + r][lane + 32 t] (coalesced), last column = tau.
template <int NSLOT, int TR>
__global__ void __launch_bounds__(256, 1)
accumulate_dense_kernel(const double* __restrict__ W, int64_t ldw, const double* __restrict__ tau, int64_t nrows,
                        int p1, double* __restrict__ parts, double* __restrict__ gR) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int nwarps = blockDim.x / 32;
  const bool R_in_smem = gR == nullptr;
  const int ntri = p1 * (p1 + 1) / 2;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int xs_bytes = align16(2 * (TR + 2) * 8);
  const int r_bytes = R_in_smem ? align16(ntri * 8) : 0;
  const int wstride = xs_bytes + r_bytes;
  unsigned char* wb = smem + (size_t)warp * wstride;
  double* xs = reinterpret_cast<double*>(wb);
  double* Rt = R_in_smem ? reinterpret_cast<double*>(wb + xs_bytes) : gR + ((size_t)blockIdx.x * nwarps + warp) * ntri;
  for (int i = lane; i < ntri; i += 32) Rt[i] = 0.0;
  double csq[NSLOT];
#pragma unroll
  for (int t = 0; t < NSLOT; ++t) csq[t] = 0.0;
  __syncwarp();
  const int p = p1 - 1;
  const int64_t ntiles = (nrows + TR - 1) / TR;
  for (int64_t tile = (int64_t)blockIdx.x * nwarps + warp; tile < ntiles; tile += (int64_t)gridDim.x * nwarps) {
    const int64_t r0 = tile * TR;
    double B[TR][NSLOT];
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      const int64_t row = r0 + r;
#pragma unroll
      for (int t = 0; t < NSLOT; ++t) {
        const int col = lane + 32 * t;
        double val = 0.0;
        if (row < nrows) {
          if (col < p) val = W[row * ldw + col];
          else if (col == p && tau != nullptr) val = tau[row];
        }
        B[r][t] = val;
        csq[t] = fma(val, val, csq[t]);
      }
    }
    fold_tile<NSLOT, TR, TR>(B, Rt, p1, xs, nullptr, 0, lane);
    __syncwarp();
  }
  __shared__ double csq_buf[8 * 32 * NSLOT];
  WarpRCtx ctx;
  if (R_in_smem) { ctx.base = smem + xs_bytes; ctx.stride = (size_t)wstride; }
  else { ctx.base = reinterpret_cast<unsigned char*>(gR + (size_t)blockIdx.x * nwarps * ntri); ctx.stride = (size_t)ntri * 8; }
  __threadfence_block();
  cta_merge_and_store<NSLOT, TR>(Rt, warpR_fn, &ctx, csq, csq_buf, p1, xs, parts, lane, warp, nwarps);
}

// ------------------------------------------------------------------ pairwise merge of parts
// CTA b (one warp) merges part[(2b+1)*stride] into part[2b*stride] (if it exists).
template <int NSLOT, int TR>
__global__ void __launch_bounds__(32) merge_pairs_kernel(double* __restrict__ parts, int nparts, int stride, int p1,
                                                         int use_smem) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x;
  const int a = 2 * blockIdx.x * stride, b = a + stride;
  if (b >= nparts) return;
  const int ntri = p1 * (p1 + 1) / 2;
  double* pa = parts + (size_t)a * part_doubles(p1);
  double* pb = parts + (size_t)b * part_doubles(p1);
  double* xs = reinterpret_cast<double*>(smem);
  double* Rt = pa;
  if (use_smem) {
    Rt = reinterpret_cast<double*>(smem + align16(2 * (TR + 2) * 8));
    for (int i = lane; i < ntri; i += 32) Rt[i] = pa[i];
    __syncwarp();
  }
  fold_factor<NSLOT, TR>(pb, Rt, p1, xs, lane);
  if (use_smem) {
    __syncwarp();
    for (int i = lane; i < ntri; i += 32) pa[i] = Rt[i];
  }
  for (int i = lane; i < p1; i += 32) pa[ntri + i] += pb[ntri + i];
}

// ------------------------------------------------------------------ squared column norms of a dense matrix
// diag(W^T W) of get_index_eliminate / eliminate_non_dynaffect (regressor.py:243,271) for ANY width: CTA b sums the
// squares of its row slab per column (thread = column, coalesced rows), a second launch adds the slabs in order.
__global__ void __launch_bounds__(256) colsq_partial_kernel(const double* __restrict__ W, int64_t ldw, int64_t rows,
                                                            int p, double* __restrict__ partial) {
  const int64_t slab = (rows + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = (int64_t)blockIdx.y * slab, r1 = r0 + slab < rows ? r0 + slab : rows;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p) return;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int64_t r = r0;
  for (; r + 3 < r1; r += 4) {
    const double x0 = W[r * ldw + c], x1 = W[(r + 1) * ldw + c], x2 = W[(r + 2) * ldw + c], x3 = W[(r + 3) * ldw + c];
    a0 = fma(x0, x0, a0); a1 = fma(x1, x1, a1); a2 = fma(x2, x2, a2); a3 = fma(x3, x3, a3);
  }
  for (; r < r1; ++r) { const double x = W[r * ldw + c]; a0 = fma(x, x, a0); }
  partial[(size_t)blockIdx.y * p + c] = (a0 + a1) + (a2 + a3);
}
__global__ void colsq_final_kernel(const double* __restrict__ partial, int nslab, int p, double* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p) return;
  double acc = 0.0;
  for (int s = 0; s < nslab; ++s) acc += partial[(size_t)s * p + c];
  out[c] = acc;
}

// ------------------------------------------------------------------ packing helpers
__global__ void tri_to_dense_kernel(const double* __restrict__ part, int p1, double* __restrict__ dense,
                                    double nrows_as_double) {
  const int ntri = p1 * (p1 + 1) / 2;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < p1 * p1; idx += gridDim.x * blockDim.x) {
    const int r = idx / p1, c = idx % p1;
    dense[idx] = c >= r ? part[tri_off(r, p1) + c - r] : 0.0;
  }
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < p1; i += blockDim.x) dense[(size_t)p1 * p1 + i] = part[ntri + i];
    if (threadIdx.x == 0) dense[(size_t)p1 * p1 + p1] = nrows_as_double;
    // column order of the factor (always the caller's order since generation 5; kept in the exchange format)
    for (int i = threadIdx.x; i < p1 - 1; i += blockDim.x) dense[(size_t)p1 * p1 + p1 + 1 + i] = (double)i;
  }
}

// packed exchange buffers (dense square + norms + rows + column order) -> parts (packed triangles); the row
// counts of the n buffers are added to *rows_acc
__global__ void dense_to_tri_kernel(const double* __restrict__ packed, int n, int p1, double* __restrict__ parts,
                                    double* __restrict__ rows_acc) {
  const int ntri = p1 * (p1 + 1) / 2;
  const int64_t pk = (int64_t)p1 * p1 + 2 * (int64_t)p1;  // fg_state_packed_doubles(p1 - 1)
  for (int k = blockIdx.y; k < n; k += gridDim.y) {
    const double* src = packed + (size_t)k * pk;
    double* dst = parts + (size_t)k * part_doubles(p1);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < p1 * p1; idx += gridDim.x * blockDim.x) {
      const int r = idx / p1, c = idx % p1;
      if (c >= r) dst[tri_off(r, p1) + c - r] = src[idx];
    }
    if (blockIdx.x == 0)
      for (int i = threadIdx.x; i < p1; i += blockDim.x) dst[ntri + i] = src[(size_t)p1 * p1 + i];
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    double acc = 0.0;
    for (int k = 0; k < n; ++k) acc += packed[(size_t)k * pk + (size_t)p1 * p1 + p1];
    *rows_acc += acc;
  }
}

// ------------------------------------------------------------------ kernel (3): small stage
// A (m rows x n cols, row-major, ld = n) = selected columns of an upper-triangular factor.
// gather: A[r][j] = (r <= src_col[j]) ? Rsrc[r][src_col[j]] : 0, Rsrc either packed (tri) or dense.
__global__ void gather_cols_kernel(const double* __restrict__ Rsrc, int src_p1, int src_is_tri,
                                   const int* __restrict__ src_col, int m, int n, double* __restrict__ A) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < m * n; idx += gridDim.x * blockDim.x) {
    const int r = idx / n, j = idx % n;
    const int c = src_col[j];
    double v = 0.0;
    if (r <= c) v = src_is_tri ? Rsrc[tri_off(r, src_p1) + c - r] : Rsrc[(size_t)r * src_p1 + c];
    A[idx] = v;
  }
}

// In-place Householder QR of the m x n matrix A (row-major, ld = n) by one CTA; A, exts (n ints) and dg
// (n doubles) may live in shared or global memory.  exts[j] = 1 + last possibly non-zero row of column j on
// entry.  Column c's reflector spans rows c .. E_c-1 with E_c = max(exts[0..c]) (fill-in of earlier
// reflectors).  On exit the upper triangle of A holds R (first min(m,n) rows, diagonal from dg) and every
// entry below the diagonal is zero.  Contains __syncthreads(): all threads of the CTA must call it.
__device__ inline void small_qr_body(double* A, int m, int n, const int* exts, double* dg) {
  const int kmax = m < n ? m : n;
  int E = 0;
  for (int c = 0; c < kmax; ++c) {
    const int e = exts[c] < m ? exts[c] : m;
    E = e > E ? e : E;
    if (E < c + 1) E = c + 1;
    const double alpha = A[(size_t)c * n + c];
    double sigma = 0.0;
    for (int r = c + 1; r < E; ++r) { const double x = A[(size_t)r * n + c]; sigma = fma(x, x, sigma); }
    if (sigma == 0.0) {
      if (threadIdx.x == 0) dg[c] = alpha;
      continue;  // uniform: every thread computed the same sigma
    }
    const double nrm = sqrt(fma(alpha, alpha, sigma));
    const double beta = alpha >= 0.0 ? -nrm : nrm;
    const double inv = 1.0 / (alpha - beta);
    const double tauh = (beta - alpha) / beta;
    for (int j = c + 1 + threadIdx.x; j < n; j += blockDim.x) {
      double d = 0.0;
      for (int r = c + 1; r < E; ++r) d = fma(A[(size_t)r * n + c], A[(size_t)r * n + j], d);
      const double rho = A[(size_t)c * n + j];
      const double sdot = fma(d, inv, rho);
      A[(size_t)c * n + j] = fma(-tauh, sdot, rho);
      const double g = tauh * sdot * inv;
      for (int r = c + 1; r < E; ++r) A[(size_t)r * n + j] = fma(-g, A[(size_t)r * n + c], A[(size_t)r * n + j]);
    }
    if (threadIdx.x == 0) dg[c] = beta;
    __syncthreads();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < m * n; i += blockDim.x) {
    const int r = i / n, c = i % n;
    if (r > c) A[i] = 0.0;
    else if (r == c && r < kmax) A[i] = dg[r];
  }
  __syncthreads();
}

// The matrix is staged in shared memory when `use_smem` (m*n*8 + n*8 + n*4 bytes).
__global__ void __launch_bounds__(1024) small_qr_kernel(double* __restrict__ Ag, int m, int n,
                                                        const int* __restrict__ ext, int use_smem) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* A = Ag;
  double* dg;
  if (use_smem) {
    A = reinterpret_cast<double*>(smem);
    dg = A + (size_t)m * n;
    for (int i = threadIdx.x; i < m * n; i += blockDim.x) A[i] = Ag[i];
  } else {
    dg = reinterpret_cast<double*>(smem);
  }
  int* exts = reinterpret_cast<int*>(dg + n);
  for (int i = threadIdx.x; i < n; i += blockDim.x) exts[i] = ext[i];
  __syncthreads();
  small_qr_body(A, m, n, exts, dg);
  if (use_smem)
    for (int i = threadIdx.x; i < m * n; i += blockDim.x) Ag[i] = A[i];
}

// Back substitution of one right-hand side by one warp (r <= 32 US_M = 256), see upper_solve_kernel below.
constexpr int US_M = 8;
__device__ __forceinline__ void upper_solve_warp(const double* Rg, int ldr, int r, int j, int nrhs, double* X, int lane) {
  double b[US_M], rc[US_M];
#pragma unroll
  for (int m = 0; m < US_M; ++m) {
    const int k = lane + 32 * m;
    b[m] = k < r ? Rg[(size_t)k * ldr + r + j] : 0.0;
    rc[m] = k < r ? Rg[(size_t)k * ldr + (r - 1)] : 0.0;
  }
  for (int i = r - 1; i >= 0; --i) {
    double rn[US_M];
#pragma unroll
    for (int m = 0; m < US_M; ++m) {
      const int k = lane + 32 * m;
      rn[m] = (i > 0 && k < i) ? Rg[(size_t)k * ldr + (i - 1)] : 0.0;
    }
    const int s = i >> 5;
    double bs = b[0], rs = rc[0];
#pragma unroll
    for (int m = 1; m < US_M; ++m)
      if (s == m) { bs = b[m]; rs = rc[m]; }
    const double xi = __shfl_sync(0xffffffffu, bs, i & 31) / __shfl_sync(0xffffffffu, rs, i & 31);
    if (lane == 0) X[(size_t)i * nrhs + j] = xi;
#pragma unroll
    for (int m = 0; m < US_M; ++m) {
      const int k = lane + 32 * m;
      if (k < i) b[m] = fma(-rc[m], xi, b[m]);
      rc[m] = rn[m];
    }
  }
}

// ------------------------------------------------------------------ kernel (3) in one launch
// get_index_eliminate + get_baseIndex + double_QR on the factor, one CTA, everything in shared memory:
//   keep      = columns with colsq >= tol_e (regressor.py:271-274) that the caller did not drop,
//   Re        = qr(R[:, keep + rhs])                       -> |diag| decides base / regroup (qrdecomposition.py:117-121)
//   Rg        = qr(Re[:, base + regroup + rhs])            -> X = R1^-1 [R2 | Q1^T tau]  (beta, phi_b)
// Output blob (doubles): [0] nk, [1] r, [2..3] unused, colsq[p1], diag[p1], order[p1] (kept-column indices:
// base first, then regroup), Rg[n*n] (n = nk + 1), X[r * (nk - r + 1)].  Re is also written to Re_out (n*n).
__global__ void __launch_bounds__(256) fused_solve_kernel(const double* __restrict__ part, int p1,
                                                          const int* __restrict__ force_drop, double tol_e, double tol_qr,
                                                          double* __restrict__ Re_out, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int p = p1 - 1, ntri = p1 * (p1 + 1) / 2;
  double* A = reinterpret_cast<double*>(smem);          // p1 x n, then the reduced factor in its first n rows
  double* A2 = A + (size_t)p1 * p1;                      // n x n regrouped
  double* dg = A2 + (size_t)p1 * p1;                     // p1
  int* cols = reinterpret_cast<int*>(dg + p1);           // p1: source column of every gathered column
  int* exts = cols + p1;                                 // p1
  int* order = exts + p1;                                // p1
  __shared__ int s_nk, s_r;
  double* o_colsq = out + 4;
  double* o_diag = o_colsq + p1;
  double* o_order = o_diag + p1;
  double* o_Rg = o_order + p1;
  for (int c = threadIdx.x; c < p1; c += blockDim.x) {
    const double sq = part[ntri + c];
    o_colsq[c] = sq;
    order[c] = c < p && ((force_drop != nullptr && force_drop[c] != 0) || (tol_e >= 0.0 && sq < tol_e));  // drop flag
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int nk = 0;
    for (int c = 0; c < p; ++c)
      if (!order[c]) { cols[nk] = c; exts[nk] = c + 1; ++nk; }
    cols[nk] = p; exts[nk] = p1;
    s_nk = nk;
  }
  __syncthreads();
  const int nk = s_nk, n = nk + 1;
  for (int idx = threadIdx.x; idx < p1 * n; idx += blockDim.x) {
    const int r = idx / n, j = idx % n, c = cols[j];
    A[idx] = r <= c ? part[tri_off(r, p1) + c - r] : 0.0;
  }
  __syncthreads();
  small_qr_body(A, p1, n, exts, dg);
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) Re_out[idx] = A[idx];
  for (int j = threadIdx.x; j < n; j += blockDim.x) o_diag[j] = fabs(A[(size_t)j * n + j]);
  if (threadIdx.x == 0) {
    int r = 0;
    for (int i = 0; i < nk; ++i) if (fabs(A[(size_t)i * n + i]) > tol_qr) order[r++] = i;
    s_r = r;
    int k = r;
    for (int i = 0; i < nk; ++i) if (!(fabs(A[(size_t)i * n + i]) > tol_qr)) order[k++] = i;
    order[nk] = nk;
    out[0] = nk; out[1] = r; out[2] = 0.0; out[3] = 0.0;
  }
  __syncthreads();
  const int r = s_r;
  for (int j = threadIdx.x; j < n; j += blockDim.x) { o_order[j] = order[j]; exts[j] = order[j] + 1; }
  if (r == 0) return;
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
    const int i = idx / n, j = idx % n, c = order[j];
    A2[idx] = i <= c ? A[(size_t)i * n + c] : 0.0;
  }
  __syncthreads();
  small_qr_body(A2, n, n, exts, dg);
  for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) o_Rg[idx] = A2[idx];
  const int nrhs = nk - r + 1;
  double* X = o_Rg + (size_t)n * n;
  const int lane = threadIdx.x & 31;
  for (int j = threadIdx.x >> 5; j < nrhs; j += blockDim.x >> 5) upper_solve_warp(A2, n, r, j, nrhs, X, lane);
}

// Back substitution on the regrouped factor Rg ((nk+1) x (nk+1) dense, ld = ldr):
//   X[:, j] = R1^-1 Rg[0:r, r + j]  for j in [0, nrhs)   (nrhs = n_regroup + 1, last = rhs)
// X is r x nrhs row-major.  One thread per right-hand side (any r).
__global__ void upper_solve_serial_kernel(const double* __restrict__ Rg, int ldr, int r, int nrhs,
                                          double* __restrict__ X) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nrhs) return;
  for (int i = r - 1; i >= 0; --i) {
    double acc = Rg[(size_t)i * ldr + r + j];
    for (int k = i + 1; k < r; ++k) acc = fma(-Rg[(size_t)i * ldr + k], X[(size_t)k * nrhs + j], acc);
    X[(size_t)i * nrhs + j] = acc / Rg[(size_t)i * ldr + i];
  }
}

// Same result, one WARP per right-hand side, column-oriented (r <= 128): lane l keeps b[l], b[l + 32], ...;
// step i divides b_i by R_ii and every lane updates its own entries with column i of R1, which was requested
// one step earlier.  r dependent steps of ~100 cycles instead of r^2 / 2 dependent global round trips.
__global__ void __launch_bounds__(128) upper_solve_kernel(const double* __restrict__ Rg, int ldr, int r, int nrhs,
                                                          double* __restrict__ X) {
  const int lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= nrhs) return;
  upper_solve_warp(Rg, ldr, r, j, nrhs, X, lane);
}

// ------------------------------------------------------------------ column-pivoted QR of the small factor
// scipy.linalg.qr(W_e, pivoting=True) of QR_pivoting (qrdecomposition.py:38) on the (reduced) R factor: since
// R^T R = W_e^T W_e, the partial column norms LAPACK's dgeqp3 pivots on are the same for R as for W_e.
// A: m x n row-major in global memory (staged in shared memory); the first npiv columns take part in the
// pivoting, the others (the rhs) are carried along.  perm[j] = source column of pivoted column j.
// Partial norms are recomputed exactly at every step (dgeqp3 downdates them); ties go to the lowest index.
__global__ void __launch_bounds__(256) pivot_qr_kernel(double* __restrict__ Ag, int m, int n, int npiv,
                                                       int* __restrict__ perm_out) {
  extern __shared__ __align__(16) unsigned char smem[];
  double* A = reinterpret_cast<double*>(smem);
  double* nrm = A + (size_t)m * n;             // n
  int* perm = reinterpret_cast<int*>(nrm + n);  // n
  __shared__ int s_pvt;
  for (int i = threadIdx.x; i < m * n; i += blockDim.x) A[i] = Ag[i];
  for (int j = threadIdx.x; j < n; j += blockDim.x) perm[j] = j;
  __syncthreads();
  const int kmax = m < npiv ? m : npiv;
  for (int c = 0; c < kmax; ++c) {
    for (int j = c + threadIdx.x; j < npiv; j += blockDim.x) {
      double sacc = 0.0;
      for (int r = c; r < m; ++r) sacc = fma(A[(size_t)r * n + j], A[(size_t)r * n + j], sacc);
      nrm[j] = sacc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int best = c;
      for (int j = c + 1; j < npiv; ++j) if (nrm[j] > nrm[best]) best = j;
      s_pvt = best;
      const int t = perm[c]; perm[c] = perm[best]; perm[best] = t;
    }
    __syncthreads();
    const int pv = s_pvt;
    if (pv != c)
      for (int r = threadIdx.x; r < m; r += blockDim.x) {
        const double t = A[(size_t)r * n + c]; A[(size_t)r * n + c] = A[(size_t)r * n + pv]; A[(size_t)r * n + pv] = t;
      }
    __syncthreads();
    const double alpha = A[(size_t)c * n + c];
    double sigma = 0.0;
    for (int r = c + 1; r < m; ++r) { const double x = A[(size_t)r * n + c]; sigma = fma(x, x, sigma); }
    if (sigma != 0.0) {  // uniform
      const double nr = sqrt(fma(alpha, alpha, sigma));
      const double beta = alpha >= 0.0 ? -nr : nr;
      const double inv = 1.0 / (alpha - beta);
      const double tauh = (beta - alpha) / beta;
      for (int j = c + 1 + threadIdx.x; j < n; j += blockDim.x) {
        double d = 0.0;
        for (int r = c + 1; r < m; ++r) d = fma(A[(size_t)r * n + c], A[(size_t)r * n + j], d);
        const double rho = A[(size_t)c * n + j];
        const double sdot = fma(d, inv, rho);
        A[(size_t)c * n + j] = fma(-tauh, sdot, rho);
        const double g = tauh * sdot * inv;
        for (int r = c + 1; r < m; ++r) A[(size_t)r * n + j] = fma(-g, A[(size_t)r * n + c], A[(size_t)r * n + j]);
      }
      __syncthreads();
      for (int r = c + 1 + threadIdx.x; r < m; r += blockDim.x) A[(size_t)r * n + c] = 0.0;
      if (threadIdx.x == 0) A[(size_t)c * n + c] = beta;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < m * n; i += blockDim.x) Ag[i] = A[i];
  for (int j = threadIdx.x; j < n; j += blockDim.x) perm_out[j] = perm[j];
}

}  // namespace fg
